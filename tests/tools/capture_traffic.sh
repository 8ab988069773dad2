#!/bin/bash
# DRAM traffic + duration per launch of the two roofline kernels at the bench configuration (run only after the plain
# command exited 0): layer-1 GEMMs of the timed step's first sub-batch, and the six decode-attention launches of its
# first decode step (eager launches: --no-graphs, so that -s counts the launches of the step).  Output: gpurun_out/traffic_gemm2.csv, gpurun_out/traffic_tattn.csv
set -e
mkdir -p gpurun_out
python bench.py --quick --no-graphs --steps 1 --warmup 1 > gpurun_out/quick_s3c.log 2>&1
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second"
ncu --metrics $M --clock-control none -k regex:gemm2_kernel -s 284 -c 4 --csv --log-file gpurun_out/traffic_gemm2.csv \
    python bench.py --quick --no-graphs --steps 1 --warmup 1 > gpurun_out/ncu_t1.log 2>&1
ncu --metrics $M --clock-control none -k regex:text_attention_kernel -s 84 -c 6 --csv --log-file gpurun_out/traffic_tattn.csv \
    python bench.py --quick --no-graphs --steps 1 --warmup 1 > gpurun_out/ncu_t2.log 2>&1
tail -1 gpurun_out/quick_s3c.log
