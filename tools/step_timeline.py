#!/usr/bin/env python
"""Per-launch timeline of one batch step of the captioning path (development tool; run on the GPU box).

    python tools/step_timeline.py [--batch 512] [--out gpurun_out/r02_timeline] [--graph]

Runs warm-up steps, then ONE step under torch.profiler (Kineto / CUPTI activity records: every kernel of the process,
including the ones libgitb200.so launches through ctypes, with start and duration on the GPU clock) and writes
  <out>.csv  : one line per kernel launch (start_us, dur_us, gap_before_us, name)
  <out>.md   : sum of kernel time vs wall time of the step, idle time per kernel boundary summed by (previous, next)
               kernel pair -- i.e. where the part of the step that is in no kernel goes.
CUPTI adds a few microseconds of host cost per launch; the host stays ahead of the GPU at this batch size, so the
GPU-side gaps are the ones an un-profiled run has (the step's wall time under the profiler is printed beside the
un-profiled one for that reason)."""
import argparse
import collections
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def short(name):
    for k in ("gemm2_kernel", "gemm_tcgen05_kernel", "gemv_skinny", "attention_tc_kernel", "text_attention_kernel", "text_attention_combine",
              "layernorm_kernel", "search_step_kernel", "search_rows_kernel", "search_merge_walk_kernel", "im2col_kernel", "store_text_kv_kernel", "embed_text_kernel", "cls_rows_kernel",
              "search_finalize_kernel", "search_init_kernel", "cast_", "preprocess_kernel"):
        if k in name:
            if k == "attention_tc_kernel":
                return "attention_tc<1,true>" if "true" in name or "1, 1" in name or "(bool)1" in name else "attention_tc<0,false>"
            return k
    return name[:40]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_timeline"))
    ap.add_argument("--sweep-rows", type=int, default=-1)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--beam", type=int, default=1)
    ap.add_argument("--max-steps", type=int, default=15)
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    g = importlib.import_module("real-time-video-captioning_b200")
    gm = importlib.import_module("real-time-video-captioning_b200.model")
    dev = torch.device("cuda", 0)
    eng, model = bench.build_engine(g, gm, torch, {"num_image_with_embedding": 6}, 0)
    del model
    sp = g.SearchConfig(beam_size=args.beam, max_steps=args.max_steps)
    B = args.batch
    if args.sweep_rows >= 0:
        eng.set_sweep_rows(args.sweep_rows)
    eng.reserve(B, 6, args.beam, args.max_steps)
    if not args.graph:
        eng.set_graph_segments(False)  # the eager timeline: one launch per kernel
    if args.graph:
        eng.set_graph_max_clips(B)
        eng.set_early_exit(0)
        torch.cuda.set_stream(torch.cuda.Stream(dev))
    frames = torch.randn(B, 6, 3, 224, 224, device=dev, generator=torch.Generator(device=dev).manual_seed(100))
    tok = torch.empty(B, 1, args.max_steps, dtype=torch.int32, device=dev)
    lp = torch.empty(B, 1, dtype=torch.float32, device=dev)
    import ctypes
    c = sp.to_c()
    stream = torch.cuda.current_stream(dev)

    def step():
        rc = eng.lib.gitb200_caption(eng.h, ctypes.c_void_p(frames.data_ptr()), B, 6, ctypes.byref(c), ctypes.c_void_p(tok.data_ptr()),
                                     ctypes.c_void_p(lp.data_ptr()), None, ctypes.c_void_p(stream.cuda_stream))
        assert rc == 0, eng.lib.gitb200_last_error(eng.h)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    plain_ms = e0.elapsed_time(e1) / 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ks = sorted(((e.time_range.start, e.time_range.end - e.time_range.start, e.name) for e in evs), key=lambda x: x[0])
    ks = [k for k in ks if "memcpy" not in k[2].lower() and "memset" not in k[2].lower()]
    if not ks:
        raise SystemExit("no CUDA kernel records: CUPTI unavailable?")
    t_first, t_last = ks[0][0], max(s + d for s, d, _ in ks)
    wall = (t_last - t_first) / 1e3
    busy = sum(d for _, d, _ in ks) / 1e3
    per_kernel = collections.defaultdict(lambda: [0, 0.0])
    gaps = collections.defaultdict(lambda: [0, 0.0])
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".csv", "w") as fh:
        fh.write("start_us,dur_us,gap_before_us,kernel\n")
        prev_end, prev_name = None, None
        for s, d, name in ks:
            n = short(name)
            gap = 0.0 if prev_end is None else max(0.0, s - prev_end)
            fh.write(f"{s - t_first:.2f},{d:.2f},{gap:.2f},{n}\n")
            per_kernel[n][0] += 1
            per_kernel[n][1] += d
            if prev_end is not None:
                gaps[(prev_name, n)][0] += 1
                gaps[(prev_name, n)][1] += gap
            prev_end, prev_name = max(prev_end or 0, s + d), n
    idle = wall - busy
    with open(args.out + ".md", "w") as fh:
        fh.write(f"# Timeline of one {B}-clip step (GIT-base, 6 frames, beam {args.beam} max {args.max_steps}){' -- CUDA graph replay' if args.graph else ''}\n\n")
        fh.write(f"`python tools/step_timeline.py --batch {B} --beam {args.beam} --max-steps {args.max_steps}{' --graph' if args.graph else ''}` (torch.profiler CUDA activity records = CUPTI kernel timestamps)\n\n")
        fh.write(f"* un-profiled step (CUDA events, mean of 3): **{plain_ms:.2f} ms**\n")
        fh.write(f"* profiled step, first kernel start -> last kernel end: {wall:.2f} ms; {len(ks)} kernel launches\n")
        fh.write(f"* sum of kernel durations: **{busy:.2f} ms = {100 * busy / wall:.1f} %** of the profiled step; in no kernel: {idle:.2f} ms ({100 * idle / wall:.1f} %)\n\n")
        fh.write("| kernel | launches | total ms | share of step | mean us |\n|---|---:|---:|---:|---:|\n")
        for n, (cnt, tot) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{n}` | {cnt} | {tot / 1e3:.3f} | {100 * tot / 1e3 / wall:.1f} % | {tot / cnt:.1f} |\n")
        fh.write("\nIdle time by kernel boundary (previous -> next), top 15:\n\n| boundary | count | idle ms | mean us |\n|---|---:|---:|---:|\n")
        for (a, b), (cnt, tot) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:15]:
            fh.write(f"| `{a}` -> `{b}` | {cnt} | {tot / 1e3:.3f} | {tot / cnt:.2f} |\n")
    print(open(args.out + ".md").read())


if __name__ == "__main__":
    main()
