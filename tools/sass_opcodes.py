#!/usr/bin/env python
"""Per-kernel histogram of the Blackwell-native opcodes in the shipped library (evidence for the tcgen05 / TMA claim).

    python tools/sass_opcodes.py [path/to/libgitb200.so] > profiles/r02_sass_opcodes.md

Disassembles the sm_100a cubin of the .so with `cuobjdump -sass` and counts, per kernel, the SASS mnemonics that only the
5th-generation tensor-core / TMA / TMEM path produces (B200_PROFILING.md): UTCHMMA* (tcgen05.mma), UTMALDG* / UTMASTG*
(cp.async.bulk.tensor load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR* (tcgen05.commit), plus the legacy HMMA.16816
(mma.sync) for contrast.  Runs on the build box: no GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "real-time-video-captioning_b200", "libgitb200.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "ELECT", "MUFU.EX2", "HMMA", "FFMA2", "USETMAXREG"]


def kernel_name(demangled: str) -> str:
    """`void ns::kernel<1, (bool)1>(args...)` -> `kernel<1, (bool)1>` (parameter list = the last top-level parenthesis)."""
    d = demangled.strip()
    if d.endswith(")"):
        depth = 0
        for i in range(len(d) - 1, -1, -1):
            depth += d[i] == ")"
            depth -= d[i] == "("
            if depth == 0:
                d = d[:i]
                break
    return d.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for p in PATTERNS:
            if op.startswith(p):
                kernels[cur][op if p in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "HMMA") else p] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    print("# SASS opcode evidence: " + os.path.basename(LIB) + " (sm_100a), `python tools/sass_opcodes.py`\n")
    print("Counts of instructions per kernel from `cuobjdump -sass`.  `UTCHMMA*` = tcgen05.mma (`.2CTA` = cta_group::2), `UTMALDG*` / `UTMASTG*` = TMA tensor")
    print("load / store, `LDTM` / `STTM` = tcgen05.ld / st (TMEM), `UTCBAR*` = tcgen05.commit, `HMMA.16816` = legacy mma.sync.\n")
    print("| kernel | SASS instr. | tensor-core / TMA / TMEM opcodes |")
    print("|---|---:|---|")
    for k, c in kernels.items():
        ops = {o: n for o, n in c.items() if o != "_total"}
        if not any(o.startswith(("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "UTCBAR")) for o in ops):
            continue
        nm = kernel_name(names[k])
        print(f"| `{nm}` | {c['_total']} | " + ", ".join(f"`{o}` x{n}" for o, n in sorted(ops.items())) + " |")
    others = [kernel_name(names[k]) for k, c in kernels.items()
              if not any(o.startswith(("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "UTCBAR")) for o in c)]
    print(f"\nKernels without tensor-core / TMA opcodes (CUDA-core, HBM-bound or glue): {len(others)} -- " + ", ".join(f"`{o}`" for o in sorted(set(others))))
    ldd = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    libs = sorted({l.split()[0] for l in ldd.splitlines() if l.strip()})
    print("\n`ldd`: " + ", ".join(f"`{l}`" for l in libs) + " -- no cuBLAS / cuDNN / CUTLASS runtime, no torch.")


if __name__ == "__main__":
    main()
