"""Build liblovr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m law_of_vision_representation_in_mllms_b200.build [--force] [--verbose] [--out PATH -DNAME=1 ...]

`--out` / `-D...` build an experiment variant next to the product library (A/B timing of a compile-time switch; load it
with LOVR_B200_LIB=PATH).  The product library is always liblovr_b200.so with no extra defines.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblovr_b200.so")
SOURCES = ["lovr_b200.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "--use_fast_math=false"]


def _newest_source_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def build(force=False, verbose=False, out=OUT, defines=()):
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _newest_source_mtime():
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + list(defines) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building liblovr_b200.so")
    return out


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else OUT
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=_out,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
