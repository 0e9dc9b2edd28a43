"""Build libb200ring.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python matrix-fhe-lattigo_b200/build.py [--force]

nvcc cross-compiles without a GPU.  Objects go to csrc/_obj/, the library to lib/libb200ring.so
(git-ignored, shipped to the GPU box with the repo snapshot).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "lib", "libb200ring.so")
SOURCES = ["b200ring.cu", "ewise.cu", "ntt_generic.cu", "ntt_big.cu", "keyswitch.cu", "primes.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-ccbin", "g++"]


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    hdr.append(os.path.join(os.path.dirname(HERE), "include", "b200ring.h"))
    return max(os.path.getmtime(h) for h in hdr)


def _compile(src, force, hdr_mtime, tag="", defines=()):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src.replace(".cu", tag + ".o"))
    if not force and os.path.exists(o) and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_mtime):
        return o, False
    subprocess.check_call([NVCC] + FLAGS + ["-D" + d for d in defines] + ["-c", s, "-o", o])
    return o, True


def build(force=False, verbose=True, tag="", defines=()):
    """tag/defines: build an A/B tuning variant lib/libb200ring<tag>.so (select it with B200RING_LIB)."""
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    lib = LIB.replace(".so", tag + ".so")
    hm = _deps()
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        tuned = ("ntt_big.cu", "keyswitch.cu")   # the sources that have compile-time tuning knobs
        res = list(ex.map(lambda s: _compile(s, force, hm, tag if s in tuned else "", defines if s in tuned else ()), SOURCES))
    objs = [o for o, _ in res]
    if force or any(ch for _, ch in res) or not os.path.exists(lib):
        subprocess.check_call([NVCC, "-shared", "-o", lib] + objs + ["-cudart", "static", "-ccbin", "g++", "-Wno-deprecated-gpu-targets"])
        if verbose:
            print("built", lib)
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    tags = [a[6:] for a in sys.argv[1:] if a.startswith("--tag=")]
    build(force="--force" in sys.argv, tag=("_" + tags[0]) if tags else "", defines=defs)
