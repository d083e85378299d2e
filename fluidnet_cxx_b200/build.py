"""In-tree build of libfluidstep_b200.so (hand-written sm_100a CUDA + the C-ABI).

    python -m fluidnet_cxx_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The .so lands in
fluidnet_cxx_b200/_lib/ (git-ignored, travels to the GPU box with the snapshot).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "libfluidstep_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr"]

# translation unit -> extra flags.  The stencil TUs are bit-exact restatements of ATen's
# op-by-op fp32 arithmetic: no FMA contraction (DESIGN.md "bit-exactness").
UNITS = {
    "stencils.cu": ["-fmad=false"],
    "jacobi_blocked.cu": ["-fmad=false"],
    "step.cu": ["-fmad=false"],
    "step2d.cu": ["-fmad=false"],
    "host_util.cpp": [],
    "halo.cu": [],
    "conv.cu": [],
    "conv_tc.cu": [],
}
OPTIONAL_UNITS = {}


def _deps():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + \
        [os.path.join(os.path.dirname(HERE), "include", "fluidstep.h"), os.path.abspath(__file__)]


def _stamp():
    h = hashlib.sha1()
    for p in _deps():
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def units():
    u = dict(UNITS)
    for k, v in OPTIONAL_UNITS.items():
        if os.path.exists(os.path.join(CSRC, k)):
            u[k] = v
    return u


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp_file = os.path.join(OUT_DIR, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB}")

    def compile_one(item):
        src, extra = item
        obj = os.path.join(OBJ_DIR, src + ".o")
        cmd = [NVCC] + ARCH + COMMON + extra + ["-Xptxas", "-v"] + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        with open(os.path.join(OBJ_DIR, src + ".ptxas.log"), "w") as f:
            f.write(r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, units().items()))
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
