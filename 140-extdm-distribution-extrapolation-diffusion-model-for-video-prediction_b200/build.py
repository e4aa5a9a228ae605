"""Build libextdm_b200.so (all CUDA kernels + the C ABI) for sm_100a, in-tree, with nvcc.

    python build.py            # incremental (per-source object files)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# development only: EXTDM_BUILD_TAG=x EXTDM_NVCC_DEFS="-DFOO" builds libextdm_b200_x.so beside the product (lib.py: EXTDM_LIB)
TAG = os.environ.get("EXTDM_BUILD_TAG", "")
OUT = os.path.join(HERE, f"libextdm_b200{'_' + TAG if TAG else ''}.so")
BUILD = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
SOURCES = ["api.cu", "conv_gemm.cu", "unet_elementwise.cu", "attention.cu", "stw_fused.cu", "stw_tc.cu", "attn_ws32.cu", "attn_core32.cu", "traj.cu", "lfae_cond.cu", "sampler.cu", "warp.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("EXTDM_NVCC_DEFS", "").split()


def _needs(obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    common = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "extdm_b200.h"), __file__]

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        if _needs(obj, [path] + common):
            r = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj], capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
            with open(obj + ".ptxas.log", "w") as f:
                f.write(r.stderr)
            return obj, True
        return obj, False

    with ThreadPoolExecutor(max_workers=6) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(OUT):
        r = subprocess.run([nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                   "-cudart", "static"], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
