"""Build libbsde_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python bsde-tensor-networks_b200/build.py [--force]

Objects go to bsde-tensor-networks_b200/build/, the library next to this file.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbsde_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["api.cu", "assembly.cu", "dense.cu", "interface.cu", "model.cu", "salsa.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "vh_lstsq.cuh", os.path.join("..", "..", "include", "bsde_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False, defines=(), out=OUT, objdir=OBJ):
    """defines / out / objdir: alternative builds of kernel variants for A/B measurements (bench only)."""
    OBJ = objdir
    os.makedirs(OBJ, exist_ok=True)
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        if not force and _newer(obj, [os.path.join(CSRC, src)] + hdrs):
            return obj
        cmd = [NVCC] + FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(sources)) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or not _newer(out, objs):
        cmd = [NVCC, "-shared", "-o", out] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    if defs:      # variant build: python build.py -DNAME=VALUE ... -> libbsde_b200_variant.so
        print(build(force=True, defines=defs, out=os.path.join(HERE, "libbsde_b200_variant.so"), objdir=os.path.join(HERE, "build_variant")))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
