"""In-tree build of libjaxrk_b200.so (hand-written sm_100a kernels + the C ABI).

`python -m jaxrk_b200.build` or `jaxrk_b200.build.build()`.  nvcc cross-compiles without
a GPU; the .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libjaxrk_b200.so")
SOURCES = ["capi.cu", "gram_direct.cu", "gram_tf32.cu", "gram_tf32_pm0.cu", "gram_tf32_pm1.cu", "gram_tf32_pm2.cu",
           "gram_groupsum.cu", "kmatw.cu", "kmatw_pm0.cu", "kmatw_pm1.cu", "kmatw_pm2.cu", "kmatw_pm3.cu",
           "kmatw_tc.cu", "kmatw_tc2.cu", "kmatw_plan.cu", "kgrad.cu", "kgrad_tc.cu", "gram_affine.cu"]
NVCC = os.environ.get("JRK_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("JRK_EXTRA_NVCC_FLAGS", "").split()


PER_FILE_FLAGS = {}      # {source: [extra nvcc flags]}


def _deps_mtime():
    names = [f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    paths = [os.path.join(CSRC, f) for f in names] + [os.path.join(HERE, "..", "include", "jrk.h")]
    return max(os.path.getmtime(p) for p in paths)


def _compile(src, force, hdr_mtime, verbose):
    s = os.path.join(CSRC, src)
    o = os.path.join(OBJ, src.replace(".cu", ".o"))
    if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(s), hdr_mtime):
        return o, ""
    r = subprocess.run([NVCC] + FLAGS + PER_FILE_FLAGS.get(src, []) + ["-c", s, "-o", o], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return o, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr = _deps_mtime()
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr, verbose), SOURCES))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
