"""In-tree build of libmpcr.so (nvcc, sm_100a only).

The solve kernel has 6 template variants (restoration mode x integrator, three models each); each is its own
translation unit (csrc/mpcr_solve_variant.cu with -D flags) so that they compile in parallel.
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmpcr.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
VARIANTS = [(r, i) for r in (0, 1, 2) for i in (0, 1)]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
        [os.path.join(HERE, "..", "include", "mpcr.h")]


def units():
    """(object name, source, extra flags) of every translation unit."""
    u = [("mpcr.o", os.path.join(CSRC, "mpcr.cu"), [])]
    for r, i in VARIANTS:
        u.append((f"solve_r{r}i{i}.o", os.path.join(CSRC, "mpcr_solve_variant.cu"),
                  [f"-DMPCR_V_RESTO={r}", f"-DMPCR_V_IG={i}"]))
    return u


def build(force=False, verbose=False, jobs=None):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not force and os.path.exists(LIB):
        t = os.path.getmtime(LIB)
        if all(os.path.getmtime(d) <= t for d in _deps()):
            return LIB
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(unit):
        obj, src, flags = unit
        out = os.path.join(OBJ, obj)
        subprocess.check_call([nvcc] + NVCC_FLAGS + extra + flags + ["-c", src, "-o", out])
        return out

    with ThreadPoolExecutor(max_workers=jobs or min(len(units()), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, units()))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
