"""In-tree nvcc build of rsrl_b200/csrc/librsrl_b200.so for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "librsrl_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]
# --fmad=false for the units whose arithmetic oracle/oracle32.cpp reproduces bit for bit: every FMA there is explicit (hostdev.h)
NOFMAD = ["--fmad=false"]
HEADERS = ["hostdev.h", "device.cuh", "core.cuh", "kernels.cuh", "persistent.cuh", "dyn.cuh", "tile.cuh", "fourier4.cuh", "launch.h", os.path.join("..", "..", "include", "rsrl_b200.h")]


# headers only some translation units include (kept out of HEADERS so that editing them does not rebuild everything)
EXTRA_DEPS = {"abi.cu": ["f4tc_launch.h", "twotable.cuh"], "inst.cu": ["twotable.cuh"], "f4tc_inst.cu": ["f4tc_launch.h", "f4tc.cuh"]}


def _units():
    # RSRL_BUILD_DOMAINS=0 (development only) leaves the CartPole / Acrobot instantiations out for fast iteration;
    # the default builds everything.
    doms = {int(d) for d in os.environ.get("RSRL_BUILD_DOMAINS", "0,1,2").split(",")}
    units = [("abi.o", "abi.cu", NOFMAD), ("f4tc.o", "f4tc_inst.cu", []), ("domains_ex.o", "domains_ex.cu", NOFMAD)]
    for rname, rtype in (("f32", "float"), ("f64", "double")):
        units.append((f"tile_{rname}.o", "tile_inst.cu", [f"-DRSRL_REAL={rtype}", f"-DRSRL_SUFFIX={rname}"]))
        units.append((f"f4_{rname}.o", "f4_inst.cu", [f"-DRSRL_REAL={rtype}", f"-DRSRL_SUFFIX={rname}"]))
        for dom in (0, 1, 2):
            suffix = f"{rname}_d{dom}"
            defs = [f"-DRSRL_REAL={rtype}", f"-DRSRL_DOM={dom}", f"-DRSRL_SUFFIX={suffix}"] + NOFMAD
            if dom not in doms:
                defs.append("-DRSRL_EMPTY")
                units.append((f"inst_{suffix}_empty.o", "inst.cu", defs))
            else:
                units.append((f"inst_{suffix}.o", "inst.cu", defs))
    return units


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    # The library is newer than every source: nothing to do, even when the object files did not travel with the tree
    # (the GPU box receives the prebuilt .so without csrc/build/).
    srcs = sorted({os.path.join(CSRC, u[1]) for u in _units()}) + [os.path.join(CSRC, h) for hs in EXTRA_DEPS.values() for h in hs]
    def _objects_fresh():  # (an RSRL_BUILD_ONLY build relinks the library over objects that are older than the headers)
        for obj, src, _ in _units():
            o = os.path.join(OBJ, obj)
            if os.path.exists(o) and _stale(o, [os.path.join(CSRC, src)] + hdrs + [os.path.join(CSRC, h) for h in EXTRA_DEPS.get(src, [])]):
                return False
        return True
    if not force and not _stale(LIB, srcs + hdrs) and _objects_fresh():
        return LIB
    jobs = []
    # RSRL_BUILD_ONLY=inst_f32_d0,abi (development only): recompile just these objects, trust the others if they exist
    only = [x for x in os.environ.get("RSRL_BUILD_ONLY", "").split(",") if x]
    for obj, src, defs in _units():
        o, s = os.path.join(OBJ, obj), os.path.join(CSRC, src)
        extra = [os.path.join(CSRC, h) for h in EXTRA_DEPS.get(src, [])]
        if only and os.path.exists(o) and not any(obj.startswith(x) for x in only):
            continue
        if force or _stale(o, [s] + hdrs + extra):
            jobs.append([NVCC] + FLAGS + defs + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, u[0]) for u in _units()]
    if jobs or _stale(LIB, objs):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
