"""Builds libmfgpu.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python dealii_cuda_b200/build.py [--force] [-v]   (run as a script: importing the package needs the built library)
The .so lands in dealii_cuda_b200/lib/ (git-ignored, travels to the GPU box).
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
OBJ = os.path.join(OUT, "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]

# (source, object suffix, extra defines)
UNITS = [("capi.cu", "", []), ("mesh.cu", "", []), ("vector.cu", "", []), ("operators.cu", "", []), ("exchange.cu", "", []), ("solver.cu", "", []), ("mg_transfer.cu", "", []), ("stage_plan.cu", "", []), ("multigrid.cu", "", []), ("coloring.cu", "", []), ("adaptive_mesh.cu", "", []), ("partition.cu", "", []), ("sparse_matrix.cu", "", []), ("ball_mesh.cu", "", [])]
for dim in (2, 3):
    for f64 in (0, 1):
        UNITS.append(("kernels_v0_inst.cu", f"_d{dim}_f{f64}", [f"-DMFG_INST_DIM={dim}", f"-DMFG_INST_F64={f64}"]))
        UNITS.append(("kernels_general_inst.cu", f"_d{dim}_f{f64}", [f"-DMFG_INST_DIM={dim}", f"-DMFG_INST_F64={f64}"]))
for f64 in (0, 1):
    UNITS.append(("kernels_stage_inst.cu", f"_f{f64}", [f"-DMFG_INST_F64={f64}"]))
    UNITS.append(("kernels_slab3_inst.cu", f"_f{f64}", [f"-DMFG_INST_F64={f64}"]))


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "include", "mfgpu.h")]


def _stamp(src, defs):
    h = hashlib.sha1()
    for f in [src] + _headers():
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(defs + FLAGS + ARCH).encode())
    return h.hexdigest()


def _compile(unit, force):
    src, suffix, defs = unit
    srcp = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src.replace(".cu", suffix + ".o"))
    stampf = obj + ".stamp"
    stamp = _stamp(srcp, defs)
    if not force and os.path.exists(obj) and os.path.exists(stampf) and open(stampf).read() == stamp:
        return obj, False, ""
    cmd = [NVCC] + ARCH + FLAGS + defs + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stampf, "w") as fh:
        fh.write(stamp)
    with open(obj + ".ptxas.log", "w") as fh:
        fh.write(r.stderr)
    return obj, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        res = list(ex.map(lambda u: _compile(u, force), UNITS))
    objs = [r[0] for r in res]
    so = os.path.join(OUT, "libmfgpu.so")
    if any(r[1] for r in res) or not os.path.exists(so):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", so] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        for r in res:
            if r[2]:
                print(r[2])
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
