"""TEST INFRASTRUCTURE: builds libmfgpu_emu.so -- the library's own sources compiled for the CPU.

The .cu / .cuh files of dealii_cuda_b200/csrc are copied into a scratch directory with three textual changes -- kernel launches
`k<<<grid, block, smem, stream>>>(args)` become `emu_launch4(grid, block, smem, stream, k, args)`, shared-memory declarations
become static / a pointer into one buffer, inline PTX disappears -- and compiled by g++ against tests/emu/cuda_emu_runtime.h
("device" memory = host memory, the CUDA threads of a block are fibers switched at barriers; tests/emu/cub/cub.cuh stands in for the
CUB scans of mesh.cu).  The PTX helpers of the slab3 and staged kernels (bulk-async copy + mbarrier, cp.async) get host bodies that copy at issue:
every unit of the library is built, nothing is stubbed.  The result exports the same C ABI: the Python binding loads it when MFG_EMULATED_LIB names it, and GPU tests of
code paths that need no fast kernel can run on the CPU (tests/test_emulated_library.py)."""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "dealii_cuda_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")

UNITS = [("capi.cu", "", []), ("mesh.cu", "", []), ("vector.cu", "", []), ("operators.cu", "", []), ("solver.cu", "", []), ("multigrid.cu", "", []), ("mg_transfer.cu", "", []),
         ("exchange.cu", "", []), ("stage_plan.cu", "", []), ("coloring.cu", "", []), ("adaptive_mesh.cu", "", []), ("partition.cu", "", []),
         ("sparse_matrix.cu", "", []), ("ball_mesh.cu", "", [])]
for dim in (2, 3):
    for f64 in (0, 1):
        UNITS.append(("kernels_v0_inst.cu", "_d%d_f%d" % (dim, f64), ["-DMFG_INST_DIM=%d" % dim, "-DMFG_INST_F64=%d" % f64]))
        UNITS.append(("kernels_general_inst.cu", "_d%d_f%d" % (dim, f64), ["-DMFG_INST_DIM=%d" % dim, "-DMFG_INST_F64=%d" % f64]))
for f64 in (0, 1):
    UNITS.append(("kernels_slab3_inst.cu", "_f%d" % f64, ["-DMFG_INST_F64=%d" % f64]))
    UNITS.append(("kernels_stage_inst.cu", "_f%d" % f64, ["-DMFG_INST_F64=%d" % f64]))


def _match_paren(s, i):
    """index just behind the parenthesis that closes the one opening at s[i]"""
    depth = 0
    while True:
        c = s[i]
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def rewrite_launches(s):
    out, pos = [], 0
    while True:
        k = s.find("<<<", pos)
        if k < 0:
            out.append(s[pos:])
            return "".join(out)
        # the kernel expression in front of <<<: identifier, optional template arguments, or a parenthesised expression
        j = k
        while j > 0 and s[j - 1] in " \t":
            j -= 1
        if s[j - 1] == ">":                                   # template arguments
            depth, j = 0, j - 1
            while True:
                if s[j] == ">":
                    depth += 1
                elif s[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while j > 0 and (s[j - 1].isalnum() or s[j - 1] in "_:"):
            j -= 1
        kernel = s[j:k].strip()
        e = s.index(">>>", k)
        cfg = s[k + 3:e]
        a = e + 3
        while s[a] in " \t\n":
            a += 1
        assert s[a] == "(", (kernel, s[k:k + 80])
        b = _match_paren(s, a)
        args = s[a + 1:b - 1].strip()
        parts = [p.strip() for p in _split_top(cfg)]
        while len(parts) < 4:
            parts.append("0" if len(parts) == 2 else "nullptr")
        out.append(s[pos:j])
        # (the kernel goes through a generic lambda: its template arguments may be deduced from the call, as nvcc does at a launch)
        out.append("emu_launch4(%s, %s, %s, %s, [&](auto... emu_a) { %s(emu_a...); }%s)"
                   % (parts[0], parts[1], parts[2], parts[3], kernel, (", " + args) if args else ""))
        pos = b


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur); cur = ""
        else:
            cur += c
    parts.append(cur)
    return parts


def remove_asm(s):
    out, pos = [], 0
    while True:
        k = s.find("asm volatile", pos)
        if k < 0:
            out.append(s[pos:])
            return "".join(out)
        a = s.index("(", k)
        b = _match_paren(s, a)
        while s[b] in " \t":
            b += 1
        assert s[b] == ";", s[k:b + 20]
        out.append(s[pos:k])
        out.append("/* inline PTX removed for the CPU emulation */")
        pos = b      # keep the ';'


# what the PTX of these helpers does, as host code, inserted at the top of their bodies (the asm itself is removed): the bulk-async
# copy and cp.async complete at issue, so the waits (mbarrier, cp.async.wait_group) have nothing left to wait for
PTX_HELPERS = {
    "__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)\n{":
        "\n  std::memcpy(smem_dst, gsrc, bytes);",
    "template <int BYTES> __device__ __forceinline__ void slab3_cp_zfill(void *smem_dst, const void *gsrc, bool valid)\n{":
        "\n  if (valid) std::memcpy(smem_dst, gsrc, BYTES); else std::memset(smem_dst, 0, BYTES);",
    "template <int BYTES> __device__ __forceinline__ void cp_async_elem(void *smem_dst, const void *gsrc)\n{":
        "\n  std::memcpy(smem_dst, gsrc, BYTES);",
}


def transform(text):
    text = text.replace("#include <cuda_runtime.h>", '#include "cuda_emu_runtime.h"')
    for head, body in PTX_HELPERS.items():
        if head in text:
            text = text.replace(head, head + body)
    text = remove_asm(text)
    text = rewrite_launches(text)
    # extern __shared__ [__align__(16)] T name[];  ->  T *name = reinterpret_cast<T *>(emu_dyn_smem);
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];", r"\1 *\2 = reinterpret_cast<\1 *>(emu_dyn_smem);", text)
    text = re.sub(r"\b__shared__\s+", "static ", text)        # one block is alive at a time: a static is shared by its threads
    return text


def build(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    mirror = os.path.join(out_dir, "dealii_cuda_b200", "csrc")
    os.makedirs(mirror, exist_ok=True)
    os.makedirs(os.path.join(out_dir, "include"), exist_ok=True)
    with open(os.path.join(out_dir, "include", "mfgpu.h"), "w") as f:
        f.write(open(os.path.join(ROOT, "include", "mfgpu.h")).read())
    for name in os.listdir(CSRC):
        if name.endswith((".cu", ".cuh", ".h")):
            text = open(os.path.join(CSRC, name)).read()
            # the slab3 / staged kernels themselves are not compiled (TMA, mbarrier, cp.async): only what their callers see
            with open(os.path.join(mirror, name), "w") as f:
                f.write(transform(text))
    objs = []

    def compile_unit(u):
        src, suffix, defs = u
        obj = os.path.join(out_dir, src.replace(".cu", suffix + ".o"))
        cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-pthread", "-w", "-x", "c++", "-DMFG_EMULATION", "-I", EMU, "-I", mirror] + defs + \
              ["-c", os.path.join(mirror, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed for %s:\n%s" % (src, r.stderr[-6000:]))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    stub = os.path.join(out_dir, "emu_stubs.o")   # (nothing is stubbed any more: the unit stays as the place for it)
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-pthread", "-w", "-DMFG_EMULATION", "-I", EMU, "-I", mirror, "-c",
                        os.path.join(EMU, "emu_stubs.cc"), "-o", stub], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed for emu_stubs.cc:\n%s" % r.stderr[-6000:])
    so = os.path.join(out_dir, "libmfgpu_emu.so")
    r = subprocess.run(["g++", "-shared", "-pthread", "-o", so] + objs + [stub], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation link failed:\n%s" % r.stderr[-6000:])
    return so


def build_examples(out_dir, so):
    """examples/ against libmfgpu_emu.so: the host-only drivers as they are, generic_ops.cu / poisson.cu (user-written functors on the
    header-only FEEvaluationGpu path) through the same textual transformation.  Returns the directory that stands in for examples/_build."""
    inc = os.path.join(out_dir, "include", "dealii_cuda_b200")
    ex = os.path.join(out_dir, "examples")
    bld = os.path.join(ex, "_build")
    for d in (inc, bld):
        os.makedirs(d, exist_ok=True)
    src_inc = os.path.join(ROOT, "include", "dealii_cuda_b200")
    for name in os.listdir(src_inc):
        with open(os.path.join(inc, name), "w") as f:
            f.write(transform(open(os.path.join(src_inc, name)).read()))
    for name in os.listdir(os.path.join(ROOT, "examples")):
        if name.endswith((".cu", ".cc", ".h")):
            with open(os.path.join(ex, name), "w") as f:
                f.write(transform(open(os.path.join(ROOT, "examples", name)).read()))
    link = ["-L", os.path.dirname(so), "-lmfgpu_emu", "-Wl,-rpath," + os.path.dirname(so), "-pthread"]
    base = ["g++", "-std=c++20", "-O1", "-w", "-DMFG_EMULATION", "-I", EMU]
    jobs = [(["-x", "c++", os.path.join(ex, "generic_ops.cu"), "-shared", "-fPIC", "-o", os.path.join(bld, "libgeneric_ops.so")]),
            (["-x", "c++", os.path.join(ex, "poisson.cu"), "-o", os.path.join(bld, "poisson")]),
            ([os.path.join(ex, "bmop.cc"), "-o", os.path.join(bld, "bmop")]),
            (["-DADAPTIVE_GRID", os.path.join(ex, "bmop.cc"), "-o", os.path.join(bld, "bmop_adaptive")]),
            (["-DBALL_GRID", os.path.join(ex, "bmop.cc"), "-o", os.path.join(bld, "bmop_ball")]),
            # (degree 2 builds: the Q4 drivers take minutes on the emulation at any useful size)
            (["-DADAPTIVE_GRID", "-DDEGREE_FE=2", os.path.join(ex, "bmop.cc"), "-o", os.path.join(bld, "bmop_adaptive_q2")]),
            (["-DBALL_GRID", "-DDEGREE_FE=2", os.path.join(ex, "bmop.cc"), "-o", os.path.join(bld, "bmop_ball_q2")]),
            (["-DADAPTIVE_GRID", "-DDIMENSION=2", "-DDEGREE_FE=2", os.path.join(ex, "bmop.cc"), "-o", os.path.join(bld, "bmop_adaptive_2d_q2")]),
            (["-DDIMENSION=2", "-DDEGREE_FE=2", os.path.join(ex, "partitioned_mg.cc"), "-o", os.path.join(bld, "partitioned_mg_2d_q2")]),
            (["-DDEGREE_FE=2", os.path.join(ex, "partitioned_mg.cc"), "-o", os.path.join(bld, "partitioned_mg_q2")])]

    def run(j):
        # (-x c++ applies to the files behind it: the link options go last, behind `-x none`)
        r = subprocess.run(base + j + ["-x", "none"] + link, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed for %s:\n%s" % (j, r.stderr[-6000:]))

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as exr:
        list(exr.map(run, jobs))
    return bld


def build_package(out_dir, so):
    """a copy of the Python binding next to the emulated library (as lib/libmfgpu.so): put `<result>` in front of the repository on
    PYTHONPATH and `import dealii_cuda_b200` binds libmfgpu_emu.so -- the product package itself has no emulation switch"""
    import shutil
    pkg = os.path.join(out_dir, "pkg", "dealii_cuda_b200")
    os.makedirs(os.path.join(pkg, "lib"), exist_ok=True)
    for name in os.listdir(os.path.join(ROOT, "dealii_cuda_b200")):
        if name.endswith(".py"):
            shutil.copy(os.path.join(ROOT, "dealii_cuda_b200", name), os.path.join(pkg, name))
    dst = os.path.join(pkg, "lib", "libmfgpu.so")
    if os.path.lexists(dst):
        os.remove(dst)
    os.symlink(so, dst)
    return os.path.join(out_dir, "pkg")


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/mfg_emu"
    so = build(out)
    print(so, build_examples(out, so), build_package(out, so))
