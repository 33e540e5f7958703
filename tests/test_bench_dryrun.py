"""bench.py and the torchrun workers of the multi-GPU tests, END TO END ON THE CPU: the torch side runs for real (CPU tensors, gloo
process group), torch.cuda is shimmed (tests/emu/torch_cpu_shim.py), dealii_cuda_b200 is bound to libmfgpu_emu.so -- the library's
own sources compiled for the CPU, whose "device" memory is host memory, so a CPU tensor's data_ptr() is a valid device pointer.
Every Python statement of the N = 1 line, of distributed.bench_main (weak and strong, 2 ranks: exchange plan, NCCL-path exchange,
distributed CG, multigrid over the partition with DistributedLevel, the watchdog scaffold, the JSON line) and of the multigrid
worker executes before its first run on hardware; CUDA graphs and symmetric memory do not exist here, so the eager / all_to_all
fallbacks are the paths taken.  A pre-flight against NameErrors and format strings in code the driver runs once -- not a measurement."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "emu", "run_with_cpu_shim.py")
_port = [29700 + os.getpid() % 200]


def run_ranks(emu, world, script, args, timeout=600):
    """one process per rank (RANK / WORLD_SIZE / MASTER_* as torchrun sets them; LOCAL_RANK = 0: the emulation has one device)"""
    _port[0] += 1
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port[0]),
                   PYTHONPATH=os.pathsep.join([emu["pkg"], ROOT]), MFG_EMULATION="1", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, RUNNER, script] + [str(a) for a in args], cwd=emu["pkg"], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=timeout))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    return [p.returncode for p in procs], [o[0] for o in outs], [o[1] for o in outs]


def the_line(stdout):
    lines = [l for l in stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, stdout[-2000:]
    return json.loads(lines[0])


def test_bench_single_gpu_line_on_the_emulation(emu):
    rc, out, err = run_ranks(emu, 1, os.path.join(ROOT, "bench.py"), ["--gpus", 1, "--steps", 3, "--warmup", 3, "--refine", 2, "--degree", 2,
                                                                    "--e2e-steps", 2, "--cpu-steps", 2])
    assert rc == [0], err[0][-3000:]
    d = the_line(out[0])
    assert d["n_gpus"] == 1 and d["unit"] == "DoFs/s" and d["value"] > 0 and d["roofline"]["frac"] > 0 and d["cpu_baseline"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 729 * 8 and d["e2e"]["value"] > 0
    assert d["cg_solve"]["iterations"] > 0 and d["cg_solve"]["rel_error"] < 1e-9
    assert "error" not in d["mg_solve"] and d["mg_solve"]["iterations"] <= 8 and d["mg_solve"]["rel_error"] < 1e-8


@pytest.mark.parametrize("scaling", ["weak", "strong"])
def test_bench_two_rank_line_on_the_emulation(emu, scaling):
    rc, out, err = run_ranks(emu, 2, os.path.join(ROOT, "bench.py"), ["--gpus", 2, "--steps", 3, "--warmup", 3, "--refine", 2, "--degree", 1,
                                                                    "--e2e-steps", 2, "--scaling", scaling])
    assert rc == [0, 0], err[0][-3000:] + err[1][-3000:]
    assert not [l for l in out[1].splitlines() if l.startswith("{")], "only rank 0 prints the line"
    d = the_line(out[0])
    assert d["n_gpus"] == 2 and d["scaling"] == scaling and d["value"] > 0 and d["roofline"]["frac"] > 0 and "note" not in d
    assert d["e2e"]["value"] > 0 and d["e2e"]["steps"] >= 6
    assert d["cg_solve"]["iterations"] > 0 and d["cg_solve"]["rel_error"] < 1e-9
    assert "error" not in d["mg_solve"] and d["mg_solve"]["iterations"] <= 8 and d["mg_solve"]["rel_error"] < 1e-8, d["mg_solve"]
    # the overlapped apply: slab3 kernel with the interface groups first, its self-check against the plain sequence
    assert "variant 50" in d["roofline"]["kernel"] and d["overlap"].startswith("interface cell groups first")
    assert d["selfcheck"]["overlapped_vs_sequential_max_rel_diff"] < 1e-12


def test_bench_watchdog_prints_the_apply_line(emu):
    """a section behind the timed region that does not finish in time: rank 0 still prints the measured apply line, all ranks leave"""
    os.environ["MFG_BENCH_WATCHDOG_S"] = "1"
    try:
        rc, out, err = run_ranks(emu, 2, os.path.join(ROOT, "bench.py"), ["--gpus", 2, "--steps", 3, "--warmup", 3, "--refine", 2, "--degree", 1, "--e2e-steps", 2])
    finally:
        del os.environ["MFG_BENCH_WATCHDOG_S"]
    assert rc == [0, 0], err[0][-3000:] + err[1][-3000:]
    d = the_line(out[0])
    assert d["value"] > 0 and d["roofline"]["frac"] > 0 and "did not finish within 1 s" in d["note"]


def test_bench_single_gpu_watchdog_prints_the_apply_line(emu):
    """N = 1: the CG / multigrid solves behind the timed region run under a watchdog too; the apply line (with roofline, e2e and the CPU
    baseline, all measured before them) is printed once when they do not finish in time"""
    os.environ["MFG_BENCH_WATCHDOG_S"] = "0.001"
    try:
        rc, out, err = run_ranks(emu, 1, os.path.join(ROOT, "bench.py"), ["--gpus", 1, "--steps", 3, "--warmup", 3, "--refine", 2, "--degree", 2,
                                                                        "--e2e-steps", 2, "--cpu-steps", 2])
    finally:
        del os.environ["MFG_BENCH_WATCHDOG_S"]
    assert rc == [0], err[0][-3000:]
    d = the_line(out[0])
    assert d["value"] > 0 and d["roofline"]["frac"] > 0 and d["e2e"]["value"] > 0 and d["cpu_baseline"]["value"] > 0
    assert "did not finish within 0.001 s" in d["note"]


def test_multigrid_worker_on_the_emulation(emu):
    """tests/multirank_mg_worker.py (the torchrun worker of the late GPU test): multigrid over the partition with one box per rank"""
    rc, out, err = run_ranks(emu, 2, os.path.join(ROOT, "tests", "multirank_mg_worker.py"), [1, 2, "weak"])
    assert rc == [0, 0], err[0][-3000:] + err[1][-3000:]
    assert sum(o.count("MULTIRANK_MG_OK") for o in out) == 2, out


@pytest.mark.parametrize("mode", [["--adaptive", "--dim", 2, "--refine", 2, "--degree", 2], ["--spmv", "--refine", 2, "--degree", 2]])
def test_bench_extra_modes_on_the_emulation(emu, mode):
    """bench.py --adaptive (BASELINE configs[3]: pseudo-adaptive mesh with hanging nodes, apply + Jacobi-CG + local-smoothing MG-CG) and
    --spmv (the assembled-matrix competitor row): never run on hardware, every statement runs here"""
    rc, out, err = run_ranks(emu, 1, os.path.join(ROOT, "bench.py"), mode + ["--steps", 2, "--warmup", 3])
    assert rc == [0], err[0][-3000:]
    d = the_line(out[0])
    assert d["value"] > 0 and d["roofline"]["frac"] > 0
    if "--adaptive" in mode:
        assert "hanging-node constraints" in d["config"]["workload"] and d["cg_solve"]["iterations"] > 0
        assert "error" not in d["mg_solve"] and d["mg_solve"]["iterations"] <= 12 and d["mg_solve"]["rel_error"] < 1e-8
    else:
        assert d["matrix_free"]["value"] > 0 and d["matrix_bytes"] > 0


def test_multirank_apply_worker_on_the_emulation(emu):
    """tests/multirank_worker.py (the torchrun worker of tests/test_gpu_multirank.py, hardware-verified at N = 2, 4, 8): the 2-rank apply,
    global dot product and distributed CG against the GLOBAL oracle mesh, FP64 and FP32, strong partition -- here over gloo"""
    rc, out, err = run_ranks(emu, 2, os.path.join(ROOT, "tests", "multirank_worker.py"), [2, 2, "strong"])
    assert rc == [0, 0], err[0][-3000:] + err[1][-3000:]
    assert sum(o.count("MULTIRANK_OK") for o in out) == 2, out
