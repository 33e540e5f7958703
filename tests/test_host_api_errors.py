"""Error behaviour of the library's host-side entry points (no device needed): bad arguments come back as a negative status with a
message in mfg_last_error (the reference asserts / throws dealii::ExcMessage), never as a crash."""
import ctypes as C

import numpy as np
import pytest

import dealii_cuda_b200 as mf
from dealii_cuda_b200 import _capi, partition
from dealii_cuda_b200._capi import lib


def test_adaptive_mesh_call_order_and_arguments():
    with pytest.raises(mf.MfgError):
        mf.AdaptiveMesh(4, 2)                                  # dim must be 2 or 3
    with pytest.raises(mf.MfgError):
        mf.AdaptiveMesh(3, 9)                                  # degree 1..8
    am = mf.AdaptiveMesh(2, 2).refine_global(1)
    with pytest.raises(mf.MfgError, match="distribute_dofs"):
        am.arrays()                                            # DoFs are not distributed yet
    with pytest.raises(mf.MfgError, match="distribute_dofs"):
        am.build_mg(0)
    with pytest.raises(mf.MfgError, match="one flag per active cell"):
        am.set_refine_flags([1, 0])
    am.distribute_dofs()
    with pytest.raises(mf.MfgError, match="min_level"):
        am.build_mg(2)                                         # above the coarsest active level
    am.build_mg(0)
    with pytest.raises(mf.MfgError, match="bad level"):
        am.mg_level(7)
    am.refine_global(1)                                        # refining invalidates DoFs and hierarchy
    with pytest.raises(mf.MfgError):
        am.arrays()
    with pytest.raises(mf.MfgError, match="unrefined"):
        am.pseudo_adaptive_refinement(4)                       # starts from the single cell only


def test_partition_arguments():
    with pytest.raises(mf.MfgError):
        partition.rank_coords(0, 3, 3)                         # 3 ranks: no grid
    with pytest.raises(mf.MfgError):
        partition.rank_coords(5, 4, 3)                         # rank out of range
    with pytest.raises(mf.MfgError, match="more ranks than cells"):
        partition.box_for_rank(0, 8, 3, 0, strong=True)        # a single cell cannot be cut into 2 x 2 x 2
    with pytest.raises(mf.MfgError, match="out of range"):
        partition.ExchangePlan(0, 2, {1: np.array([5], np.uint32)}, 3)      # shared DoF beyond n_local
    with pytest.raises(mf.MfgError):
        partition.ExchangePlan(0, 2, {0: np.array([1], np.uint32)}, 3)      # a rank does not exchange with itself


def test_ball_mesh_and_sparse_matrix_arguments():
    with pytest.raises(mf.MfgError):
        mf.BallMesh(1, 2)
    with pytest.raises(mf.MfgError, match="radius"):
        mf.BallMesh(2, 2, radius=-1.0)
    bm = mf.BallMesh(2, 2)
    with pytest.raises(mf.MfgError, match="distribute_dofs"):
        lib_check = _capi.check
        out = np.zeros(4)
        lib_check(lib.mfg_umesh_get_support_points(bm.h, out.ctypes.data_as(C.POINTER(C.c_double))))
    l2g = np.array([[0, 1, 2, 9]], np.uint32)                  # an index beyond n_dofs
    with pytest.raises(mf.MfgError, match="out of range"):
        mf.assemble_laplace_csr(2, 1, l2g, 4, np.ones(1), np.ones((1, 4)), np.zeros(0, np.uint32))
    with pytest.raises(mf.MfgError, match="out of range"):
        mf.assemble_laplace_csr(2, 1, np.array([[0, 1, 2, 3]], np.uint32), 4, np.ones(1), np.ones((1, 4)), np.array([7], np.uint32))


def test_graph_coloring_and_stage_plan_reject_bad_input():
    with pytest.raises(mf.MfgError):
        mf.graph_coloring(np.array([[0, 1, 2, 99]], np.uint32), 4)   # conflict index beyond n_indices


def test_numa_binding_helper_is_harmless_without_topology():
    """bench's NUMA placement of pinned buffers: the cpulist parser, and no change of the affinity where the GPU / sysfs topology is
    not visible (this container)"""
    import os
    from dealii_cuda_b200.distributed import bind_to_gpu_numa_node, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and parse_cpulist("5") == {5} and parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
