"""ctypes binding of libmfgpu.so -- the C ABI declared in include/mfgpu.h.

There is no CPU fallback: if the CUDA library has not been built, importing
this module raises; if no CUDA device is usable, Context() raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmfgpu.so")

F32, F64 = 0, 1
SCATTER_ATOMIC, SCATTER_COLOR = 0, 1


class MfgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mfgpu error %d: %s" % (code, msg))
        self.code = code


class BoxDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("degree", C.c_int), ("log2_cells", C.c_int * 3), ("origin", C.c_double * 3),
                ("h", C.c_double), ("dirichlet_faces", C.c_uint32)]


class MfDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("degree", C.c_int), ("dtype", C.c_int), ("n_cells", C.c_uint32), ("n_dofs", C.c_uint32),
                ("loc2glob", C.POINTER(C.c_uint32)), ("geometry", C.c_int), ("inv_jac", C.POINTER(C.c_double)),
                ("JxW", C.POINTER(C.c_double)), ("quadrature_points", C.POINTER(C.c_double)), ("scatter", C.c_int),
                ("n_colors", C.c_uint32), ("color_offsets", C.POINTER(C.c_uint32)), ("constraint_mask", C.POINTER(C.c_uint32))]


class GpuData(C.Structure):
    """mfg_gpu_data: device arrays of a MatrixFreeGpu object for user-written kernels (MatrixFreeGpu::get_gpu_data)."""
    _fields_ = [("loc2glob", C.c_void_p), ("JxW", C.c_void_p), ("inv_jac", C.c_void_p), ("quadrature_points", C.c_void_p),
                ("constraint_mask", C.c_void_p), ("color_offsets", C.POINTER(C.c_uint32)),
                ("n_cells", C.c_uint32), ("n_dofs", C.c_uint32), ("n_colors", C.c_uint32), ("n_plain_cells", C.c_uint32),
                ("dim", C.c_int), ("degree", C.c_int), ("general", C.c_int), ("use_coloring", C.c_int), ("dtype", C.c_int),
                ("cuda_stream", C.c_void_p), ("shape_values", C.c_double * 81), ("shape_gradients", C.c_double * 81),
                ("colloc_gradients", C.c_double * 81)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libmfgpu.so is not built (%s). Run `python dealii_cuda_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32p, dp, sz = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.c_size_t
    pp = C.POINTER(C.c_void_p)
    sig = {
        "mfg_last_error": (C.c_char_p, []),
        "mfg_version": (C.c_char_p, []),
        "mfg_ctx_create": (C.c_int, [C.c_int, vp, pp]),
        "mfg_ctx_destroy": (C.c_int, [vp]),
        "mfg_ctx_set_stream": (C.c_int, [vp, vp]),
        "mfg_ctx_synchronize": (C.c_int, [vp]),
        "mfg_ctx_device_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(sz)]),
        "mfg_vec_create": (C.c_int, [vp, C.c_int, sz, pp]),
        "mfg_vec_wrap": (C.c_int, [vp, C.c_int, sz, vp, pp]),
        "mfg_vec_destroy": (C.c_int, [vp]),
        "mfg_vec_resize": (C.c_int, [vp, sz]),
        "mfg_vec_size": (sz, [vp]),
        "mfg_vec_dtype": (C.c_int, [vp]),
        "mfg_vec_data": (vp, [vp]),
        "mfg_vec_from_host": (C.c_int, [vp, vp, sz]),
        "mfg_vec_to_host": (C.c_int, [vp, vp, sz]),
        "mfg_vec_copy": (C.c_int, [vp, vp]),
        "mfg_vec_swap": (C.c_int, [vp, vp]),
        "mfg_vec_fill": (C.c_int, [vp, C.c_double]),
        "mfg_vec_sadd": (C.c_int, [vp, C.c_double, C.c_double, vp]),
        "mfg_vec_equ": (C.c_int, [vp, C.c_double, vp]),
        "mfg_vec_scale": (C.c_int, [vp, vp]),
        "mfg_vec_divide": (C.c_int, [vp, vp]),
        "mfg_vec_invert": (C.c_int, [vp]),
        "mfg_vec_scal": (C.c_int, [vp, C.c_double]),
        "mfg_vec_dot": (C.c_int, [vp, vp, dp]),
        "mfg_vec_add_and_dot": (C.c_int, [vp, C.c_double, vp, vp, dp]),
        "mfg_vec_l2_norm": (C.c_int, [vp, dp]),
        "mfg_vec_all_zero": (C.c_int, [vp, C.POINTER(C.c_int)]),
        "mfg_vec_copy_with_indices": (C.c_int, [vp, vp, vp, vp, sz]),
        "mfg_mesh_create_box": (C.c_int, [vp, C.POINTER(BoxDesc), pp]),
        "mfg_mesh_hyper_cube": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, pp]),
        "mfg_mesh_destroy": (C.c_int, [vp]),
        "mfg_mesh_n_cells": (C.c_uint32, [vp]),
        "mfg_mesh_n_dofs": (C.c_uint32, [vp]),
        "mfg_mesh_dofs_per_cell": (C.c_uint32, [vp]),
        "mfg_mesh_n_constrained": (C.c_uint32, [vp]),
        "mfg_mesh_loc2glob_device": (vp, [vp]),
        "mfg_mesh_constrained_device": (vp, [vp]),
        "mfg_mesh_get_loc2glob": (C.c_int, [vp, u32p]),
        "mfg_mesh_get_constrained": (C.c_int, [vp, u32p]),
        "mfg_mesh_get_support_points": (C.c_int, [vp, dp]),
        "mfg_mesh_get_cell_coords": (C.c_int, [vp, u32p]),
        "mfg_mesh_lattice_to_dof": (C.c_int, [vp, sz, u32p, u32p]),
        "mfg_mesh_color_cells": (C.c_int, [vp, u32p, u32p]),
        "mfg_mf_reinit": (C.c_int, [vp, C.POINTER(MfDesc), pp]),
        "mfg_mf_reinit_from_mesh": (C.c_int, [vp, vp, C.c_int, C.c_int, pp]),
        "mfg_mf_destroy": (C.c_int, [vp]),
        "mfg_mf_n_dofs": (C.c_uint32, [vp]),
        "mfg_mf_n_cells": (C.c_uint32, [vp]),
        "mfg_mf_n_colors": (C.c_uint32, [vp]),
        "mfg_mf_memory_consumption": (sz, [vp]),
        "mfg_mf_get_gpu_data": (C.c_int, [vp, C.POINTER(GpuData)]),
        "mfg_shape_info": (C.c_int, [C.c_int, dp, dp, dp, dp]),
        "mfg_hanging_node_weights": (C.c_int, [C.c_int, dp]),
        "mfg_ch_create": (C.c_int, [vp, C.c_int, u32p, sz, u32p, sz, pp]),
        "mfg_ch_create_from_mesh": (C.c_int, [vp, C.c_int, vp, pp]),
        "mfg_ch_destroy": (C.c_int, [vp]),
        "mfg_ch_n_constrained": (sz, [vp]),
        "mfg_ch_set_constrained_values": (C.c_int, [vp, vp, C.c_double]),
        "mfg_ch_save_constrained_values": (C.c_int, [vp, vp]),
        "mfg_ch_save_constrained_values2": (C.c_int, [vp, vp, vp]),
        "mfg_ch_load_constrained_values": (C.c_int, [vp, vp]),
        "mfg_ch_load_and_add_constrained_values": (C.c_int, [vp, vp, vp]),
        "mfg_ch_copy_edge_values": (C.c_int, [vp, vp, vp]),
        "mfg_laplace_create": (C.c_int, [vp, vp, C.c_int, C.c_int, pp]),
        "mfg_laplace_create_from_arrays": (C.c_int, [vp, vp, vp, dp, pp]),
        "mfg_laplace_set_coefficient": (C.c_int, [vp, dp]),
        "mfg_laplace_destroy": (C.c_int, [vp]),
        "mfg_laplace_m": (C.c_uint32, [vp]),
        "mfg_laplace_set_variant": (C.c_int, [vp, C.c_int]),
        "mfg_laplace_vmult": (C.c_int, [vp, vp, vp]),
        "mfg_laplace_vmult_add": (C.c_int, [vp, vp, vp]),
        "mfg_laplace_vmult_ptr": (C.c_int, [vp, vp, vp]),
        "mfg_laplace_vmult_add_ptr": (C.c_int, [vp, vp, vp]),
        "mfg_laplace_vmult_host": (C.c_int, [vp, vp, vp]),
        "mfg_laplace_vmult_host_async": (C.c_int, [vp, vp, vp, C.c_int]),
        "mfg_laplace_host_sync": (C.c_int, [vp]),
        "mfg_laplace_compute_diagonal": (C.c_int, [vp]),
        "mfg_laplace_get_diagonal_inverse": (C.c_int, [vp, pp]),
        "mfg_laplace_memory_consumption": (sz, [vp]),
        "mfg_laplace_launches_per_vmult": (C.c_int, [vp]),
        "mfg_laplace_cell_launches_per_vmult": (C.c_int, [vp]),
        "mfg_laplace_enable_kernel_timing": (C.c_int, [vp, C.c_int]),
        "mfg_laplace_kernel_time_ms": (C.c_int, [vp, dp, C.POINTER(C.c_int)]),
        "mfg_graph_coloring": (C.c_int, [C.c_uint32, C.c_uint32, u32p, C.c_uint32, u32p, u32p]),
        "mfg_laplace_active_variant": (C.c_int, [vp]),
        "mfg_laplace_set_option": (C.c_int, [vp, C.c_char_p, C.c_int]),
        "mfg_laplace_stage_stats": (C.c_int, [vp, u32p]),
        "mfg_chebyshev_create": (C.c_int, [vp, C.c_int, C.c_double, C.c_int, pp]),
        "mfg_chebyshev_destroy": (C.c_int, [vp]),
        "mfg_chebyshev_vmult": (C.c_int, [vp, vp, vp]),
        "mfg_chebyshev_step": (C.c_int, [vp, vp, vp]),
        "mfg_chebyshev_info": (C.c_int, [vp, dp, dp, dp, dp, C.POINTER(C.c_int)]),
        "mfg_vec_chebyshev_update": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_double, C.c_double, C.c_int, C.c_int]),
        "mfg_mg_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, pp]),
        "mfg_mg_destroy": (C.c_int, [vp]),
        "mfg_mg_vcycle": (C.c_int, [vp, vp, vp]),
        "mfg_mg_level_operator": (C.c_int, [vp, C.c_int, pp]),
        "mfg_mg_info": (C.c_int, [vp, C.c_int, dp, C.POINTER(C.c_long), C.POINTER(sz)]),
        "mfg_mg_solve_cg": (C.c_int, [vp, vp, vp, C.c_double, C.c_int, C.POINTER(C.c_int), dp, dp]),
        "mfg_cgd_init": (C.c_int, [vp, vp]),
        "mfg_cgd_dot": (C.c_int, [vp, C.c_int, vp, vp, vp, sz, vp]),
        "mfg_cgd_alpha": (C.c_int, [vp, vp]),
        "mfg_cgd_residual": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, sz, vp, C.c_int]),
        "mfg_cgd_beta": (C.c_int, [vp, vp, C.c_double, C.c_int]),
        "mfg_cgd_advance": (C.c_int, [vp, C.c_int, vp, vp, vp, sz, vp, C.c_int]),
        "mfg_stage_plan_build": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, u32p, C.c_int, pp]),
        "mfg_stage_plan_info": (C.c_int, [vp, u32p]),
        "mfg_stage_plan_get": (C.c_int, [vp, u32p, u32p, C.POINTER(C.c_uint16), u32p]),
        "mfg_stage_plan_destroy": (C.c_int, [vp]),
        "mfg_mgt_build": (C.c_int, [vp, vp, vp, C.c_int, pp]),
        "mfg_mgt_destroy": (C.c_int, [vp]),
        "mfg_mgt_prolongate": (C.c_int, [vp, vp, vp]),
        "mfg_mgt_restrict_and_add": (C.c_int, [vp, vp, vp]),
        "mfg_solver_cg": (C.c_int, [vp, vp, vp, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), dp, dp]),
        "mfg_exchange_create": (C.c_int, [vp, C.c_int, u32p, sz, u32p, sz, u32p, C.POINTER(C.c_int32), sz, pp]),
        "mfg_exchange_destroy": (C.c_int, [vp]),
        "mfg_exchange_pack": (C.c_int, [vp, vp, vp]),
        "mfg_exchange_accumulate": (C.c_int, [vp, vp, vp]),
        "mfg_exchange_accumulate_stream": (C.c_int, [vp, vp, vp, vp]),
        "mfg_laplace_set_interface_dofs": (C.c_int, [vp, C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_uint32)]),
        "mfg_laplace_vmult_part_ptr": (C.c_int, [vp, vp, vp, C.c_int, vp]),
        "mfg_exchange_pack_stream": (C.c_int, [vp, vp, vp, vp]),
        "mfg_exchange_push_stream": (C.c_int, [vp, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_int, vp]),
        "mfg_vec_dot_masked": (C.c_int, [vp, vp, vp, dp]),
        "mfg_laplace_bmop": (C.c_int, [vp, vp, vp, C.c_int, C.c_double, C.POINTER(C.c_float)]),
        "mfg_partition_rank_coords": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "mfg_partition_box": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(BoxDesc)]),
        "mfg_partition_global_n_dofs": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
        "mfg_partition_interface_points": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int,
                                                     C.POINTER(C.c_int), C.POINTER(sz), u32p]),
        "mfg_partition_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_int), C.POINTER(sz), u32p,
                                                C.c_int, C.POINTER(C.c_int), C.POINTER(sz), u32p, pp]),
        "mfg_partition_plan_destroy": (C.c_int, [vp]),
        "mfg_partition_plan_sizes": (C.c_int, [vp, C.POINTER(sz)]),
        "mfg_partition_plan_get": (C.c_int, [vp, C.POINTER(C.c_int), u32p, u32p, u32p, u32p, u32p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8)]),
        "mfg_csr_assemble_laplace": (C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, u32p, dp, dp, u32p, sz, pp]),
        "mfg_csr_destroy": (C.c_int, [vp]),
        "mfg_csr_sizes": (C.c_int, [vp, u32p, C.POINTER(sz)]),
        "mfg_csr_get": (C.c_int, [vp, u32p, u32p, dp]),
        "mfg_spm_create": (C.c_int, [vp, C.c_int, vp, pp]),
        "mfg_spm_create_from_mesh": (C.c_int, [vp, vp, C.c_int, pp]),
        "mfg_spm_destroy": (C.c_int, [vp]),
        "mfg_spm_m": (C.c_uint32, [vp]),
        "mfg_spm_n_nonzero_elements": (sz, [vp]),
        "mfg_spm_memory_consumption": (sz, [vp]),
        "mfg_spm_vmult": (C.c_int, [vp, vp, vp]),
        "mfg_umesh_hyper_ball": (C.c_int, [C.c_int, C.c_int, C.c_double, pp]),
        "mfg_umesh_destroy": (C.c_int, [vp]),
        "mfg_umesh_refine_global": (C.c_int, [vp, C.c_int]),
        "mfg_umesh_distribute_dofs": (C.c_int, [vp]),
        "mfg_umesh_n_cells": (C.c_uint32, [vp]),
        "mfg_umesh_n_vertices": (C.c_uint32, [vp]),
        "mfg_umesh_n_dofs": (C.c_uint32, [vp]),
        "mfg_umesh_n_boundary": (C.c_uint32, [vp]),
        "mfg_umesh_get_mesh": (C.c_int, [vp, dp, u32p]),
        "mfg_umesh_get_support_points": (C.c_int, [vp, dp]),
        "mfg_mf_reinit_from_umesh": (C.c_int, [vp, vp, C.c_int, pp]),
        "mfg_umesh_get_arrays": (C.c_int, [vp, u32p, u32p, dp, dp, dp, dp]),
        "mfg_laplace_create_from_umesh": (C.c_int, [vp, vp, C.c_int, pp]),
        "mfg_amesh_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "mfg_mgt_build_from_blocks": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_uint32, u32p, u32p, dp, C.c_uint32, C.c_uint32, pp]),
        "mfg_amg_create": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, pp]),
        "mfg_amg_destroy": (C.c_int, [vp]),
        "mfg_amg_vcycle": (C.c_int, [vp, vp, vp]),
        "mfg_amg_active_operator": (C.c_int, [vp, pp]),
        "mfg_amg_level_operator": (C.c_int, [vp, C.c_int, pp]),
        "mfg_amg_vmult_interface_down": (C.c_int, [vp, C.c_int, vp, vp]),
        "mfg_amg_vmult_interface_up": (C.c_int, [vp, C.c_int, vp, vp]),
        "mfg_amg_prolongate": (C.c_int, [vp, C.c_int, vp, vp]),
        "mfg_amg_restrict_and_add": (C.c_int, [vp, C.c_int, vp, vp]),
        "mfg_amg_copy_to_level": (C.c_int, [vp, C.c_int, vp, vp]),
        "mfg_amg_copy_from_level": (C.c_int, [vp, C.c_int, vp, vp]),
        "mfg_amg_info": (C.c_int, [vp, C.c_int, dp, C.POINTER(C.c_long), C.POINTER(sz), C.POINTER(sz)]),
        "mfg_amg_solve_cg": (C.c_int, [vp, vp, vp, C.c_double, C.c_int, C.POINTER(C.c_int), dp, dp]),
        "mfg_amesh_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, pp]),
        "mfg_amesh_destroy": (C.c_int, [vp]),
        "mfg_amesh_set_limit_level_difference_at_vertices": (C.c_int, [vp, C.c_int]),
        "mfg_amesh_refine_global": (C.c_int, [vp, C.c_int]),
        "mfg_amesh_set_refine_flags": (C.c_int, [vp, C.POINTER(C.c_uint8), sz]),
        "mfg_amesh_mark_cells_in_annulus": (C.c_int, [vp, C.c_double, C.c_double, dp]),
        "mfg_amesh_mark_cells_on_shell": (C.c_int, [vp, C.c_double, dp]),
        "mfg_amesh_mark_octant": (C.c_int, [vp]),
        "mfg_amesh_execute_refinement": (C.c_int, [vp]),
        "mfg_amesh_pseudo_adaptive_refinement": (C.c_int, [vp, C.c_int]),
        "mfg_amesh_n_active_cells": (C.c_uint32, [vp]),
        "mfg_amesh_n_levels": (C.c_uint32, [vp]),
        "mfg_amesh_get_active_cells": (C.c_int, [vp, u32p]),
        "mfg_amesh_n_level_cells": (C.c_uint32, [vp, C.c_int]),
        "mfg_amesh_get_level_cells": (C.c_int, [vp, C.c_int, u32p]),
        "mfg_amesh_distribute_dofs": (C.c_int, [vp]),
        "mfg_amesh_n_boundary": (C.c_uint32, [vp]),
        "mfg_amesh_get_boundary": (C.c_int, [vp, u32p]),
        "mfg_amesh_get_support_points": (C.c_int, [vp, dp]),
        "mfg_mf_reinit_from_amesh": (C.c_int, [vp, vp, C.c_int, pp]),
        "mfg_amesh_build_mg": (C.c_int, [vp, C.c_int]),
        "mfg_amesh_mg_level_sizes": (C.c_int, [vp, C.c_int, u32p]),
        "mfg_amesh_mg_level_get": (C.c_int, [vp, C.c_int, u32p, u32p, u32p, dp, u32p, u32p, u32p, u32p, dp]),
        "mfg_amesh_n_dofs": (C.c_uint32, [vp]),
        "mfg_amesh_n_constrained": (C.c_uint32, [vp]),
        "mfg_amesh_n_hanging": (C.c_uint32, [vp]),
        "mfg_amesh_get_arrays": (C.c_int, [vp, u32p, u32p, u32p, u32p, u32p, dp, dp, dp]),
        "mfg_laplace_create_from_amesh": (C.c_int, [vp, vp, C.c_int, pp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError if a declared symbol is missing
        f.restype = res
        f.argtypes = args
    return L, sorted(sig)


lib, DECLARED_SYMBOLS = _load()


def check(rc):
    if rc != 0:
        raise MfgError(rc, lib.mfg_last_error().decode(errors="replace"))
