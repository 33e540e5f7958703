// capi.cu -- extern "C" entry points of libmfgpu.so (include/mfgpu.h).
#include <cstdlib>
#include "operators.cuh"
#include "kernels_stage.cuh"

namespace mfg {
static thread_local std::string g_last_error;
void set_last_error(const std::string &msg) { g_last_error = msg; }
}  // namespace mfg

using namespace mfg;

extern "C" {

const char *mfg_last_error(void) { return g_last_error.c_str(); }
const char *mfg_version(void) { return "mfgpu 0.1 (sm_100a)"; }

// ---- context ----------------------------------------------------------------
int mfg_ctx_create(int device, void *stream, mfg_ctx **out)
{
  return guarded([&] {
    MFG_REQUIRE(out != nullptr, "out is null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(MFG_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    MFG_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    MFG_CUDA(cudaSetDevice(device));
    std::unique_ptr<mfg_ctx> ctx(new mfg_ctx);
    ctx->device = device; ctx->stream = (cudaStream_t)stream;
    cudaDeviceProp prop;
    MFG_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount; ctx->cc_major = prop.major; ctx->cc_minor = prop.minor; ctx->l2_bytes = (size_t)prop.l2CacheSize;
    MFG_CUDA(cudaMalloc(&ctx->red_dev, RED_SCRATCH_DOUBLES * sizeof(double)));
    MFG_CUDA(cudaMallocHost(&ctx->red_host, RED_HOST_DOUBLES * sizeof(double)));
    *out = ctx.release();
  });
}
int mfg_ctx_destroy(mfg_ctx *ctx)
{
  return guarded([&] {
    if (!ctx) return;
    if (ctx->red_dev) cudaFree(ctx->red_dev);
    if (ctx->red_host) cudaFreeHost(ctx->red_host);
    delete ctx;
  });
}
int mfg_ctx_set_stream(mfg_ctx *ctx, void *stream) { return guarded([&] { MFG_REQUIRE(ctx, "ctx is null"); ctx->stream = (cudaStream_t)stream; }); }
int mfg_ctx_synchronize(mfg_ctx *ctx) { return guarded([&] { MFG_REQUIRE(ctx, "ctx is null"); MFG_CUDA(cudaStreamSynchronize(ctx->stream)); }); }
int mfg_ctx_device_info(mfg_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes)
{
  return guarded([&] {
    MFG_REQUIRE(ctx, "ctx is null");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
  });
}

// ---- GpuVector ----------------------------------------------------------------
int mfg_vec_create(mfg_ctx *ctx, mfg_dtype dt, size_t n, mfg_vec **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && out, "null argument");
    std::unique_ptr<mfg_vec> v(new mfg_vec);
    v->ctx = ctx; v->dt = dt; v->n = v->cap = n; v->owns = true;
    if (n)
      {
        cudaError_t e = cudaMalloc(&v->p, n * v->esize());
        if (e != cudaSuccess) throw Error(MFG_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        MFG_CUDA(cudaMemsetAsync(v->p, 0, n * v->esize(), ctx->stream));
      }
    *out = v.release();
  });
}
int mfg_vec_wrap(mfg_ctx *ctx, mfg_dtype dt, size_t n, void *device_ptr, mfg_vec **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && out && (device_ptr || n == 0), "null argument");
    mfg_vec *v = new mfg_vec;
    v->ctx = ctx; v->dt = dt; v->n = v->cap = n; v->p = device_ptr; v->owns = false;
    *out = v;
  });
}
int mfg_vec_destroy(mfg_vec *v)
{
  return guarded([&] { if (!v) return; if (v->owns && v->p) cudaFree(v->p); delete v; });
}
int mfg_vec_resize(mfg_vec *v, size_t n)
{
  return guarded([&] {
    MFG_REQUIRE(v, "null vector");
    if (n == v->n) return;
    MFG_REQUIRE(v->owns, "cannot resize a wrapped vector");
    if (n > v->cap)
      {
        if (v->p) MFG_CUDA(cudaFree(v->p));
        v->p = nullptr; v->cap = 0;
        cudaError_t e = cudaMalloc(&v->p, n * v->esize());
        if (e != cudaSuccess) throw Error(MFG_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        v->cap = n;
      }
    v->n = n;
  });
}
size_t mfg_vec_size(const mfg_vec *v) { return v ? v->n : 0; }
mfg_dtype mfg_vec_dtype(const mfg_vec *v) { return v ? v->dt : MFG_F64; }
void *mfg_vec_data(mfg_vec *v) { return v ? v->p : nullptr; }
int mfg_vec_from_host(mfg_vec *v, const void *host, size_t n)
{
  return guarded([&] {
    MFG_REQUIRE(v && host, "null argument");
    // fromHost prints and returns on a size mismatch (gpu_vec.cu:148-159); here it is an error
    MFG_REQUIRE(n == v->n, "fromHost: size mismatch");
    if (n) { MFG_CUDA(cudaMemcpyAsync(v->p, host, n * v->esize(), cudaMemcpyHostToDevice, v->ctx->stream)); MFG_CUDA(cudaStreamSynchronize(v->ctx->stream)); }
  });
}
int mfg_vec_to_host(const mfg_vec *v, void *host, size_t n)
{
  return guarded([&] {
    MFG_REQUIRE(v && host, "null argument");
    MFG_REQUIRE(n == v->n, "copyToHost: size mismatch");
    if (n) { MFG_CUDA(cudaMemcpyAsync(host, v->p, n * v->esize(), cudaMemcpyDeviceToHost, v->ctx->stream)); MFG_CUDA(cudaStreamSynchronize(v->ctx->stream)); }
  });
}
int mfg_vec_copy(mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] {
    MFG_REQUIRE(dst && src, "null argument");
    if (dst->n != src->n) { int rc = mfg_vec_resize(dst, src->n); if (rc) throw Error(rc, mfg_last_error()); }
    if (dst->dt == src->dt) { if (src->n) MFG_CUDA(cudaMemcpyAsync(dst->p, src->p, src->n * src->esize(), cudaMemcpyDeviceToDevice, dst->ctx->stream)); }
    else vec_equ(dst, 1.0, src);
  });
}
int mfg_vec_swap(mfg_vec *a, mfg_vec *b)
{
  return guarded([&] {
    MFG_REQUIRE(a && b, "null argument");
    MFG_REQUIRE(a->dt == b->dt, "swap: dtypes differ");
    std::swap(a->p, b->p); std::swap(a->n, b->n); std::swap(a->cap, b->cap); std::swap(a->owns, b->owns);
  });
}
int mfg_vec_fill(mfg_vec *v, double a) { return guarded([&] { MFG_REQUIRE(v, "null vector"); vec_fill(v, a); }); }
int mfg_vec_sadd(mfg_vec *v, double s, double a, const mfg_vec *x) { return guarded([&] { MFG_REQUIRE(v && x, "null vector"); vec_sadd(v, s, a, x); }); }
int mfg_vec_equ(mfg_vec *v, double a, const mfg_vec *x) { return guarded([&] { MFG_REQUIRE(v && x, "null vector"); vec_equ(v, a, x); }); }
int mfg_vec_scale(mfg_vec *v, const mfg_vec *x) { return guarded([&] { MFG_REQUIRE(v && x, "null vector"); vec_scale(v, x); }); }
int mfg_vec_divide(mfg_vec *v, const mfg_vec *x) { return guarded([&] { MFG_REQUIRE(v && x, "null vector"); vec_divide(v, x); }); }
int mfg_vec_invert(mfg_vec *v) { return guarded([&] { MFG_REQUIRE(v, "null vector"); vec_invert(v); }); }
int mfg_vec_scal(mfg_vec *v, double a) { return guarded([&] { MFG_REQUIRE(v, "null vector"); vec_scal(v, a); }); }
int mfg_vec_dot(const mfg_vec *a, const mfg_vec *b, double *out) { return guarded([&] { MFG_REQUIRE(a && b && out, "null argument"); *out = vec_dot(a, b); }); }
int mfg_vec_add_and_dot(mfg_vec *v, double a, const mfg_vec *x, const mfg_vec *w, double *out)
{
  return guarded([&] { MFG_REQUIRE(v && x && w && out, "null argument"); *out = vec_add_and_dot(v, a, x, w); });
}
int mfg_vec_l2_norm(const mfg_vec *v, double *out) { return guarded([&] { MFG_REQUIRE(v && out, "null argument"); *out = std::sqrt(vec_dot(v, v)); }); }
int mfg_vec_all_zero(const mfg_vec *v, int *out) { return guarded([&] { MFG_REQUIRE(v && out, "null argument"); *out = vec_all_zero(v) ? 1 : 0; }); }
int mfg_vec_copy_with_indices(mfg_vec *dst, const mfg_vec *src, const uint32_t *dst_idx, const uint32_t *src_idx, size_t n)
{
  return guarded([&] { MFG_REQUIRE(dst && src && (n == 0 || (dst_idx && src_idx)), "null argument"); vec_copy_with_indices(dst, src, dst_idx, src_idx, n); });
}

// ---- mesh ---------------------------------------------------------------------
int mfg_mesh_create_box(mfg_ctx *ctx, const mfg_box_desc *desc, mfg_mesh **out)
{
  return guarded([&] { MFG_REQUIRE(ctx && desc && out, "null argument"); *out = build_box_mesh(ctx, *desc); });
}
int mfg_mesh_hyper_cube(mfg_ctx *ctx, int dim, int degree, int n_refine, double left, double right, mfg_mesh **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && out, "null argument");
    MFG_REQUIRE(right > left, "hyper_cube: right must exceed left");
    MFG_REQUIRE(n_refine >= 0 && n_refine <= 10, "n_refine out of range");
    mfg_box_desc d;
    d.dim = dim; d.degree = degree; d.h = (right - left) / (double)(1u << n_refine); d.dirichlet_faces = 0x3f;
    for (int k = 0; k < 3; ++k) { d.log2_cells[k] = n_refine; d.origin[k] = left; }
    *out = build_box_mesh(ctx, d);
  });
}
int mfg_mesh_destroy(mfg_mesh *m) { return guarded([&] { delete m; }); }
uint32_t mfg_mesh_n_cells(const mfg_mesh *m) { return m ? m->n_cells : 0; }
uint32_t mfg_mesh_n_dofs(const mfg_mesh *m) { return m ? m->n_dofs : 0; }
uint32_t mfg_mesh_dofs_per_cell(const mfg_mesh *m) { return m ? m->npc : 0; }
uint32_t mfg_mesh_n_constrained(const mfg_mesh *m) { return m ? m->n_constrained : 0; }
const uint32_t *mfg_mesh_loc2glob_device(const mfg_mesh *m) { return m ? m->l2g.p : nullptr; }
const uint32_t *mfg_mesh_constrained_device(const mfg_mesh *m) { return m ? m->constrained.p : nullptr; }
int mfg_mesh_get_loc2glob(const mfg_mesh *m, uint32_t *host) { return guarded([&] { MFG_REQUIRE(m && host, "null argument"); m->l2g.download(host, m->ctx->stream); }); }
int mfg_mesh_get_constrained(const mfg_mesh *m, uint32_t *host) { return guarded([&] { MFG_REQUIRE(m && (host || !m->n_constrained), "null argument"); m->constrained.download(host, m->ctx->stream); }); }
int mfg_mesh_get_cell_coords(const mfg_mesh *m, uint32_t *host) { return guarded([&] { MFG_REQUIRE(m && host, "null argument"); mesh_cell_coords(m, host); }); }
int mfg_mesh_get_support_points(const mfg_mesh *m, double *host) { return guarded([&] { MFG_REQUIRE(m && host, "null argument"); mesh_support_points(m, host); }); }
int mfg_mesh_lattice_to_dof(const mfg_mesh *m, size_t n, const uint32_t *lattice_xyz, uint32_t *dof)
{
  return guarded([&] {
    MFG_REQUIRE(m && (n == 0 || (lattice_xyz && dof)), "null argument");
    for (size_t i = 0; i < n; ++i)
      for (int d = 0; d < m->dim; ++d) MFG_REQUIRE(lattice_xyz[3 * i + d] <= (uint32_t)m->p * m->nc[d], "lattice point outside the mesh");
    mesh_lattice_to_dof(m, n, lattice_xyz, dof);
  });
}
int mfg_mesh_color_cells(const mfg_mesh *m, uint32_t *color_of_cell, uint32_t *n_colors)
{
  return guarded([&] {
    MFG_REQUIRE(m && color_of_cell && n_colors, "null argument");
    std::vector<uint32_t> col; uint32_t nc = 0;
    mesh_parity_colors(m, col, nc);
    std::copy(col.begin(), col.end(), color_of_cell); *n_colors = nc;
  });
}

// ---- MatrixFreeGpu --------------------------------------------------------------
int mfg_mf_reinit(mfg_ctx *ctx, const mfg_mf_desc *desc, mfg_mf **out)
{
  return guarded([&] { MFG_REQUIRE(ctx && desc && out, "null argument"); *out = mf_from_desc(ctx, *desc); });
}
int mfg_mf_reinit_from_mesh(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_scatter scatter, mfg_mf **out)
{
  return guarded([&] { MFG_REQUIRE(ctx && mesh && out, "null argument"); *out = mf_from_mesh(ctx, mesh, dt, scatter); });
}
int mfg_mf_destroy(mfg_mf *mf) { return guarded([&] { delete mf; }); }
uint32_t mfg_mf_n_dofs(const mfg_mf *mf) { return mf ? mf->n_dofs : 0; }
uint32_t mfg_mf_n_cells(const mfg_mf *mf) { return mf ? mf->n_cells : 0; }
uint32_t mfg_mf_n_colors(const mfg_mf *mf) { return mf ? mf->n_colors() : 0; }
size_t mfg_mf_memory_consumption(const mfg_mf *mf)
{
  if (!mf) return 0;
  return mf->idx.bytes() + mf->cell_perm.bytes() + mf->color_offsets.size() * sizeof(uint32_t);
}
int mfg_shape_info(int degree, double *shape_values, double *shape_gradients, double *q_points, double *q_weights)
{
  return guarded([&] {
    MFG_REQUIRE(degree >= 1 && degree <= 8, "degree must be in 1..8");
    const FEData1D fe = make_fe_data(degree);
    const int n = fe.n;
    if (shape_values) std::copy(fe.val.begin(), fe.val.end(), shape_values);
    if (shape_gradients) std::copy(fe.grad.begin(), fe.grad.end(), shape_gradients);
    if (q_points) std::copy(fe.qpts.begin(), fe.qpts.begin() + n, q_points);
    if (q_weights) std::copy(fe.qwts.begin(), fe.qwts.begin() + n, q_weights);
  });
}

int mfg_hanging_node_weights(int degree, double *weights)
{
  return guarded([&] {
    MFG_REQUIRE(degree >= 1 && degree <= 8 && weights, "degree must be in 1..8 and weights non-null");
    const FEData1D fe = make_fe_data(degree);
    std::copy(fe.hanging.begin(), fe.hanging.end(), weights);
  });
}

// ---- ConstraintHandlerGpu ---------------------------------------------------------
int mfg_ch_create(mfg_ctx *ctx, mfg_dtype dt, const uint32_t *constrained_host, size_t n_constrained, const uint32_t *edge_host, size_t n_edge, mfg_ch **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && out && (constrained_host || !n_constrained) && (edge_host || !n_edge), "null argument");
    *out = ch_create(ctx, dt, constrained_host, n_constrained, edge_host, n_edge);
  });
}
int mfg_ch_create_from_mesh(mfg_ctx *ctx, mfg_dtype dt, const mfg_mesh *mesh, mfg_ch **out)
{
  return guarded([&] { MFG_REQUIRE(ctx && mesh && out, "null argument"); *out = ch_from_mesh(ctx, dt, mesh); });
}
int mfg_ch_destroy(mfg_ch *ch) { return guarded([&] { delete ch; }); }
size_t mfg_ch_n_constrained(const mfg_ch *ch) { return ch ? ch->n() : 0; }
int mfg_ch_set_constrained_values(mfg_ch *ch, mfg_vec *v, double val) { return guarded([&] { MFG_REQUIRE(ch && v, "null argument"); ch_set(ch, v, val); }); }
int mfg_ch_save_constrained_values(mfg_ch *ch, mfg_vec *v) { return guarded([&] { MFG_REQUIRE(ch && v, "null argument"); ch_save(ch, v); }); }
int mfg_ch_save_constrained_values2(mfg_ch *ch, const mfg_vec *v1, mfg_vec *v2) { return guarded([&] { MFG_REQUIRE(ch && v1 && v2, "null argument"); ch_save2(ch, v1, v2); }); }
int mfg_ch_load_constrained_values(mfg_ch *ch, mfg_vec *v) { return guarded([&] { MFG_REQUIRE(ch && v, "null argument"); ch_load(ch, v); }); }
int mfg_ch_load_and_add_constrained_values(mfg_ch *ch, mfg_vec *v1, mfg_vec *v2) { return guarded([&] { MFG_REQUIRE(ch && v1 && v2, "null argument"); ch_load_and_add(ch, v1, v2); }); }
int mfg_ch_copy_edge_values(mfg_ch *ch, mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] { MFG_REQUIRE(ch && dst && src, "null argument"); vec_copy_with_indices(dst, src, ch->edge.p, ch->edge.p, ch->edge.n); });
}

// ---- LaplaceOperatorGpu -------------------------------------------------------------
int mfg_laplace_create(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_scatter scatter, mfg_laplace **out)
{
  return guarded([&] { MFG_REQUIRE(ctx && mesh && out, "null argument"); *out = laplace_from_mesh(ctx, mesh, dt, scatter); });
}
int mfg_laplace_create_from_arrays(mfg_ctx *ctx, mfg_mf *mf, mfg_ch *ch, const double *coefficient_host, mfg_laplace **out)
{
  return guarded([&] { MFG_REQUIRE(ctx && out, "null argument"); *out = laplace_from_arrays(ctx, mf, ch, coefficient_host); });
}
int mfg_laplace_set_coefficient(mfg_laplace *op, const double *coefficient_host)
{
  return guarded([&] { MFG_REQUIRE(op && coefficient_host, "null argument"); laplace_set_coefficient_host(op, coefficient_host); });
}
int mfg_laplace_destroy(mfg_laplace *op)
{
  return guarded([&] {
    if (!op) return;
    if (op->inv_diag && op->inv_diag->p) cudaFree(op->inv_diag->p);
    for (cudaEvent_t e : op->ev) cudaEventDestroy(e);
    for (auto &sl : op->slots) { if (sl.h2d) { cudaEventDestroy(sl.h2d); cudaEventDestroy(sl.done); cudaEventDestroy(sl.d2h); } }
    if (op->h2d_stream) { cudaStreamDestroy(op->h2d_stream); cudaStreamDestroy(op->d2h_stream); }
    if (op->owns_mf) delete op->mf;
    if (op->owns_ch) delete op->ch;
    delete op;
  });
}
uint32_t mfg_laplace_m(const mfg_laplace *op) { return op ? op->mf->n_dofs : 0; }
int mfg_laplace_set_option(mfg_laplace *op, const char *name, int value)
{
  return guarded([&] {
    MFG_REQUIRE(op && name, "null argument");
    const std::string k(name);
    if (k == "cg_fused") op->cg_fused = value != 0;
    else if (k == "stage_sync") op->stage_sync = value != 0;
    else if (k == "stage_merge_dirs") { op->stage_merge_dirs = value & 7; op->st_built = false; }
    else throw Error(MFG_ERR_INVALID, "unknown option '" + k + "'");
  });
}
int mfg_laplace_set_variant(mfg_laplace *op, int variant) { return guarded([&] { MFG_REQUIRE(op, "null operator"); op->variant = variant; }); }

static void check_vecs(const mfg_laplace *op, const mfg_vec *dst, const mfg_vec *src)
{
  MFG_REQUIRE(op && dst && src, "null argument");
  MFG_REQUIRE(dst->dt == op->mf->dt && src->dt == op->mf->dt, "vector dtype differs from operator dtype");
  MFG_REQUIRE(dst->n == op->mf->n_dofs && src->n == op->mf->n_dofs, "vector size differs from operator size");
}
int mfg_laplace_vmult(mfg_laplace *op, mfg_vec *dst, const mfg_vec *src) { return guarded([&] { check_vecs(op, dst, src); laplace_vmult(op, dst->p, src->p, false); }); }
int mfg_laplace_vmult_add(mfg_laplace *op, mfg_vec *dst, const mfg_vec *src) { return guarded([&] { check_vecs(op, dst, src); laplace_vmult(op, dst->p, src->p, true); }); }
int mfg_laplace_vmult_ptr(mfg_laplace *op, void *dst_dev, const void *src_dev) { return guarded([&] { MFG_REQUIRE(op && dst_dev && src_dev, "null argument"); laplace_vmult(op, dst_dev, src_dev, false); }); }
int mfg_laplace_vmult_add_ptr(mfg_laplace *op, void *dst_dev, const void *src_dev) { return guarded([&] { MFG_REQUIRE(op && dst_dev && src_dev, "null argument"); laplace_vmult(op, dst_dev, src_dev, true); }); }
int mfg_mf_get_gpu_data(mfg_mf *mf, mfg_gpu_data *out) { return guarded([&] { MFG_REQUIRE(mf && out, "null argument"); mf_get_gpu_data(mf, out); }); }
int mfg_laplace_set_interface_dofs(mfg_laplace *op, const uint32_t *dofs_host, size_t n, uint32_t *n_interface_groups)
{
  return guarded([&] {
    MFG_REQUIRE(op && (dofs_host || !n), "null argument");
    const uint32_t k = laplace_set_interface_dofs(op, dofs_host, n);
    if (n_interface_groups) *n_interface_groups = k;
  });
}
int mfg_laplace_vmult_part_ptr(mfg_laplace *op, void *dst_dev, const void *src_dev, int part, void *cuda_stream)
{
  return guarded([&] { MFG_REQUIRE(op && dst_dev && src_dev, "null argument"); laplace_vmult(op, dst_dev, src_dev, false, part, cuda_stream); });
}
int mfg_laplace_vmult_host(mfg_laplace *op, void *dst_host, const void *src_host)
{
  return guarded([&] {
    MFG_REQUIRE(op && dst_host && src_host, "null argument");
    const size_t bytes = (size_t)op->mf->n_dofs * (op->mf->dt == MFG_F64 ? 8 : 4);
    if (op->host_stage_src.n != bytes) { op->host_stage_src.alloc(bytes); op->host_stage_dst.alloc(bytes); }
    cudaStream_t s = op->ctx->stream;
    MFG_CUDA(cudaMemcpyAsync(op->host_stage_src.p, src_host, bytes, cudaMemcpyHostToDevice, s));
    laplace_vmult(op, op->host_stage_dst.p, op->host_stage_src.p, false);
    MFG_CUDA(cudaMemcpyAsync(dst_host, op->host_stage_dst.p, bytes, cudaMemcpyDeviceToHost, s));
    MFG_CUDA(cudaStreamSynchronize(s));
  });
}
int mfg_laplace_vmult_host_async(mfg_laplace *op, void *dst_host, const void *src_host, int slot)
{
  return guarded([&] {
    MFG_REQUIRE(op && dst_host && src_host, "null argument");
    MFG_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
    const size_t bytes = (size_t)op->mf->n_dofs * (op->mf->dt == MFG_F64 ? 8 : 4);
    if (!op->h2d_stream)
      {
        MFG_CUDA(cudaStreamCreateWithFlags(&op->h2d_stream, cudaStreamNonBlocking));
        MFG_CUDA(cudaStreamCreateWithFlags(&op->d2h_stream, cudaStreamNonBlocking));
      }
    mfg_laplace::HostSlot &sl = op->slots[slot];
    if (sl.src.n != bytes)
      {
        sl.src.alloc(bytes); sl.dst.alloc(bytes);
        if (!sl.h2d)
          {
            MFG_CUDA(cudaEventCreateWithFlags(&sl.h2d, cudaEventDisableTiming));
            MFG_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
            MFG_CUDA(cudaEventCreateWithFlags(&sl.d2h, cudaEventDisableTiming));
          }
        sl.used = false;
      }
    cudaStream_t cs = op->ctx->stream;
    // the slot's staging buffers are free once its previous apply has been computed (src) and copied out (dst)
    if (sl.used) MFG_CUDA(cudaStreamWaitEvent(op->h2d_stream, sl.done, 0));
    MFG_CUDA(cudaMemcpyAsync(sl.src.p, src_host, bytes, cudaMemcpyHostToDevice, op->h2d_stream));
    MFG_CUDA(cudaEventRecord(sl.h2d, op->h2d_stream));
    MFG_CUDA(cudaStreamWaitEvent(cs, sl.h2d, 0));
    if (sl.used) MFG_CUDA(cudaStreamWaitEvent(cs, sl.d2h, 0));
    laplace_vmult(op, sl.dst.p, sl.src.p, false);
    MFG_CUDA(cudaEventRecord(sl.done, cs));
    MFG_CUDA(cudaStreamWaitEvent(op->d2h_stream, sl.done, 0));
    MFG_CUDA(cudaMemcpyAsync(dst_host, sl.dst.p, bytes, cudaMemcpyDeviceToHost, op->d2h_stream));
    MFG_CUDA(cudaEventRecord(sl.d2h, op->d2h_stream));
    sl.used = true;
  });
}
int mfg_laplace_host_sync(mfg_laplace *op)
{
  return guarded([&] {
    MFG_REQUIRE(op, "null operator");
    if (op->h2d_stream) { MFG_CUDA(cudaStreamSynchronize(op->h2d_stream)); MFG_CUDA(cudaStreamSynchronize(op->d2h_stream)); }
    MFG_CUDA(cudaStreamSynchronize(op->ctx->stream));
  });
}
int mfg_laplace_compute_diagonal(mfg_laplace *op) { return guarded([&] { MFG_REQUIRE(op, "null operator"); laplace_compute_diagonal(op); }); }
int mfg_laplace_get_diagonal_inverse(mfg_laplace *op, mfg_vec **out)
{
  return guarded([&] {
    MFG_REQUIRE(op && out, "null argument");
    // Assert(diagonal_is_available, ExcNotInitialized()) laplace_operator_gpu.h:427
    MFG_REQUIRE(op->diagonal_is_available, "get_diagonal_inverse: compute_diagonal has not been called");
    *out = op->inv_diag.get();
  });
}
size_t mfg_laplace_memory_consumption(const mfg_laplace *op)
{
  if (!op) return 0;
  return mfg_mf_memory_consumption(op->mf) + op->cw.bytes() + op->cbits.bytes() + op->ch->constrained.bytes() + op->ch->edge.bytes() +
         op->ch->tmp_src.bytes() + op->ch->tmp_dst.bytes() + (op->inv_diag ? op->inv_diag->n * op->inv_diag->esize() : 0);
}
int mfg_laplace_enable_kernel_timing(mfg_laplace *op, int on)
{
  return guarded([&] { MFG_REQUIRE(op, "null operator"); op->timing = on != 0; op->timing_stride = on > 1 ? on : 1; op->timing_counter = 0; if (!on) op->ev_used = 0; });
}
int mfg_laplace_kernel_time_ms(mfg_laplace *op, double *total_ms, int *n_launches)
{
  return guarded([&] { MFG_REQUIRE(op, "null operator"); laplace_kernel_time(op, total_ms, n_launches); });
}
int mfg_laplace_active_variant(const mfg_laplace *op)
{
  // (a requested variant the operator cannot run throws inside: no exception may cross the C boundary -- -1, mfg_last_error says why)
  int       v = 0;
  const int rc = guarded([&] { v = op ? laplace_active_variant(op) : 0; });
  return rc == MFG_OK ? v : -1;
}
int mfg_laplace_stage_stats(const mfg_laplace *op, uint32_t out[8])
{
  return guarded([&] { MFG_REQUIRE(op && out, "null argument"); for (int i = 0; i < 8; ++i) out[i] = op->st_stats[i]; });
}
struct mfg_stage_plan { mfg::StagePlan plan; mfg::StageGeom geom; };
int mfg_stage_plan_build(int degree, mfg_dtype dt, uint32_t n_plain, uint32_t n_cells, uint32_t n_dofs, const uint32_t *idx_host, int merge_dirs,
                         mfg_stage_plan **out)
{
  return guarded([&] {
    MFG_REQUIRE(out && idx_host, "null argument");
    MFG_REQUIRE(n_plain <= n_cells, "n_plain must not exceed n_cells");
    std::unique_ptr<mfg_stage_plan> p(new mfg_stage_plan);
    p->geom = stage_geom(degree, dt);
    StagePlanIn in;
    in.n = p->geom.n; in.cw = p->geom.cw; in.hc = p->geom.hc; in.wb = dt == MFG_F64 ? 8 : 4; in.xcap = p->geom.xcap; in.hmax = p->geom.hmax; in.ocap = p->geom.ocap; in.lcap = p->geom.lcap; in.nclass = p->geom.nclass;
    in.n_plain = n_plain; in.n_cells = n_cells; in.n_dofs = n_dofs; in.idx = idx_host; in.merge_dirs = merge_dirs;
    build_stage_plan(in, p->plan);
    *out = p.release();
  });
}
int mfg_stage_plan_info(const mfg_stage_plan *p, uint32_t info[16])
{
  return guarded([&] {
    MFG_REQUIRE(p && info, "null argument");
    const uint32_t v[16] = {p->plan.n_groups, p->plan.n_patterns, (uint32_t)p->plan.pstride, (uint32_t)p->plan.halo.size(), (uint32_t)p->plan.fallback.size(),
                            (uint32_t)p->geom.cw, (uint32_t)p->geom.hc, (uint32_t)p->geom.xcap, (uint32_t)p->geom.n, p->plan.n_staged, (uint32_t)p->geom.lcap, (uint32_t)p->geom.ocap,
                            (uint32_t)(p->plan.rd_wavefronts / std::max<uint32_t>(1, p->plan.n_staged)), (uint32_t)(p->plan.wr_wavefronts / std::max<uint32_t>(1, p->plan.n_staged)),
                            (uint32_t)(p->plan.cp_wavefronts / std::max<uint32_t>(1, p->plan.n_staged)), 0};
    for (int i = 0; i < 16; ++i) info[i] = v[i];
  });
}
int mfg_stage_plan_get(const mfg_stage_plan *p, uint32_t *gdesc, uint32_t *halo, uint16_t *ptab, uint32_t *fallback)
{
  return guarded([&] {
    MFG_REQUIRE(p, "null argument");
    if (gdesc) std::copy(p->plan.gdesc.begin(), p->plan.gdesc.end(), gdesc);
    if (halo) std::copy(p->plan.halo.begin(), p->plan.halo.end(), halo);
    if (ptab) std::copy(p->plan.ptab.begin(), p->plan.ptab.end(), ptab);
    if (fallback) std::copy(p->plan.fallback.begin(), p->plan.fallback.end(), fallback);
  });
}
int mfg_stage_plan_destroy(mfg_stage_plan *p) { return guarded([&] { delete p; }); }
int mfg_laplace_launches_per_vmult(const mfg_laplace *op) { return op ? laplace_launches_per_vmult(op) : 0; }
int mfg_laplace_cell_launches_per_vmult(const mfg_laplace *op) { return op ? (int)op->mf->n_colors() + (op->mf->hn_mask.n ? 1 : 0) : 0; }
int mfg_laplace_bmop(mfg_laplace *op, mfg_vec *dst, mfg_vec *src, int k, double init, float *elapsed_ms)
{
  return guarded([&] {
    check_vecs(op, dst, src);
    MFG_REQUIRE(k >= 0, "k must be non-negative");
    cudaStream_t s = op->ctx->stream;
    cudaEvent_t e0, e1;
    MFG_CUDA(cudaEventCreate(&e0)); MFG_CUDA(cudaEventCreate(&e1));
    vec_fill(dst, init);
    MFG_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < k; ++i)
      {
        std::swap(dst->p, src->p); std::swap(dst->cap, src->cap); std::swap(dst->owns, src->owns);  // dst.swap(src)  bmop.cu:143
        laplace_vmult(op, dst->p, src->p, false);
      }
    MFG_CUDA(cudaEventRecord(e1, s));
    MFG_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MFG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (elapsed_ms) *elapsed_ms = ms;
  });
}

}  // extern "C"
