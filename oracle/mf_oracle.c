/*
 * mf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the algorithm of the reference's hot path
 * (kalj/dealii-cuda: LaplaceOperatorGpu::vmult and its deal.II MatrixFree CPU
 * twin LaplaceOperatorCpu::vmult) for uniform hyper_cube meshes, plus the
 * pieces of deal.II the reference delegates to (FE_Q numbering, Gauss/GLL
 * data, boundary constraints, graph coloring) restated from their published
 * behaviour.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may load this library; the product (libmfgpu.so) never does.
 *
 * PARITY STATUS: "parity unpinned" at the deal.II boundary.  deal.II is not in
 * the reference tree nor in this image, and the reference's tests hold no
 * golden vectors for this path.  The oracle is pinned against
 * (1) the DoF-map / norm fixtures of SURVEY.md Appendix A.3 / B (themselves a
 * restatement, not deal.II output), (2) an independent assembled-matrix
 * operator following the procedure of test_laplace_op.cu:50-120, and (3)
 * analytic identities (A*1=0, u^T A u = |Omega| for u=x, symmetry).
 *
 * Reference anchors (file:line relative to /root/reference):
 *   operator semantics      laplace_operator_cpu.cc:125-143, 180-211
 *                           laplace_operator_gpu.h:216-223, 247-303
 *   cell kernel pieces      matrix_free_gpu/fee_gpu.cuh:197-365
 *                           matrix_free_gpu/tensor_ops.cuh:179-261
 *   data layout             matrix_free_gpu/matrix_free_gpu.cu:283-339
 *   coefficient             poisson_common.h:146-158
 *   mesh                    poisson_common.h:58-72, bmop_common.h:108-120
 *   benchmark loop          bmop.cu:135-153, bmop-cpu.cc:138-155
 *   diagonal                laplace_operator_gpu.h:355-421, laplace_operator_cpu.cc:294-353
 *   assembled check         test_laplace_op.cu:50-120
 *   constraints list        matrix_free_gpu/constraint_handler_gpu.cu:69-95
 *   coloring                matrix_free_gpu/coloring.cc:8-33
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXN 9 /* degree <= 8 */

typedef struct orc_mesh
{
  int      dim, p, n, r;
  uint32_t N;       /* cells per direction = 2^r (cube meshes) */
  int      lg[3];   /* log2 cells per direction (box meshes) */
  uint32_t nc[3];
  double   origin[3];
  uint32_t dirichlet_faces;
  double   left, right, h;
  uint32_t n_cells, n_dofs, npc; /* npc = n^dim */
  uint32_t *l2g;    /* [n_cells][npc], lexicographic local order (x fastest) */
  uint32_t *cxyz;   /* [n_cells][3] integer cell coordinates, cells in Morton order */
  uint32_t *dof_lattice; /* [n_dofs][3] lattice coordinate (0..p*N) of every DoF */
  uint32_t n_constrained;
  uint32_t *constrained;   /* ascending */
  uint8_t  *is_constrained;
  double   *coef;   /* [n_cells][npc] at Gauss points, q lexicographic */
  double   xq[ORC_MAXN], wq[ORC_MAXN], xn[ORC_MAXN];
  double   sv[ORC_MAXN * ORC_MAXN], sg[ORC_MAXN * ORC_MAXN]; /* [i*n+q] */
  uint32_t *lex2hier; /* hierarchic index of lexicographic local dof i */
  /* parity coloring for the threaded baseline */
  uint32_t *color_cells; uint32_t color_off[9];
} orc_mesh;

/* ------------------------------------------------------------------------ */
/* 1-D data: Gauss-Legendre (deal.II QGauss) and Gauss-Lobatto (FE_Q nodes)  */
/* ------------------------------------------------------------------------ */

static void legendre(int n, long double x, long double *P, long double *dP)
{
  long double p0 = 1.0L, p1 = x;
  if (n == 0) { *P = 1.0L; *dP = 0.0L; return; }
  for (int k = 2; k <= n; ++k)
    {
      long double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
      p0 = p1; p1 = p2;
    }
  *P  = p1;
  *dP = n * (x * p1 - p0) / (x * x - 1.0L);
}

/* n-point Gauss-Legendre on [0,1] */
void orc_gauss(int n, double *x, double *w)
{
  for (int i = 0; i < n; ++i)
    {
      long double z = -cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L));
      long double P, dP;
      for (int it = 0; it < 100; ++it)
        {
          legendre(n, z, &P, &dP);
          long double dz = P / dP;
          z -= dz;
          if (fabsl(dz) < 1e-19L) break;
        }
      legendre(n, z, &P, &dP);
      x[i] = (double)(0.5L * (z + 1.0L));
      w[i] = (double)(1.0L / ((1.0L - z * z) * dP * dP));
    }
}

/* n-point Gauss-Lobatto nodes on [0,1] (roots of P'_{n-1} and the end points) */
void orc_gauss_lobatto(int n, double *x)
{
  const int m = n - 1;
  x[0] = 0.0; x[n - 1] = 1.0;
  for (int i = 1; i < n - 1; ++i)
    {
      long double z = -cosl(3.14159265358979323846264338327950288L * i / m);
      for (int it = 0; it < 100; ++it)
        {
          long double P, dP;
          legendre(m, z, &P, &dP);
          /* f = P'_m ; f' = P''_m = (2 z P' - m(m+1) P)/(1-z^2) */
          long double ddP = (2.0L * z * dP - m * (m + 1.0L) * P) / (1.0L - z * z);
          long double dz  = dP / ddP;
          z -= dz;
          if (fabsl(dz) < 1e-19L) break;
        }
      x[i] = (double)(0.5L * (z + 1.0L));
    }
  /* symmetrise */
  for (int i = 0; i < n / 2; ++i)
    {
      double a = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
      x[i] = a; x[n - 1 - i] = 1.0 - a;
    }
  if (n % 2) x[n / 2] = 0.5;
}

/* Lagrange basis on nodes xn evaluated at points xq: val[i*nq+q], grad[i*nq+q]
 * (matrix_free_gpu.cu:502-513: shape_values[i*n+q] = phi_i(x_q)) */
static void lagrange_eval(int n, const double *xn, int nq, const double *xq, double *val, double *grad)
{
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < nq; ++q)
      {
        long double v = 1.0L, g = 0.0L;
        for (int m = 0; m < n; ++m)
          if (m != i) v *= ((long double)xq[q] - xn[m]) / ((long double)xn[i] - xn[m]);
        for (int l = 0; l < n; ++l)
          {
            if (l == i) continue;
            long double t = 1.0L / ((long double)xn[i] - xn[l]);
            for (int m = 0; m < n; ++m)
              if (m != i && m != l) t *= ((long double)xq[q] - xn[m]) / ((long double)xn[i] - xn[m]);
            g += t;
          }
        val[i * nq + q]  = (double)v;
        grad[i * nq + q] = (double)g;
      }
}

void orc_shape_1d(int p, double *val, double *grad, double *nodes, double *qpts, double *qwts)
{
  const int n = p + 1;
  double xn[ORC_MAXN], xq[ORC_MAXN], wq[ORC_MAXN];
  orc_gauss_lobatto(n, xn);
  orc_gauss(n, xq, wq);
  lagrange_eval(n, xn, n, xq, val, grad);
  if (nodes) memcpy(nodes, xn, n * sizeof(double));
  if (qpts) memcpy(qpts, xq, n * sizeof(double));
  if (qwts) memcpy(qwts, wq, n * sizeof(double));
}

/* ------------------------------------------------------------------------ */
/* FE_Q hierarchic -> lexicographic                 */
/* h2l[hier] = lexicographic index                                           */
/* ------------------------------------------------------------------------ */
void orc_hier_to_lex(int dim, int p, uint32_t *h2l)
{
  const uint32_t n = p + 1, L = p - 1;
  uint32_t c = 0;
  if (dim == 2)
    {
      h2l[c++] = 0; h2l[c++] = p; h2l[c++] = n * p; h2l[c++] = n * p + p;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = (i + 1) * n;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = (i + 2) * n - 1;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = 1 + i;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = 1 + i + n * (n - 1);
      for (uint32_t i = 0; i < L; ++i)
        for (uint32_t j = 0; j < L; ++j) h2l[c++] = n * (i + 1) + j + 1;
    }
  else
    {
      const uint32_t n2 = n * n;
      h2l[c++] = 0; h2l[c++] = p; h2l[c++] = n * p; h2l[c++] = (n + 1) * p;
      h2l[c++] = n2 * p; h2l[c++] = (n2 + 1) * p; h2l[c++] = (n2 + n) * p; h2l[c++] = (n2 + n + 1) * p;
      /* lines 0..11 */
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = (i + 1) * n;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = n - 1 + (i + 1) * n;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = 1 + i;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = 1 + i + n * (n - 1);
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = (n - 1) * n2 + (i + 1) * n;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = (n - 1) * (n2 + 1) + (i + 1) * n;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = n2 * (n - 1) + i + 1;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = n2 * (n - 1) + i + 1 + n * (n - 1);
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = (i + 1) * n2;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = n - 1 + (i + 1) * n2;
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = (i + 1) * n2 + n * (n - 1);
      for (uint32_t i = 0; i < L; ++i) h2l[c++] = n - 1 + (i + 1) * n2 + n * (n - 1);
      /* quads 0..5 */
      for (uint32_t i = 0; i < L; ++i) for (uint32_t j = 0; j < L; ++j) h2l[c++] = (i + 1) * n2 + n * (j + 1);
      for (uint32_t i = 0; i < L; ++i) for (uint32_t j = 0; j < L; ++j) h2l[c++] = (i + 1) * n2 + n - 1 + n * (j + 1);
      for (uint32_t i = 0; i < L; ++i) for (uint32_t j = 0; j < L; ++j) h2l[c++] = (j + 1) * n2 + i + 1;
      for (uint32_t i = 0; i < L; ++i) for (uint32_t j = 0; j < L; ++j) h2l[c++] = (j + 1) * n2 + n * (n - 1) + i + 1;
      for (uint32_t i = 0; i < L; ++i) for (uint32_t j = 0; j < L; ++j) h2l[c++] = n * (i + 1) + j + 1;
      for (uint32_t i = 0; i < L; ++i) for (uint32_t j = 0; j < L; ++j) h2l[c++] = (n - 1) * n2 + n * (i + 1) + j + 1;
      /* hex interior */
      for (uint32_t i = 0; i < L; ++i)
        for (uint32_t j = 0; j < L; ++j)
          for (uint32_t k = 0; k < L; ++k) h2l[c++] = n2 * (i + 1) + n * (j + 1) + k + 1;
    }
}

/* ------------------------------------------------------------------------ */
/* mesh + DoF numbering                                                      */
/* ------------------------------------------------------------------------ */

static uint32_t ipow_u(uint32_t b, int e) { uint32_t r = 1; while (e-- > 0) r *= b; return r; }

/* Coefficient<dim,Number>::value, poisson_common.h:146-158 */
static inline double coefficient_value(const double *x, int dim)
{
  double s = 0; for (int d = 0; d < dim; ++d) s += x[d] * x[d];
  return 1.0 / (0.05 + 2.0 * s);
}

void orc_destroy(orc_mesh *m)
{
  if (!m) return;
  free(m->l2g); free(m->cxyz); free(m->dof_lattice); free(m->constrained); free(m->is_constrained);
  free(m->coef); free(m->lex2hier); free(m->color_cells); free(m);
}

/* Box of 2^lg[d] cells per direction with edge h and lower corner origin, FE_Q(p), homogeneous Dirichlet data
 * on the faces selected by dirichlet_faces (bit 2d: lower face of direction d, bit 2d+1: upper face).
 * hyper_cube(left,right)^dim + refine_global(r) (bmop.cu:111-132) is the special case lg = (r,r,r), all faces.
 * Cells are ordered along the (generalised) Morton curve, x least significant. */
orc_mesh *orc_create_box(int dim, int p, const int *lg, const double *origin, double h, uint32_t dirichlet_faces)
{
  if (dim < 2 || dim > 3 || p < 1 || p > 8) return NULL;
  orc_mesh *m = (orc_mesh *)calloc(1, sizeof(orc_mesh));
  m->dim = dim; m->p = p; m->n = p + 1; m->h = h; m->dirichlet_faces = dirichlet_faces;
  m->n_cells = 1;
  for (int d = 0; d < 3; ++d)
    {
      m->lg[d] = d < dim ? lg[d] : 0; m->nc[d] = 1u << m->lg[d]; m->origin[d] = d < dim ? origin[d] : 0.0;
      m->n_cells *= m->nc[d];
    }
  m->r = m->lg[0]; m->N = m->nc[0]; m->left = m->origin[0]; m->right = m->origin[0] + h * m->nc[0];
  const uint32_t n = m->n, npc = ipow_u(n, dim);
  m->npc = npc;
  orc_shape_1d(p, m->sv, m->sg, m->xn, m->xq, m->wq);

  /* deal.II order after global refinement: children of cell k are 2^dim*k.., child index = x + 2y + 4z
   * => Morton order, x = LSB; directions with fewer cells run out of bits first */
  m->cxyz = (uint32_t *)calloc((size_t)m->n_cells * 3, sizeof(uint32_t));
  for (uint32_t c = 0; c < m->n_cells; ++c)
    {
      uint32_t x[3] = {0, 0, 0}; int pos = 0;
      for (int b = 0; b < 11; ++b)
        for (int d = 0; d < dim; ++d)
          if (b < m->lg[d]) { x[d] |= ((c >> pos) & 1u) << b; ++pos; }
      for (int d = 0; d < 3; ++d) m->cxyz[3 * (size_t)c + d] = x[d];
    }

  /* first-touch numbering: cells in order; per cell hierarchic order
   * (vertices, lines, quads, hex). DoFs identified by lattice coordinate. */
  uint32_t *h2l = (uint32_t *)malloc(npc * sizeof(uint32_t));
  orc_hier_to_lex(dim, p, h2l);
  m->lex2hier = (uint32_t *)malloc(npc * sizeof(uint32_t));
  for (uint32_t hI = 0; hI < npc; ++hI) m->lex2hier[h2l[hI]] = hI;
  const uint32_t M[3] = {p * m->nc[0] + 1, p * m->nc[1] + 1, dim == 3 ? p * m->nc[2] + 1 : 1};
  const size_t nlat = (size_t)M[0] * M[1] * M[2];
  uint32_t *lat = (uint32_t *)malloc(nlat * sizeof(uint32_t));
  memset(lat, 0xff, nlat * sizeof(uint32_t));
  m->l2g = (uint32_t *)malloc((size_t)m->n_cells * npc * sizeof(uint32_t));
  uint32_t next = 0;
  for (uint32_t c = 0; c < m->n_cells; ++c)
    {
      const uint32_t *cx = &m->cxyz[3 * (size_t)c];
      for (uint32_t hI = 0; hI < npc; ++hI)
        {
          const uint32_t li = h2l[hI];
          const uint32_t i = li % n, j = (li / n) % n, k = (dim == 3) ? li / (n * n) : 0;
          const size_t   a = (size_t)(cx[0] * p + i) + (size_t)M[0] * ((cx[1] * p + j) + (size_t)M[1] * (dim == 3 ? cx[2] * p + k : 0));
          if (lat[a] == 0xffffffffu) lat[a] = next++;
          m->l2g[(size_t)c * npc + li] = lat[a];
        }
    }
  m->n_dofs = next;
  free(h2l);

  m->dof_lattice = (uint32_t *)malloc((size_t)m->n_dofs * 3 * sizeof(uint32_t));
  m->is_constrained = (uint8_t *)calloc(m->n_dofs, 1);
  for (uint32_t z = 0; z < M[2]; ++z)
    for (uint32_t y = 0; y < M[1]; ++y)
      for (uint32_t x = 0; x < M[0]; ++x)
        {
          const uint32_t g = lat[(size_t)x + (size_t)M[0] * (y + (size_t)M[1] * z)];
          m->dof_lattice[3 * (size_t)g + 0] = x; m->dof_lattice[3 * (size_t)g + 1] = y; m->dof_lattice[3 * (size_t)g + 2] = z;
          /* interpolate_boundary_values(dof_handler,0,ZeroFunction): every DoF on a Dirichlet face */
          const uint32_t X[3] = {x, y, z};
          int onb = 0;
          for (int d = 0; d < dim; ++d)
            {
              if (X[d] == 0 && ((dirichlet_faces >> (2 * d)) & 1u)) onb = 1;
              if (X[d] == M[d] - 1 && ((dirichlet_faces >> (2 * d + 1)) & 1u)) onb = 1;
            }
          m->is_constrained[g] = (uint8_t)onb;
        }
  free(lat);
  /* ConstraintHandlerGpu::reinit: ascending list of constrained indices
   * (constraint_handler_gpu.cu:77-83) */
  uint32_t nc = 0;
  for (uint32_t g = 0; g < m->n_dofs; ++g) nc += m->is_constrained[g];
  m->n_constrained = nc;
  m->constrained = (uint32_t *)malloc((nc ? nc : 1) * sizeof(uint32_t));
  nc = 0;
  for (uint32_t g = 0; g < m->n_dofs; ++g) if (m->is_constrained[g]) m->constrained[nc++] = g;

  /* coefficient at quadrature points (laplace_operator_gpu.h:191-211) */
  m->coef = (double *)malloc((size_t)m->n_cells * npc * sizeof(double));
  for (uint32_t c = 0; c < m->n_cells; ++c)
    {
      const uint32_t *cx = &m->cxyz[3 * (size_t)c];
      for (uint32_t q = 0; q < npc; ++q)
        {
          const uint32_t qi[3] = {q % n, (q / n) % n, (dim == 3) ? q / (n * n) : 0};
          double x[3];
          for (int d = 0; d < dim; ++d) x[d] = m->origin[d] + m->h * (cx[d] + m->xq[qi[d]]);
          m->coef[(size_t)c * npc + q] = coefficient_value(x, dim);
        }
    }

  /* parity coloring (2^dim colors) for the threaded CPU baseline */
  m->color_cells = (uint32_t *)malloc(m->n_cells * sizeof(uint32_t));
  {
    const int ncol = 1 << dim; uint32_t cnt[9] = {0};
    for (uint32_t c = 0; c < m->n_cells; ++c)
      {
        const uint32_t *cx = &m->cxyz[3 * (size_t)c];
        cnt[(cx[0] & 1) + 2 * (cx[1] & 1) + 4 * (cx[2] & 1) + 1]++;
      }
    m->color_off[0] = 0;
    for (int k = 0; k < ncol; ++k) m->color_off[k + 1] = m->color_off[k] + cnt[k + 1];
    uint32_t pos[8]; for (int k = 0; k < ncol; ++k) pos[k] = m->color_off[k];
    for (uint32_t c = 0; c < m->n_cells; ++c)
      {
        const uint32_t *cx = &m->cxyz[3 * (size_t)c];
        m->color_cells[pos[(cx[0] & 1) + 2 * (cx[1] & 1) + 4 * (cx[2] & 1)]++] = c;
      }
  }
  return m;
}

/* hyper_cube(left,right)^dim, refine_global(r), FE_Q(p), Dirichlet on the whole
 * boundary (bmop.cu:111-132), coefficient at Gauss(p+1) points. */
orc_mesh *orc_create(int dim, int p, int r, double left, double right)
{
  if (r < 0 || r > 10) return NULL;
  const int    lg[3] = {r, r, r};
  const double origin[3] = {left, left, left};
  orc_mesh *m = orc_create_box(dim, p, lg, origin, (right - left) / (double)(1u << r), 0x3f);
  if (m) { m->left = left; m->right = right; }
  return m;
}

/* replace the coefficient by a constant (used by the analytic identities) */
void orc_set_constant_coefficient(orc_mesh *m, double a)
{
  for (size_t i = 0; i < (size_t)m->n_cells * m->npc; ++i) m->coef[i] = a;
}
/* drop all constraints (used by the analytic identities) */
void orc_clear_constraints(orc_mesh *m)
{
  memset(m->is_constrained, 0, m->n_dofs); m->n_constrained = 0;
}

uint32_t orc_n_cells(const orc_mesh *m) { return m->n_cells; }
uint32_t orc_n_dofs(const orc_mesh *m) { return m->n_dofs; }
uint32_t orc_dofs_per_cell(const orc_mesh *m) { return m->npc; }
uint32_t orc_n_constrained(const orc_mesh *m) { return m->n_constrained; }
const uint32_t *orc_loc2glob(const orc_mesh *m) { return m->l2g; }
const uint32_t *orc_constrained(const orc_mesh *m) { return m->constrained; }
const uint32_t *orc_dof_lattice(const orc_mesh *m) { return m->dof_lattice; }
const uint32_t *orc_cell_coords(const orc_mesh *m) { return m->cxyz; }
const double *orc_coefficient(const orc_mesh *m) { return m->coef; }
const double *orc_shape_values(const orc_mesh *m) { return m->sv; }
const double *orc_shape_gradients(const orc_mesh *m) { return m->sg; }
const uint32_t *orc_lex2hier(const orc_mesh *m) { return m->lex2hier; }

/* ------------------------------------------------------------------------ */
/* the cell kernel (fee_gpu.cuh + tensor_ops.cuh restated, scalar)           */
/* ------------------------------------------------------------------------ */

/* out[.., q, ..] = sum_k S(k,q) in[.., k, ..] along direction `dir`.
 * tr=1: S(k,q)=M[k*n+q]  (nodes -> quadrature points, tensor_ops.cuh phi_tr=true)
 * tr=0: S(k,q)=M[q*n+k]  (quadrature points -> nodes, phi_tr=false)            */
static void contract(int dim, int n, int dir, int tr, const double *M, const double *in, double *out)
{
  const int nz = (dim == 3) ? n : 1;
  int stride = 1; for (int d = 0; d < dir; ++d) stride *= n;
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < n; ++y)
      for (int x = 0; x < n; ++x)
        {
          const int idx = x + n * (y + n * z);
          const int c[3] = {x, y, z};
          const int q = c[dir];
          const int base = idx - q * stride;
          double t = 0;
          for (int k = 0; k < n; ++k) t += (tr ? M[k * n + q] : M[q * n + k]) * in[base + k * stride];
          out[idx] = t;
        }
}

/* one cell: v = A_cell u  (LocalOperator::cell_apply, laplace_operator_gpu.h:263-281) */
static void cell_apply(const orc_mesh *m, uint32_t cell, const double *u, double *v)
{
  const int dim = m->dim, n = m->n; const uint32_t npc = m->npc;
  double g[3][ORC_MAXN * ORC_MAXN * ORC_MAXN], t1[ORC_MAXN * ORC_MAXN * ORC_MAXN], t2[ORC_MAXN * ORC_MAXN * ORC_MAXN];
  /* evaluate(false,true): grad_at_quad_pts (tensor_ops.cuh:179-217) */
  for (int d = 0; d < dim; ++d)
    {
      const double *in = u; double *bufs[2] = {t1, t2}; int b = 0;
      for (int dir = 0; dir < dim; ++dir)
        {
          double *out = (dir == dim - 1) ? g[d] : bufs[b];
          contract(dim, n, dir, 1, dir == d ? m->sg : m->sv, in, out);
          in = out; b ^= 1;
        }
    }
  /* quad_operation (laplace_operator_gpu.h:257-260) with get_gradient/submit_gradient
   * (fee_gpu.cuh:219-284): inverse Jacobian of a uniform cube cell is (1/h) I,
   * JxW = h^dim * w_q  (matrix_free_gpu.cu:315-338) */
  const double invJ = 1.0 / m->h;
  double hd = 1.0; for (int d = 0; d < dim; ++d) hd *= m->h;
  for (uint32_t q = 0; q < npc; ++q)
    {
      const int qi[3] = {(int)(q % n), (int)((q / n) % n), dim == 3 ? (int)(q / (n * n)) : 0};
      double w = hd; for (int d = 0; d < dim; ++d) w *= m->wq[qi[d]];
      const double c = m->coef[(size_t)cell * npc + q];
      for (int d = 0; d < dim; ++d)
        {
          const double gx = invJ * g[d][q];         /* get_gradient  */
          g[d][q] = (c * gx) * invJ * w;            /* submit_gradient */
        }
    }
  /* integrate(false,true): quad_int_grad (tensor_ops.cuh:219-261) */
  for (uint32_t i = 0; i < npc; ++i) v[i] = 0;
  for (int d = 0; d < dim; ++d)
    {
      const double *in = g[d]; double *bufs[2] = {t1, t2}; int b = 0;
      for (int dir = 0; dir < dim; ++dir)
        {
          double *out = bufs[b];
          contract(dim, n, dir, 0, dir == d ? m->sg : m->sv, in, out);
          in = out; b ^= 1;
        }
      for (uint32_t i = 0; i < npc; ++i) v[i] += in[i];
    }
}

/* dst += A src with identity on constrained rows
 * (laplace_operator_gpu.h:286-303 == laplace_operator_cpu.cc:180-211):
 * constrained src entries read as 0, cells never write constrained rows,
 * then dst[c] = dst_old[c] + src[c]. */
void orc_vmult_add(const orc_mesh *m, double *dst, const double *src)
{
  const uint32_t npc = m->npc;
  double u[ORC_MAXN * ORC_MAXN * ORC_MAXN], v[ORC_MAXN * ORC_MAXN * ORC_MAXN];
  for (uint32_t c = 0; c < m->n_cells; ++c)
    {
      const uint32_t *row = &m->l2g[(size_t)c * npc];
      for (uint32_t i = 0; i < npc; ++i) u[i] = m->is_constrained[row[i]] ? 0.0 : src[row[i]];
      cell_apply(m, c, u, v);
      for (uint32_t i = 0; i < npc; ++i) if (!m->is_constrained[row[i]]) dst[row[i]] += v[i];
    }
  for (uint32_t g = 0; g < m->n_dofs; ++g) if (m->is_constrained[g]) dst[g] += src[g];
}

/* cell loop over the cells [cell_begin, cell_end) only, no constrained-row identity: the partial sums one
 * partition of the mesh contributes (used by the multi-GPU partition tests) */
void orc_cell_loop_range(const orc_mesh *m, double *dst, const double *src, uint32_t cell_begin, uint32_t cell_end)
{
  const uint32_t npc = m->npc;
  double u[ORC_MAXN * ORC_MAXN * ORC_MAXN], v[ORC_MAXN * ORC_MAXN * ORC_MAXN];
  for (uint32_t c = cell_begin; c < cell_end && c < m->n_cells; ++c)
    {
      const uint32_t *row = &m->l2g[(size_t)c * npc];
      for (uint32_t i = 0; i < npc; ++i) u[i] = m->is_constrained[row[i]] ? 0.0 : src[row[i]];
      cell_apply(m, c, u, v);
      for (uint32_t i = 0; i < npc; ++i) if (!m->is_constrained[row[i]]) dst[row[i]] += v[i];
    }
}

/* vmult: dst = 0; vmult_add  (laplace_operator_gpu.h:216-223) */
void orc_vmult(const orc_mesh *m, double *dst, const double *src)
{
  memset(dst, 0, (size_t)m->n_dofs * sizeof(double));
  orc_vmult_add(m, dst, src);
}

/* bmop loop: dst=0.1; repeat k times {swap(dst,src); vmult(dst,src)} (bmop.cu:135-153).
 * On return `out` holds the final dst. */
void orc_bmop(const orc_mesh *m, int k, double init, double *out)
{
  double *a = (double *)malloc((size_t)m->n_dofs * sizeof(double));
  double *b = (double *)malloc((size_t)m->n_dofs * sizeof(double));
  double *dst = a, *src = b;
  for (uint32_t i = 0; i < m->n_dofs; ++i) dst[i] = init;
  for (int it = 0; it < k; ++it)
    {
      double *t = dst; dst = src; src = t;
      orc_vmult(m, dst, src);
    }
  memcpy(out, dst, (size_t)m->n_dofs * sizeof(double));
  free(a); free(b);
}

/* compute_diagonal (laplace_operator_gpu.h:355-421): per cell apply to every
 * local unit vector, keep entry i; scatter-add; constrained -> 1; invert. */
void orc_inverse_diagonal(const orc_mesh *m, double *inv_diag)
{
  const uint32_t npc = m->npc;
  double u[ORC_MAXN * ORC_MAXN * ORC_MAXN], v[ORC_MAXN * ORC_MAXN * ORC_MAXN];
  memset(inv_diag, 0, (size_t)m->n_dofs * sizeof(double));
  for (uint32_t c = 0; c < m->n_cells; ++c)
    {
      const uint32_t *row = &m->l2g[(size_t)c * npc];
      for (uint32_t i = 0; i < npc; ++i)
        {
          for (uint32_t j = 0; j < npc; ++j) u[j] = 0;
          u[i] = 1.0;
          cell_apply(m, c, u, v);
          inv_diag[row[i]] += v[i];
        }
    }
  for (uint32_t g = 0; g < m->n_dofs; ++g)
    {
      if (m->is_constrained[g]) inv_diag[g] = 1.0;
      inv_diag[g] = 1.0 / inv_diag[g];
    }
}

/* ------------------------------------------------------------------------ */
/* independent assembled operator (test_laplace_op.cu:50-120):               */
/* dense cell matrices from explicit shape gradients (no sum factorisation), */
/* assembled, constrained rows/cols replaced by identity; then y = K x.      */
/* Stored as a dense n_dofs x n_dofs matrix: small meshes only.              */
/* ------------------------------------------------------------------------ */
int orc_assemble_dense(const orc_mesh *m, double *K /* [n_dofs*n_dofs], row-major */)
{
  const int dim = m->dim, n = m->n; const uint32_t npc = m->npc, nd = m->n_dofs;
  if ((size_t)nd * nd > ((size_t)1 << 28)) return -1;
  memset(K, 0, (size_t)nd * nd * sizeof(double));
  double *G = (double *)malloc((size_t)npc * npc * 3 * sizeof(double)); /* G[d][i][q] real-space gradient */
  const double invJ = 1.0 / m->h;
  double hd = 1.0; for (int d = 0; d < dim; ++d) hd *= m->h;
  for (uint32_t i = 0; i < npc; ++i)
    for (uint32_t q = 0; q < npc; ++q)
      {
        const int ii[3] = {(int)(i % n), (int)((i / n) % n), dim == 3 ? (int)(i / (n * n)) : 0};
        const int qi[3] = {(int)(q % n), (int)((q / n) % n), dim == 3 ? (int)(q / (n * n)) : 0};
        for (int d = 0; d < dim; ++d)
          {
            double t = invJ;
            for (int e = 0; e < dim; ++e) t *= (e == d ? m->sg : m->sv)[ii[e] * n + qi[e]];
            G[((size_t)d * npc + i) * npc + q] = t;
          }
      }
  double *Kc = (double *)malloc((size_t)npc * npc * sizeof(double));
  for (uint32_t c = 0; c < m->n_cells; ++c)
    {
      memset(Kc, 0, (size_t)npc * npc * sizeof(double));
      for (uint32_t q = 0; q < npc; ++q)
        {
          const int qi[3] = {(int)(q % n), (int)((q / n) % n), dim == 3 ? (int)(q / (n * n)) : 0};
          double w = hd; for (int d = 0; d < dim; ++d) w *= m->wq[qi[d]];
          const double cw = m->coef[(size_t)c * npc + q] * w;
          for (uint32_t i = 0; i < npc; ++i)
            for (uint32_t j = 0; j < npc; ++j)
              {
                double s = 0;
                for (int d = 0; d < dim; ++d) s += G[((size_t)d * npc + i) * npc + q] * G[((size_t)d * npc + j) * npc + q];
                Kc[(size_t)i * npc + j] += s * cw;
              }
        }
      const uint32_t *row = &m->l2g[(size_t)c * npc];
      for (uint32_t i = 0; i < npc; ++i)
        for (uint32_t j = 0; j < npc; ++j) K[(size_t)row[i] * nd + row[j]] += Kc[(size_t)i * npc + j];
    }
  for (uint32_t g = 0; g < nd; ++g)
    if (m->is_constrained[g])
      {
        for (uint32_t j = 0; j < nd; ++j) { K[(size_t)g * nd + j] = 0; K[(size_t)j * nd + g] = 0; }
        K[(size_t)g * nd + g] = 1.0;
      }
  free(G); free(Kc);
  return 0;
}

void orc_dense_vmult(uint32_t nd, const double *K, double *y, const double *x)
{
  for (uint32_t i = 0; i < nd; ++i)
    {
      double s = 0; const double *r = &K[(size_t)i * nd];
      for (uint32_t j = 0; j < nd; ++j) s += r[j] * x[j];
      y[i] = s;
    }
}

/* ------------------------------------------------------------------------ */
/* deterministic test vector: u_i = splitmix64(seed, i) in [0,1)             */
/*                                                     */
/* ------------------------------------------------------------------------ */
void orc_fill_sm64(uint64_t seed, uint32_t nvals, double *u)
{
  for (uint32_t i = 0; i < nvals; ++i)
    {
      uint64_t z = seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      u[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
}

/* ------------------------------------------------------------------------ */
/* threaded baseline: same arithmetic, OpenMP over conflict-free colors      */
/* (the reference CPU path uses TBB partition_color,                         */
/*  laplace_operator_cpu.cc:51-52). Used only as the timed CPU baseline.     */
/* ------------------------------------------------------------------------ */
void orc_vmult_omp(const orc_mesh *m, double *dst, const double *src)
{
  const uint32_t npc = m->npc; const int ncol = 1 << m->dim;
#pragma omp parallel
  {
#pragma omp for schedule(static)
    for (uint32_t g = 0; g < m->n_dofs; ++g) dst[g] = m->is_constrained[g] ? src[g] : 0.0;
    for (int col = 0; col < ncol; ++col)
      {
#pragma omp for schedule(static)
        for (uint32_t ci = m->color_off[col]; ci < m->color_off[col + 1]; ++ci)
          {
            double u[ORC_MAXN * ORC_MAXN * ORC_MAXN], v[ORC_MAXN * ORC_MAXN * ORC_MAXN];
            const uint32_t c = m->color_cells[ci];
            const uint32_t *row = &m->l2g[(size_t)c * npc];
            for (uint32_t i = 0; i < npc; ++i) u[i] = m->is_constrained[row[i]] ? 0.0 : src[row[i]];
            cell_apply(m, c, u, v);
            for (uint32_t i = 0; i < npc; ++i) if (!m->is_constrained[row[i]]) dst[row[i]] += v[i];
          }
      }
  }
}

/* ------------------------------------------------------------------------ */
/* timed CPU baseline: what deal.II's MatrixFree/FEEvaluation does on the    */
/* host (laplace_operator_cpu.cc:125-143): sum factorisation in the          */
/* collocation form, 8 cells per SIMD batch (VectorizedArray<double> on      */
/* AVX-512), OpenMP over conflict-free colors (partition_color, :51-52).     */
/* Same bilinear form as orc_vmult, different summation order.               */
/* ------------------------------------------------------------------------ */
#define ORC_LANES 8
typedef double lane_t[ORC_LANES] __attribute__((aligned(64)));

static inline __attribute__((always_inline)) void
batch_contract(const int n, const int dim, const int dir, const int tr, const double *M, lane_t *in, lane_t *out)
{
  const int nz = dim == 3 ? n : 1;
  int stride = 1; for (int d = 0; d < dir; ++d) stride *= n;
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < n; ++y)
      for (int x = 0; x < n; ++x)
        {
          const int c[3] = {x, y, z};
          if (c[dir] != 0) continue;
          const int base = x + n * (y + n * z);
          for (int q = 0; q < n; ++q)
            {
              double acc[ORC_LANES];
              const double m0 = tr ? M[q * n + 0] : M[0 * n + q];
              for (int l = 0; l < ORC_LANES; ++l) acc[l] = m0 * in[base][l];
              for (int k = 1; k < n; ++k)
                {
                  const double mk = tr ? M[q * n + k] : M[k * n + q];
                  for (int l = 0; l < ORC_LANES; ++l) acc[l] += mk * in[base + k * stride][l];
                }
              for (int l = 0; l < ORC_LANES; ++l) out[base + q * stride][l] = acc[l];
            }
        }
}

static inline __attribute__((always_inline)) void
batch_apply(const orc_mesh *m, const int n, const int dim, const double *Dc, const double *wfac, const uint32_t *cells, const int nb,
            const double *src, double *dst)
{
  const int npc = dim == 3 ? n * n * n : n * n;
  lane_t u[ORC_MAXN * ORC_MAXN * ORC_MAXN], t[ORC_MAXN * ORC_MAXN * ORC_MAXN], g[ORC_MAXN * ORC_MAXN * ORC_MAXN], R[ORC_MAXN * ORC_MAXN * ORC_MAXN];
  for (int e = 0; e < npc; ++e)
    for (int l = 0; l < ORC_LANES; ++l)
      {
        double v = 0.0;
        if (l < nb) { const uint32_t gi = m->l2g[(size_t)cells[l] * npc + e]; v = m->is_constrained[gi] ? 0.0 : src[gi]; }
        u[e][l] = v;
      }
  /* values at the quadrature points */
  batch_contract(n, dim, 0, 0, m->sv, u, t);
  batch_contract(n, dim, 1, 0, m->sv, t, u);
  if (dim == 3) { batch_contract(n, dim, 2, 0, m->sv, u, t); } else { for (int e = 0; e < npc; ++e) for (int l = 0; l < ORC_LANES; ++l) t[e][l] = u[e][l]; }
  /* R = sum_d D_d^T ( a * JxW * J^-2 .* D_d G ) */
  for (int e = 0; e < npc; ++e) for (int l = 0; l < ORC_LANES; ++l) R[e][l] = 0.0;
  for (int d = 0; d < dim; ++d)
    {
      batch_contract(n, dim, d, 0, Dc, t, g);
      for (int e = 0; e < npc; ++e)
        for (int l = 0; l < ORC_LANES; ++l)
          g[e][l] *= (l < nb ? m->coef[(size_t)cells[l] * npc + e] : 0.0) * wfac[e];
      batch_contract(n, dim, d, 1, Dc, g, u);
      for (int e = 0; e < npc; ++e) for (int l = 0; l < ORC_LANES; ++l) R[e][l] += u[e][l];
    }
  batch_contract(n, dim, 0, 1, m->sv, R, u);
  batch_contract(n, dim, 1, 1, m->sv, u, t);
  if (dim == 3) batch_contract(n, dim, 2, 1, m->sv, t, u);
  lane_t *res = dim == 3 ? u : t;
  for (int l = 0; l < nb; ++l)
    for (int e = 0; e < npc; ++e)
      {
        const uint32_t gi = m->l2g[(size_t)cells[l] * npc + e];
        if (!m->is_constrained[gi]) dst[gi] += res[e][l];
      }
}

/* AVX-512 clone where the host has it (the reference's VectorizedArray<double> width), x86-64-v3 otherwise */
__attribute__((target_clones("avx512f", "default")))
static void batch_apply_dispatch(const orc_mesh *m, const double *Dc, const double *wfac, const uint32_t *cells, int nb, const double *src, double *dst)
{
  if (m->dim == 3)
    switch (m->n) { case 2: batch_apply(m, 2, 3, Dc, wfac, cells, nb, src, dst); break; case 3: batch_apply(m, 3, 3, Dc, wfac, cells, nb, src, dst); break;
      case 4: batch_apply(m, 4, 3, Dc, wfac, cells, nb, src, dst); break; case 5: batch_apply(m, 5, 3, Dc, wfac, cells, nb, src, dst); break;
      case 6: batch_apply(m, 6, 3, Dc, wfac, cells, nb, src, dst); break; case 7: batch_apply(m, 7, 3, Dc, wfac, cells, nb, src, dst); break;
      case 8: batch_apply(m, 8, 3, Dc, wfac, cells, nb, src, dst); break; default: batch_apply(m, 9, 3, Dc, wfac, cells, nb, src, dst); break; }
  else
    switch (m->n) { case 2: batch_apply(m, 2, 2, Dc, wfac, cells, nb, src, dst); break; case 3: batch_apply(m, 3, 2, Dc, wfac, cells, nb, src, dst); break;
      case 4: batch_apply(m, 4, 2, Dc, wfac, cells, nb, src, dst); break; case 5: batch_apply(m, 5, 2, Dc, wfac, cells, nb, src, dst); break;
      case 6: batch_apply(m, 6, 2, Dc, wfac, cells, nb, src, dst); break; case 7: batch_apply(m, 7, 2, Dc, wfac, cells, nb, src, dst); break;
      case 8: batch_apply(m, 8, 2, Dc, wfac, cells, nb, src, dst); break; default: batch_apply(m, 9, 2, Dc, wfac, cells, nb, src, dst); break; }
}

void orc_vmult_fast(const orc_mesh *m, double *dst, const double *src)
{
  const int n = m->n, dim = m->dim; const uint32_t npc = m->npc; const int ncol = 1 << dim;
  double Dc[ORC_MAXN * ORC_MAXN], dummy[ORC_MAXN * ORC_MAXN], wfac[ORC_MAXN * ORC_MAXN * ORC_MAXN];
  lagrange_eval(n, m->xq, n, m->xq, dummy, Dc); /* collocation derivative: Lagrange basis through the Gauss points */
  double hd = 1.0; for (int d = 0; d < dim; ++d) hd *= m->h;
  for (uint32_t q = 0; q < npc; ++q)
    {
      const int qi[3] = {(int)(q % n), (int)((q / n) % n), dim == 3 ? (int)(q / (n * n)) : 0};
      double w = hd / (m->h * m->h); for (int d = 0; d < dim; ++d) w *= m->wq[qi[d]];
      wfac[q] = w;
    }
#pragma omp parallel
  {
#pragma omp for schedule(static)
    for (uint32_t g = 0; g < m->n_dofs; ++g) dst[g] = m->is_constrained[g] ? src[g] : 0.0;
    for (int col = 0; col < ncol; ++col)
      {
        const uint32_t c0 = m->color_off[col], c1 = m->color_off[col + 1];
        const uint32_t nbatch = (c1 - c0 + ORC_LANES - 1) / ORC_LANES;
#pragma omp for schedule(static)
        for (uint32_t b = 0; b < nbatch; ++b)
          {
            const uint32_t first = c0 + b * ORC_LANES;
            const int nb = (int)((c1 - first) < ORC_LANES ? (c1 - first) : ORC_LANES);
            batch_apply_dispatch(m, Dc, wfac, &m->color_cells[first], nb, src, dst);
          }
      }
  }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the timed CPU arm asks for the host's threads explicitly */
void orc_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
