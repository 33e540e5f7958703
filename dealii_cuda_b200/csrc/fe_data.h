// fe_data.h -- 1-D finite element data the reference takes from deal.II:
// QGauss<1>(p+1), the FE_Q(p) Gauss-Lobatto support points and
// internal::MatrixFreeFunctions::ShapeInfo::{shape_values,shape_gradients}
// (matrix_free_gpu.cu:492-513).  Host only, computed in long double with
// barycentric Lagrange formulas.
#pragma once
#include <cmath>
#include <vector>

namespace mfg {

struct FEData1D
{
  int                 degree = 0, n = 0;
  std::vector<double> nodes;      // GLL support points on [0,1]
  std::vector<double> qpts, qwts; // Gauss-Legendre(n) on [0,1]
  std::vector<double> val;        // val[i*n+q]  = phi_i(x_q)
  std::vector<double> grad;       // grad[i*n+q] = phi_i'(x_q)   (reference-cell derivative)
  std::vector<double> colloc;     // colloc[a*n+q] = l_a'(x_q), l_a = Lagrange basis through the Gauss points
  std::vector<double> hanging;    // hanging[k*n+i] = phi_i(xi_k/2): subface interpolation, first child (hanging_nodes.cuh:580-598)
};

namespace detail {
typedef long double ld;

inline void legendre_pd(int n, ld x, ld &P, ld &dP)
{
  ld a = 1, b = x;
  if (n == 0) { P = 1; dP = 0; return; }
  for (int k = 2; k <= n; ++k) { ld c = ((2 * k - 1) * x * b - (k - 1) * a) / k; a = b; b = c; }
  P = b; dP = n * (a - x * b) / (1 - x * x);
}

// barycentric weights of a node set
inline std::vector<ld> bary_weights(const std::vector<ld> &x)
{
  const int n = (int)x.size();
  std::vector<ld> w(n, 1);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (j != i) w[i] /= (x[i] - x[j]);
  return w;
}

// value and derivative of Lagrange polynomial i (nodes x, bary weights w) at point t
inline void lagrange_at(const std::vector<ld> &x, const std::vector<ld> &w, int i, ld t, ld &v, ld &d)
{
  const int n = (int)x.size();
  // if t coincides with a node use the differentiation-matrix formulas
  for (int k = 0; k < n; ++k)
    if (t == x[k])
      {
        if (k == i)
          {
            v = 1; d = 0;
            for (int j = 0; j < n; ++j) if (j != i) d += 1 / (x[i] - x[j]);
          }
        else { v = 0; d = (w[i] / w[k]) / (x[k] - x[i]); }
        return;
      }
  // general point: l_i(t) = L(t) w_i/(t-x_i),  L = prod (t-x_j)
  ld L = 1; for (int j = 0; j < n; ++j) L *= (t - x[j]);
  ld s = 0; for (int j = 0; j < n; ++j) s += 1 / (t - x[j]);
  v = L * w[i] / (t - x[i]);
  d = v * (s - 1 / (t - x[i]));
}
}  // namespace detail

inline FEData1D make_fe_data(int degree)
{
  using detail::ld;
  const ld pi = 3.141592653589793238462643383279502884L;
  FEData1D fe;
  fe.degree = degree; fe.n = degree + 1;
  const int n = fe.n;
  std::vector<ld> xg(n), wg(n), xl(n);
  // Gauss-Legendre on [-1,1] by Newton on P_n
  for (int i = 0; i < n; ++i)
    {
      ld z = -std::cos(pi * (4 * i + 3) / (4 * n + 2)), P, dP;
      for (int it = 0; it < 60; ++it) { detail::legendre_pd(n, z, P, dP); ld dz = P / dP; z -= dz; if (std::fabs(dz) < 1e-20L) break; }
      detail::legendre_pd(n, z, P, dP);
      xg[i] = z; wg[i] = 2 / ((1 - z * z) * dP * dP);
    }
  // Gauss-Lobatto: end points + roots of P'_{n-1}; Newton with P'' from the Legendre ODE
  xl[0] = -1; xl[n - 1] = 1;
  for (int i = 1; i + 1 < n; ++i)
    {
      const int m = n - 1;
      ld z = -std::cos(pi * i / m), P, dP;
      for (int it = 0; it < 60; ++it)
        {
          detail::legendre_pd(m, z, P, dP);
          ld ddP = (2 * z * dP - m * (m + 1) * P) / (1 - z * z);
          ld dz = dP / ddP; z -= dz; if (std::fabs(dz) < 1e-20L) break;
        }
      xl[i] = z;
    }
  for (int i = 0; i < n / 2; ++i) { ld a = (xl[n - 1 - i] - xl[i]) / 2; xl[i] = -a; xl[n - 1 - i] = a; }
  if (n % 2) xl[n / 2] = 0;
  for (int i = 0; i < n / 2; ++i) { ld a = (xg[n - 1 - i] - xg[i]) / 2; xg[i] = -a; xg[n - 1 - i] = a; ld w = (wg[i] + wg[n - 1 - i]) / 2; wg[i] = wg[n - 1 - i] = w; }
  if (n % 2) xg[n / 2] = 0;
  // map to [0,1]
  std::vector<ld> xn01(n), xq01(n);
  for (int i = 0; i < n; ++i) { xn01[i] = (xl[i] + 1) / 2; xq01[i] = (xg[i] + 1) / 2; }
  fe.nodes.resize(n); fe.qpts.resize(n); fe.qwts.resize(n);
  for (int i = 0; i < n; ++i) { fe.nodes[i] = (double)xn01[i]; fe.qpts[i] = (double)xq01[i]; fe.qwts[i] = (double)(wg[i] / 2); }
  fe.val.resize(n * n); fe.grad.resize(n * n); fe.colloc.resize(n * n); fe.hanging.resize(n * n);
  const std::vector<ld> wn = detail::bary_weights(xn01), wq = detail::bary_weights(xq01);
  for (int i = 0; i < n; ++i)
    for (int q = 0; q < n; ++q)
      {
        ld v, d;
        detail::lagrange_at(xn01, wn, i, xq01[q], v, d);
        fe.val[i * n + q] = (double)v; fe.grad[i * n + q] = (double)d;
        detail::lagrange_at(xq01, wq, i, xq01[q], v, d);
        fe.colloc[i * n + q] = (double)d;
        detail::lagrange_at(xn01, wn, i, xn01[q] / 2, v, d);  // q plays the role of the fine node k
        fe.hanging[q * n + i] = (double)v;
      }
  return fe;
}

}  // namespace mfg
