// gecut_integrate.cu -- K4: cut-cell quadrature (measures and Laplacian element matrices), FP64 on CUDA cores.
//
// Reference behaviour replaced: Gridap's `Measure(trian,2)` / `∫(f)dΩ` evaluated on the triangulations the reference
// builds from SubCellData / SubFacetData (src/Interfaces/SubCellTriangulations.jl:26-97, SubFacetTriangulations.jl:30-86;
// call sites examples/PoissonCSGCutFEM/PoissonCSGCutFEM.jl:47-51,63-70, test/InterfacesTests/EmbeddedDiscretizationsTests.jl:48-56):
// degree-2 reference rule of the sub-simplex mapped through its P1 map, weight w_q*|det J_s| (facets: sqrt(det JtJ)),
// background shape functions evaluated at the glued reference point eta_q = sum_i N_i(xi_q) R_i, gradients pushed with the
// background cell Jacobian; contributions of the subcells of a cut cell are summed in subcell order (move_contributions,
// SubCellTriangulations.jl:63-72).  No step is a dense contraction, so tensor cores are deliberately unused.
#include <algorithm>
#include <cmath>

#include "gecut_internal.cuh"

namespace {

__host__ __device__ inline double det2x2(const double J[3][3]) { return J[0][0] * J[1][1] - J[0][1] * J[1][0]; }
__host__ __device__ inline double det3x3(const double J[3][3]) {
  return J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
         (J[0][0] * J[1][2] * J[2][1] + J[0][1] * J[1][0] * J[2][2] + J[0][2] * J[1][1] * J[2][0]);
}

// |det J| of a D-simplex / sqrt(det JtJ) of a (D-1)-simplex whose vertices are p[0..ds] (D doubles each)
template <int D, int DS>
__device__ __forceinline__ double simplex_meas(const double (*p)[3]) {
  if (DS == D) {
    double J[3][3];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int d = 0; d < D; ++d) J[i][d] = p[i + 1][d] - p[0][d];
    return fabs(D == 2 ? det2x2(J) : det3x3(J));
  }
  if (D == 2) {
    double a = p[1][0] - p[0][0], b = p[1][1] - p[0][1];
    return sqrt(a * a + b * b);
  }
  double a[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
  double b[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
  double n1 = a[1] * b[2] - a[2] * b[1], n2 = a[2] * b[0] - a[0] * b[2], n3 = a[0] * b[1] - a[1] * b[0];
  return sqrt(n1 * n1 + n2 * n2 + n3 * n3);
}

// sum_q w_q * meas with the degree-2 rule of the DS-simplex (weights 1/2 x2, 1/6 x3, 1/24 x4)
template <int DS>
__device__ __forceinline__ double integrate_one(double meas) {
  const double w = DS == 1 ? 0.5 : (DS == 2 ? 1.0 / 6 : 1.0 / 24);
  const int nq = DS + 1;
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < nq; ++q) s += w * meas;
  return s;
}

template <int D, int DS>
__global__ void k_measures(const int32_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ c2p,
                           const double* __restrict__ X, double* __restrict__ values) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t e = (int64_t)ids[i] - 1;
  double p[DS + 1][3];
#pragma unroll
  for (int v = 0; v <= DS; ++v) {
    int64_t pid = (int64_t)c2p[e * (DS + 1) + v] - 1;
#pragma unroll
    for (int d = 0; d < D; ++d) p[v][d] = X[pid * D + d];
  }
  values[i] = integrate_one<DS>(simplex_meas<D, DS>(p));
}

// deterministic two-level sum
__global__ void k_block_sums(const double* __restrict__ v, int64_t n, double* __restrict__ partial) {
  __shared__ double sh[256];
  double s = 0.0;
  int64_t per = (n + gridDim.x - 1) / gridDim.x;
  int64_t lo = per * blockIdx.x, hi = lo + per < n ? lo + per : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s += v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void k_final_sum(const double* __restrict__ partial, int nb, double* __restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// ---- bg cell helpers shared by host (uncut cells) and device (cut cells) ---------------------------
__host__ __device__ inline double hd_coord(const GcGrid& g, int d, int i) { return g.pmin[d] + (double)i * g.dx[d]; }

// vertex coordinates of bg cell (cube ijk, type typ): n-cube vertices x fastest, or the typ-th simplex of simplexify(p)
__host__ __device__ inline void bgcell_vertices(const GcGrid& g, const int* ijk, int typ, const uint8_t* simp_raw,
                                                double x[8][3]) {
  const int D = g.D;
  for (int m = 0; m < g.nv; ++m) {
    int lvid = g.simplexified ? simp_raw[(D - 2) * 24 + typ * 4 + m] : m;
    for (int d = 0; d < D; ++d) x[m][d] = hd_coord(g, d, ijk[d] + ((lvid >> d) & 1));
  }
}

// inv(Jt) of the bg cell map (constant per cell): grad_x phi = invJt * grad_xi phi
__host__ __device__ inline void bgcell_inv_jt(const GcGrid& g, const double x[8][3], double invJt[3][3]) {
  const int D = g.D;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) invJt[a][b] = 0.0;
  if (!g.simplexified) {
    for (int d = 0; d < D; ++d) invJt[d][d] = 1.0 / g.dx[d];
    return;
  }
  double Jt[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < D; ++i)
    for (int d = 0; d < D; ++d) Jt[i][d] = x[i + 1][d] - x[0][d];
  if (D == 2) {
    double det = det2x2(Jt);
    invJt[0][0] = Jt[1][1] / det;
    invJt[0][1] = -Jt[0][1] / det;
    invJt[1][0] = -Jt[1][0] / det;
    invJt[1][1] = Jt[0][0] / det;
  } else {
    double det = det3x3(Jt);
    invJt[0][0] = (Jt[1][1] * Jt[2][2] - Jt[1][2] * Jt[2][1]) / det;
    invJt[0][1] = (Jt[0][2] * Jt[2][1] - Jt[0][1] * Jt[2][2]) / det;
    invJt[0][2] = (Jt[0][1] * Jt[1][2] - Jt[0][2] * Jt[1][1]) / det;
    invJt[1][0] = (Jt[1][2] * Jt[2][0] - Jt[1][0] * Jt[2][2]) / det;
    invJt[1][1] = (Jt[0][0] * Jt[2][2] - Jt[0][2] * Jt[2][0]) / det;
    invJt[1][2] = (Jt[0][2] * Jt[1][0] - Jt[0][0] * Jt[1][2]) / det;
    invJt[2][0] = (Jt[1][0] * Jt[2][1] - Jt[1][1] * Jt[2][0]) / det;
    invJt[2][1] = (Jt[0][1] * Jt[2][0] - Jt[0][0] * Jt[2][1]) / det;
    invJt[2][2] = (Jt[0][0] * Jt[1][1] - Jt[0][1] * Jt[1][0]) / det;
  }
}

// reference gradients of the bg shape functions at eta (Q1 on n-cubes, P1 on simplices), vertex order x fastest
__host__ __device__ inline void bg_ref_grads(int D, bool simplex, const double* eta, double gr[8][3]) {
  if (simplex) {
    for (int d = 0; d < D; ++d) gr[0][d] = -1.0;
    for (int i = 0; i < D; ++i)
      for (int d = 0; d < D; ++d) gr[i + 1][d] = (i == d) ? 1.0 : 0.0;
    return;
  }
  const int nv = 1 << D;
  for (int a = 0; a < nv; ++a)
    for (int d = 0; d < D; ++d) {
      double v = 1.0;
      for (int e = 0; e < D; ++e) {
        int bit = (a >> e) & 1;
        if (e == d)
          v *= bit ? 1.0 : -1.0;
        else
          v *= bit ? eta[e] : (1.0 - eta[e]);
      }
      gr[a][d] = v;
    }
}

struct QuadRule {
  int n;
  double xi[8][3];
  double w[8];
};

__host__ __device__ inline QuadRule simplex_rule_deg2(int ds) {
  QuadRule q;
  q.n = ds + 1;
  for (int i = 0; i < 8; ++i) {
    q.w[i] = 0;
    for (int d = 0; d < 3; ++d) q.xi[i][d] = 0;
  }
  if (ds == 1) {
    double a = 0.5 - 0.5 / sqrt(3.0), b = 0.5 + 0.5 / sqrt(3.0);
    q.xi[0][0] = a;
    q.xi[1][0] = b;
    q.w[0] = q.w[1] = 0.5;
  } else if (ds == 2) {
    const double p[3][2] = {{1.0 / 6, 1.0 / 6}, {2.0 / 3, 1.0 / 6}, {1.0 / 6, 2.0 / 3}};
    for (int i = 0; i < 3; ++i) {
      q.xi[i][0] = p[i][0];
      q.xi[i][1] = p[i][1];
      q.w[i] = 1.0 / 6;
    }
  } else {
    double a = (5.0 - sqrt(5.0)) / 20.0, b = (5.0 + 3.0 * sqrt(5.0)) / 20.0;
    const double p[4][3] = {{a, a, a}, {b, a, a}, {a, b, a}, {a, a, b}};
    for (int i = 0; i < 4; ++i) {
      for (int d = 0; d < 3; ++d) q.xi[i][d] = p[i][d];
      q.w[i] = 1.0 / 24;
    }
  }
  return q;
}

inline QuadRule ncube_rule_deg2(int D) {
  QuadRule q;
  double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  q.n = 1 << D;
  for (int i = 0; i < q.n; ++i) {
    double w = 1.0;
    for (int d = 0; d < 3; ++d) q.xi[i][d] = 0;
    for (int d = 0; d < D; ++d) {
      q.xi[i][d] = g[(i >> d) & 1];
      w *= 0.5;
    }
    q.w[i] = w;
  }
  return q;
}

// One warp per cut bg cell: lanes own entries (a,b) of the nv x nv matrix and loop over the selected subcells of the
// cell (a contiguous run of the selection, found by binary search on the bg cell id) and the quadrature points.
template <int D>
__global__ void __launch_bounds__(128)
k_cutcell_laplacian(const GcGrid g, const GcLayout lay, const int32_t* __restrict__ sel_ids, int64_t nsel,
                    int64_t ncut, const uint8_t* __restrict__ simp_raw, double* __restrict__ K) {
  const int lane = threadIdx.x & 31;
  const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= ncut) return;
  const int nv = g.nv, nn = nv * nv;
  const int32_t bg = lay.cutcells[k];
  // run [lo,hi) of the selection with cell_to_bgcell == bg
  int64_t lo, hi;
  {
    int64_t a = 0, b = nsel;
    while (a < b) {
      int64_t m = (a + b) >> 1;
      if (lay.sc_c2bg[(int64_t)sel_ids[m] - 1] < bg) a = m + 1; else b = m;
    }
    lo = a;
    b = nsel;
    while (a < b) {
      int64_t m = (a + b) >> 1;
      if (lay.sc_c2bg[(int64_t)sel_ids[m] - 1] <= bg) a = m + 1; else b = m;
    }
    hi = a;
  }
  double acc[2] = {0.0, 0.0};
  if (hi > lo) {
    const int64_t cell0 = (int64_t)bg - 1;
    const int64_t cube = cell0 / g.cpc;
    const int typ = (int)(cell0 % g.cpc);
    int ijk[3];
    ijk[0] = (int)(cube % g.n[0]);
    if (D == 3) {
      ijk[1] = (int)((cube / g.n[0]) % g.n[1]);
      ijk[2] = (int)(cube / ((int64_t)g.n[0] * g.n[1])) + g.k0;
    } else {
      ijk[1] = (int)(cube / g.n[0]) + g.k0;
      ijk[2] = 0;
    }
    double xv[8][3], invJt[3][3];
    bgcell_vertices(g, ijk, typ, simp_raw, xv);
    bgcell_inv_jt(g, xv, invJt);
    const QuadRule q = simplex_rule_deg2(D);
    for (int64_t s = lo; s < hi; ++s) {
      const int64_t e = (int64_t)sel_ids[s] - 1;
      double p[D + 1][3], r[D + 1][3];
#pragma unroll
      for (int v = 0; v <= D; ++v) {
        int64_t pid = (int64_t)lay.sc_c2p[e * (D + 1) + v] - 1;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          p[v][d] = lay.sc_X[pid * D + d];
          r[v][d] = lay.sc_R[pid * D + d];
        }
      }
      const double dV = simplex_meas<D, D>(p);
      for (int iq = 0; iq < q.n; ++iq) {
        double N[4];
        double sm = 0;
        for (int d = 0; d < D; ++d) sm += q.xi[iq][d];
        N[0] = 1.0 - sm;
        for (int d = 0; d < D; ++d) N[d + 1] = q.xi[iq][d];
        double eta[3] = {0, 0, 0};
        for (int i = 0; i <= D; ++i)
          for (int d = 0; d < D; ++d) eta[d] += N[i] * r[i][d];
        double gr[8][3], gx[8][3];
        bg_ref_grads(D, g.simplexified != 0, eta, gr);
        for (int a = 0; a < nv; ++a)
          for (int d = 0; d < D; ++d) {
            double v = 0;
            for (int e2 = 0; e2 < D; ++e2) v += invJt[d][e2] * gr[a][e2];
            gx[a][d] = v;
          }
        const double wdV = q.w[iq] * dV;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          int idx = lane + 32 * u;
          if (idx < nn) {
            int a = idx / nv, b = idx % nv;
            double dd = 0;
            for (int d = 0; d < D; ++d) dd += gx[a][d] * gx[b][d];
            acc[u] += wdV * dd;
          }
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int idx = lane + 32 * u;
    if (idx < nn) K[k * nn + idx] = acc[u];
  }
}

}  // namespace

// measure of an uncut bg cell of type `typ` (host; tensor Gauss for n-cubes, degree-2 simplex rule otherwise)
static double host_bgcell_measure(const GcGrid& g, int typ) {
  const int D = g.D;
  if (!g.simplexified) {
    QuadRule q = ncube_rule_deg2(D);
    double detJ = g.dx[0] * g.dx[1] * (D == 3 ? g.dx[2] : 1.0);
    double s = 0;
    for (int i = 0; i < q.n; ++i) s += q.w[i] * std::fabs(detJ);
    return s;
  }
  int ijk[3] = {0, 0, 0};
  ijk[D - 1] = g.k0;
  double x[8][3];
  bgcell_vertices(g, ijk, typ, &GC_SIMPLEXIFY_RAW[0][0][0], x);
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < D; ++i)
    for (int d = 0; d < D; ++d) J[i][d] = x[i + 1][d] - x[0][d];
  double dV = std::fabs(D == 2 ? det2x2(J) : det3x3(J));
  QuadRule q = simplex_rule_deg2(D);
  double s = 0;
  for (int i = 0; i < q.n; ++i) s += q.w[i] * dV;
  return s;
}

static void host_bgcell_laplacian(const GcGrid& g, int typ, double* K) {
  const int D = g.D, nv = g.nv;
  for (int i = 0; i < nv * nv; ++i) K[i] = 0.0;
  int ijk[3] = {0, 0, 0};
  ijk[D - 1] = g.k0;
  double x[8][3], invJt[3][3];
  bgcell_vertices(g, ijk, typ, &GC_SIMPLEXIFY_RAW[0][0][0], x);
  bgcell_inv_jt(g, x, invJt);
  QuadRule q = g.simplexified ? simplex_rule_deg2(D) : ncube_rule_deg2(D);
  double dV;
  if (!g.simplexified) {
    dV = std::fabs(g.dx[0] * g.dx[1] * (D == 3 ? g.dx[2] : 1.0));
  } else {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < D; ++i)
      for (int d = 0; d < D; ++d) J[i][d] = x[i + 1][d] - x[0][d];
    dV = std::fabs(D == 2 ? det2x2(J) : det3x3(J));
  }
  for (int iq = 0; iq < q.n; ++iq) {
    double gr[8][3], gx[8][3];
    bg_ref_grads(D, g.simplexified != 0, q.xi[iq], gr);
    for (int a = 0; a < nv; ++a)
      for (int d = 0; d < D; ++d) {
        double v = 0;
        for (int e = 0; e < D; ++e) v += invJt[d][e] * gr[a][e];
        gx[a][d] = v;
      }
    double wdV = q.w[iq] * dV;
    for (int a = 0; a < nv; ++a)
      for (int b = 0; b < nv; ++b) {
        double dd = 0;
        for (int d = 0; d < D; ++d) dd += gx[a][d] * gx[b][d];
        K[a * nv + b] += wdV * dd;
      }
  }
}

int gc_integrate_measure(gecut_ctx* ctx, const gecut_selection* sel, double* values_dev, double* total_host) {
  const gecut_result* res = sel->res;
  const GcGrid& g = res->grid;
  const GcLayout& lay = res->lay;
  const int64_t n = sel->n_sub;
  double total = 0.0;
  const bool use_sub = sel->kind == GECUT_SEL_PHYSICAL || sel->kind == GECUT_SEL_CUTONLY || sel->kind == GECUT_SEL_BOUNDARY;
  if (use_sub && n > 0) {
    double* vals = values_dev;
    if (!vals) GC_CHECK(gc_malloc(ctx, (void**)&vals, sizeof(double) * n));
    const int T = 256;
    const unsigned nb = (unsigned)gc_div_up(n, T);
    if (sel->kind == GECUT_SEL_BOUNDARY) {
      if (g.D == 2)
        GC_LAUNCH(ctx, "k_measures", k_measures<2, 1><<<nb, T, 0, ctx->stream>>>(sel->sub_ids, n, lay.sf_c2p, lay.sf_X, vals));
      else
        GC_LAUNCH(ctx, "k_measures", k_measures<3, 2><<<nb, T, 0, ctx->stream>>>(sel->sub_ids, n, lay.sf_c2p, lay.sf_X, vals));
    } else {
      if (g.D == 2)
        GC_LAUNCH(ctx, "k_measures", k_measures<2, 2><<<nb, T, 0, ctx->stream>>>(sel->sub_ids, n, lay.sc_c2p, lay.sc_X, vals));
      else
        GC_LAUNCH(ctx, "k_measures", k_measures<3, 3><<<nb, T, 0, ctx->stream>>>(sel->sub_ids, n, lay.sc_c2p, lay.sc_X, vals));
    }
    const int NB = (int)std::min<int64_t>(1024, gc_div_up(n, 1024));
    double* partial = nullptr;
    GC_CHECK(gc_malloc(ctx, (void**)&partial, sizeof(double) * (NB + 1)));
    GC_LAUNCH(ctx, "k_block_sums", k_block_sums<<<NB, 256, 0, ctx->stream>>>(vals, n, partial));
    GC_LAUNCH(ctx, "k_final_sum", k_final_sum<<<1, 256, 0, ctx->stream>>>(partial, NB, partial + NB));
    double* h = (double*)ctx->h_pinned;
    GC_CUDA(cudaMemcpyAsync(h, partial + NB, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    total += h[0];
    gc_free(ctx, partial);
    if (!values_dev) gc_free(ctx, vals);
    GC_CHECK(gc_launch_check(ctx, "k_measures"));
  }
  const bool use_bg = sel->kind == GECUT_SEL_PHYSICAL || sel->kind == GECUT_SEL_ACTIVE || sel->kind == GECUT_SEL_BGONLY;
  if (use_bg) {
    for (int t = 0; t < g.cpc; ++t)
      if (sel->bg_type_count[t] > 0) total += (double)sel->bg_type_count[t] * host_bgcell_measure(g, t);
  }
  *total_host = total;
  return GECUT_OK;
}

int gc_integrate_laplacian(gecut_ctx* ctx, gecut_selection* sel, double* K_bg_host) {
  const gecut_result* res = sel->res;
  const GcGrid& g = res->grid;
  const int64_t ncut = res->sizes.num_cutcells;
  const int nn = g.nv * g.nv;
  if (K_bg_host)
    for (int t = 0; t < g.cpc; ++t) host_bgcell_laplacian(g, t, K_bg_host + (size_t)t * nn);
  if (ncut == 0) return GECUT_OK;
  if (!sel->K_cut) GC_CHECK(gc_malloc(ctx, (void**)&sel->K_cut, sizeof(double) * ncut * nn));
  const int T = 128;
  const unsigned nb = (unsigned)gc_div_up(ncut * 32, T);
  if (g.D == 2)
    GC_LAUNCH(ctx, "k_cutcell_laplacian", k_cutcell_laplacian<2><<<nb, T, 0, ctx->stream>>>(g, res->lay, sel->sub_ids, sel->n_sub, ncut, ctx->d_simplexify_raw,
                                                      sel->K_cut));
  else
    GC_LAUNCH(ctx, "k_cutcell_laplacian", k_cutcell_laplacian<3><<<nb, T, 0, ctx->stream>>>(g, res->lay, sel->sub_ids, sel->n_sub, ncut, ctx->d_simplexify_raw,
                                                      sel->K_cut));
  return gc_launch_check(ctx, "k_cutcell_laplacian");
}
