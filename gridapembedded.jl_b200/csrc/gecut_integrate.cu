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
#include <cstring>

#include "gecut_internal.cuh"

namespace {

// per-entity measures (integral of 1 with the degree-2 rule) are produced by the subtriangulation kernels while the
// vertices are on chip; a Measure over a selection only gathers them
__global__ void __launch_bounds__(256)
k_measures(const int32_t* __restrict__ ids, int64_t n, const double* __restrict__ meas, double* __restrict__ values,
           double* __restrict__ partial, unsigned int* __restrict__ ticket, double* __restrict__ out) {
  // gather + deterministic two-level sum: fixed element -> thread -> tree order inside a block, and the block that
  // finishes last adds the block partials in index order (no second launch); values may be NULL
  __shared__ double sh[256];
  __shared__ bool s_last;
  double s = 0.0;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi = lo + per < n ? lo + per : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const double v = meas[ids ? (int64_t)ids[i] - 1 : i];  // ids == NULL: the selection is 1..n
    if (values) values[i] = v;
    s += v;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = sh[0];
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double t = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += reinterpret_cast<volatile double*>(partial)[i];
  sh[threadIdx.x] = t;
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

// One warp per cut bg cell.  The selection (bg state CUT && subcell state == side, EmbeddedDiscretizations.jl:303-311) is
// evaluated on the fly from the state rows, so the kernel reads each subcell of the cell exactly once.
//   * n-cube cells (Q1): the integrand grad(phi_a).grad(phi_b) is a tensor-product polynomial; with
//     P_0(t)=(1-t)^2, P_1(t)=t(1-t), P_2(t)=t^2 it only needs the moments
//       M_d[i][j] = sum_{s,q} w_q |J_s| P_i(eta_e) P_j(eta_f)   ((e,f) = the directions other than d)
//     (27 numbers in 3-D, 6 in 2-D), accumulated per lane over (subcell, quadrature point) items, reduced across the warp
//     through shared memory, and K_ab = sum_d sgn_a sgn_b / h_d^2 * M_d[a_e+b_e][a_f+b_f].
//   * simplex cells (P1): gradients are constant, K_ab = (g_a.g_b) * sum_s sum_q w_q |J_s|.
// G LANES per cut cell (G = 4): lane g of the group walks the subcells s0+g, s0+g+G, ... of the cell's contiguous run with
// the 27 (6) moments in registers, the G partial sums are added with a fixed butterfly (reproducible), the moments of the
// warp's 32/G cells are parked in shared memory and the warp writes their matrices as contiguous 256-byte stores (lane =
// matrix entry).  Compared with one thread per cell: G times more warps and G times shorter dependent gather chains
// (ids -> reference points), and the trip counts inside a warp differ by the subcell count / G instead of the subcell
// count itself (cut hexes carry 6 .. 36 subcells).
constexpr int LAPQ_G = 4;
constexpr int LAPQ_T = 128;

template <int D>
__global__ void __launch_bounds__(LAPQ_T)
k_cutcell_laplacian_q1(const GcGrid g, const GcLayout lay, const uint32_t* __restrict__ sc_ptr, int64_t ncut,
                       const GcProgram prog, int side, int check_bg, double* __restrict__ K) {
  constexpr int G = LAPQ_G;
  constexpr int CW = 32 / G;  // cells per warp
  constexpr int NV = 1 << D;
  constexpr int NQ = D + 1;
  constexpr int NM = D == 3 ? 27 : 6;
  constexpr int NN = NV * NV;
  __shared__ double s_M[LAPQ_T / 32][CW][NM + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gl = lane % G, cw = lane / G;
  const int64_t kw0 = ((int64_t)blockIdx.x * (LAPQ_T / 32) + wid) * CW;  // first cell of this warp
  const int64_t k = kw0 + cw;
  const double qa = D == 3 ? (5.0 - sqrt(5.0)) / 20.0 : 1.0 / 6;
  const double qb = D == 3 ? (5.0 + 3.0 * sqrt(5.0)) / 20.0 : 2.0 / 3;
  // accumulators of a lane: 2-D the 6 moments themselves; 3-D the 19 MONOMIAL moments sum w t_e^a t_f^b (a, b <= 2, at most
  // two directions per monomial: 1, t_d, t_d^2, and {t_e t_f, t_e t_f^2, t_e^2 t_f, t_e^2 t_f^2} per pair) from which the 27
  // P-basis moments follow by two 3x3 triangular transforms per direction once per cell -- 30 instead of 48 FP64
  // instructions per quadrature point and 16 registers less
  constexpr int NA = D == 3 ? 19 : NM;
  double M[NA];
#pragma unroll
  for (int m = 0; m < NA; ++m) M[m] = 0.0;
  if (k < ncut && (!check_bg || gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, (int64_t)lay.cutcells[k] - 1) == GC_CUT)) {
    const uint32_t s0 = sc_ptr[2 * k] + (uint32_t)gl, s1 = sc_ptr[2 * k + 2];
    // software pipeline: connectivity row, state and measure of this lane's NEXT subcell are requested while the current
    // one is processed, so an iteration waits for one memory round trip (the reference points), not two
    struct Ids {
      int32_t v[4];
    };
    auto load_ids = [&](int64_t e) -> Ids {
      Ids o;
      if (D == 3) {
        const int4 q4 = *reinterpret_cast<const int4*>(lay.sc_c2p + e * 4);
        o.v[0] = q4.x;
        o.v[1] = q4.y;
        o.v[2] = q4.z;
        o.v[3] = q4.w;
      } else {
#pragma unroll
        for (int v = 0; v <= D; ++v) o.v[v] = lay.sc_c2p[e * (D + 1) + v];
        o.v[3] = 0;
      }
      return o;
    };
    Ids nids{};
    int nstate = 0;
    double nmeas = 0.0;
    if (s0 < s1) {
      nids = load_ids((int64_t)s0);
      nstate = gc_eval_inoutcut(prog, lay.sc_state, lay.sc_pitch, (int64_t)s0);
      nmeas = lay.sc_meas[s0];
    }
    for (uint32_t it = s0; it < s1; it += G) {
      const Ids ids_s = nids;
      const int state = nstate;
      const double meas = nmeas;
      if (it + G < s1) {
        nids = load_ids((int64_t)it + G);
        nstate = gc_eval_inoutcut(prog, lay.sc_state, lay.sc_pitch, (int64_t)it + G);
        nmeas = lay.sc_meas[it + G];
      }
      if (state != side) continue;
      double r[D + 1][3];
#pragma unroll
      for (int v = 0; v <= D; ++v) {
        const int64_t pid = (int64_t)ids_s.v[v] - 1;
#pragma unroll
        for (int d = 0; d < D; ++d) r[v][d] = lay.sc_R[pid * D + d];
      }
      const double wdV = meas * (1.0 / NQ);  // sc_meas = sum_q w_q |J_s| (all weights equal)
      double rs[3];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        double t = r[0][d];
#pragma unroll
        for (int i = 1; i <= D; ++i) t += r[i][d];
        rs[d] = qa * t;
      }
      if (D == 3) M[0] += meas;  // sum_q w_q |J_s| (4 equal weights)
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        if (D == 3) {
          double t[3], sq[3], wt[3], ws[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            // eta_q = sum_i N_i(xi_q) R_i with N_i = qb for i == q and qa otherwise
            t[d] = rs[d] + (qb - qa) * r[q][d];
            sq[d] = t[d] * t[d];
            wt[d] = wdV * t[d];
            ws[d] = wdV * sq[d];
            M[1 + d] += wt[d];
            M[4 + d] += ws[d];
          }
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const int e1 = d == 0 ? 1 : 0, f1 = d == 2 ? 1 : 2;
            M[7 + 4 * d + 0] = fma(wt[e1], t[f1], M[7 + 4 * d + 0]);
            M[7 + 4 * d + 1] = fma(wt[e1], sq[f1], M[7 + 4 * d + 1]);
            M[7 + 4 * d + 2] = fma(ws[e1], t[f1], M[7 + 4 * d + 2]);
            M[7 + 4 * d + 3] = fma(ws[e1], sq[f1], M[7 + 4 * d + 3]);
          }
        } else {
          double P[3][3];
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const double t = rs[d] + (qb - qa) * r[q][d], u = 1.0 - t;
            P[d][0] = u * u;
            P[d][1] = t * u;
            P[d][2] = t * t;
          }
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const int e1 = 1 - d;
#pragma unroll
            for (int i = 0; i < 3; ++i) M[d * 3 + i] = fma(wdV, P[e1][i], M[d * 3 + i]);
          }
        }
      }
    }
  }
  // ---- the G partial sums of a cell: fixed butterfly, then lane 0 of the group parks the moments
#pragma unroll
  for (int o = 1; o < G; o <<= 1)
#pragma unroll
    for (int m = 0; m < NA; ++m) M[m] += __shfl_xor_sync(0xffffffffu, M[m], o);
  if (gl == 0) {
    if (D == 3) {
      // monomial moments -> P-basis moments, P_0 = 1 - 2t + t^2, P_1 = t - t^2, P_2 = t^2 in each of the two directions
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int e1 = d == 0 ? 1 : 0, f1 = d == 2 ? 1 : 2;
        const double A[3][3] = {{M[0], M[1 + f1], M[4 + f1]},
                                {M[1 + e1], M[7 + 4 * d + 0], M[7 + 4 * d + 1]},
                                {M[4 + e1], M[7 + 4 * d + 2], M[7 + 4 * d + 3]}};
        double B[3][3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          B[0][b] = (A[0][b] - 2.0 * A[1][b]) + A[2][b];
          B[1][b] = A[1][b] - A[2][b];
          B[2][b] = A[2][b];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          s_M[wid][cw][d * 9 + i * 3 + 0] = (B[i][0] - 2.0 * B[i][1]) + B[i][2];
          s_M[wid][cw][d * 9 + i * 3 + 1] = B[i][1] - B[i][2];
          s_M[wid][cw][d * 9 + i * 3 + 2] = B[i][2];
        }
      }
    } else {
#pragma unroll
      for (int m = 0; m < NM; ++m) s_M[wid][cw][m] = M[m];
    }
  }
  __syncwarp();
  // ---- element matrices of the warp's cells from the moments: lane = entry of the contiguous output
  double sih2[3];
#pragma unroll
  for (int d = 0; d < D; ++d) sih2[d] = 1.0 / (g.dx[d] * g.dx[d]);
  double* out = K + kw0 * NN;
  const int64_t nvalid = (ncut - kw0) * NN;  // entries of existing cells
#pragma unroll 4
  for (int j = 0; j < CW * NN / 32; ++j) {
    const int e = lane + 32 * j;
    const int cl = e / NN, idx = e % NN;
    const int a = idx / NV, bb = idx % NV;
    const double* Mc = s_M[wid][cl];
    double val = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double sg = (((a >> d) & 1) == ((bb >> d) & 1)) ? sih2[d] : -sih2[d];
      int mi;
      if (D == 3) {
        const int e1 = d == 0 ? 1 : 0, f1 = d == 2 ? 1 : 2;
        mi = d * 9 + (((a >> e1) & 1) + ((bb >> e1) & 1)) * 3 + (((a >> f1) & 1) + ((bb >> f1) & 1));
      } else {
        const int e1 = 1 - d;
        mi = d * 3 + (((a >> e1) & 1) + ((bb >> e1) & 1));
      }
      val += sg * Mc[mi];
    }
    if (e < nvalid) out[e] = val;
  }
}

// constant gradient products g_a.g_b of the 6 (2) simplex types of a Cartesian n-cube: a kernel parameter (no H2D copy, so
// the device-resident Laplacian can return without synchronising), staged into shared memory by the kernels
struct GcGradDots {
  double v[6 * 16];
};

// Simplex bg cells (P1): one thread per cut cell accumulates V = sum over its selected subcells of sum_q w_q |J_s|; the
// matrix is V * (g_a.g_b) with the constant gradient products of the cell type (6 TET / 2 TRI types of a Cartesian
// n-cube), written warp-cooperatively so that the 32 matrices of a warp go out as contiguous 256-byte stores.
template <int D>
__global__ void __launch_bounds__(128)
k_cutcell_laplacian_p1(const GcGrid g, const GcLayout lay, const uint32_t* __restrict__ sc_ptr, int64_t ncut,
                       const GcProgram prog, int side, int check_bg, const GcGradDots gd,
                       double* __restrict__ K) {
  constexpr int NV = D + 1, NN = NV * NV;
  __shared__ double graddots[6 * 16];
  if (threadIdx.x < 6 * 16) graddots[threadIdx.x] = gd.v[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double V = 0.0;
  int ty = 0;
  if (k < ncut) {
    const int64_t cell0 = (int64_t)lay.cutcells[k] - 1;
    ty = (int)(cell0 % g.cpc);
    // one level set: every cut cell is CUT for every program over it, no need to look at the bg state row
    if (!check_bg || gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell0) == GC_CUT) {
      const uint32_t s0 = sc_ptr[2 * k], s1 = sc_ptr[2 * k + 2];
      for (uint32_t e = s0; e < s1; ++e) {
        if (gc_eval_inoutcut(prog, lay.sc_state, lay.sc_pitch, (int64_t)e) != side) continue;
        V += lay.sc_meas[e];
      }
    }
  }
  const int64_t wbase = k - lane;
#pragma unroll
  for (int j = 0; j < NN; ++j) {
    const int idx = lane + 32 * j;
    const int cl = idx / NN, en = idx % NN;
    const double v = __shfl_sync(0xffffffffu, V, cl);
    const int t = __shfl_sync(0xffffffffu, ty, cl);
    if (wbase + cl < ncut) K[wbase * NN + idx] = v * graddots[t * NN + en];
  }
}

// the same with the per-cell (IN, OUT) measures the fill kernel left behind (one level set, single-leaf geometry):
// 16 bytes read per cut cell instead of its subcell range
template <int D>
__global__ void __launch_bounds__(128)
k_cutcell_laplacian_p1_vio(const GcLayout lay, int cpc, const double2* __restrict__ cell_vio, int64_t ncut, int side,
                           const GcGradDots gd, double* __restrict__ K) {
  constexpr int NV = D + 1, NN = NV * NV;
  __shared__ double graddots[6 * 16];
  if (threadIdx.x < 6 * 16) graddots[threadIdx.x] = gd.v[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double V = 0.0;
  int ty = 0;
  if (k < ncut) {
    ty = (int)(((int64_t)lay.cutcells[k] - 1) % cpc);
    const double2 v = cell_vio[k];
    V = side == GC_IN ? v.x : v.y;
  }
  const int64_t wbase = k - lane;
#pragma unroll
  for (int j = 0; j < NN; ++j) {
    const int idx = lane + 32 * j;
    const int cl = idx / NN, en = idx % NN;
    const double v = __shfl_sync(0xffffffffu, V, cl);
    const int t = __shfl_sync(0xffffffffu, ty, cl);
    if (wbase + cl < ncut) K[wbase * NN + idx] = v * graddots[t * NN + en];
  }
}

// get_cell_measure: uncut cells (whole-cell measure per cell type when the state is `side`, else 0) ...
__global__ void k_bgcell_measure(const GcLayout lay, int64_t nc, const GcProgram prog, int side, int cpc, int use_bg,
                                 const double* __restrict__ type_meas, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nc) return;
  double v = 0.0;
  if (use_bg && gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, e) == side) v = type_meas[e % cpc];
  out[e] = v;
}

// ... then the cut cells: sum of the selected subcells in subcell order (move_contributions)
__global__ void k_cutcell_measure(const GcLayout lay, const uint32_t* __restrict__ sc_ptr, int64_t ncut, const GcProgram prog,
                                  int side, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncut) return;
  const int64_t cell0 = (int64_t)lay.cutcells[k] - 1;
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell0) != GC_CUT) return;
  double V = 0.0;
  const uint32_t s0 = sc_ptr[2 * k], s1 = sc_ptr[2 * k + 2];
  for (uint32_t e = s0; e < s1; ++e)
    if (gc_eval_inoutcut(prog, lay.sc_state, lay.sc_pitch, (int64_t)e) == side) V += lay.sc_meas[e];
  out[cell0] = V;
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

// g_a.g_b of the P1 shape functions of an uncut simplex cell of type `typ` (host)
static void host_bgcell_graddots(const GcGrid& g, int typ, double* G) {
  const int D = g.D, nv = g.nv;
  int ijk[3] = {0, 0, 0};
  ijk[D - 1] = g.k0;
  double x[8][3], invJt[3][3], gr[8][3], gx[8][3];
  bgcell_vertices(g, ijk, typ, &GC_SIMPLEXIFY_RAW[0][0][0], x);
  bgcell_inv_jt(g, x, invJt);
  bg_ref_grads(D, true, nullptr, gr);
  for (int a = 0; a < nv; ++a)
    for (int d = 0; d < D; ++d) {
      double v = 0;
      for (int e = 0; e < D; ++e) v += invJt[d][e] * gr[a][e];
      gx[a][d] = v;
    }
  for (int a = 0; a < nv; ++a)
    for (int b = 0; b < nv; ++b) {
      double dd = 0;
      for (int d = 0; d < D; ++d) dd += gx[a][d] * gx[b][d];
      G[a * nv + b] = dd;
    }
}

int gc_integrate_measure(gecut_ctx* ctx, const gecut_selection* sel, double* values_dev, double* total_host) {
  const gecut_result* res = sel->res;
  const GcGrid& g = res->grid;
  const GcLayout& lay = res->lay;
  const int64_t n = sel->n_sub;
  double total = 0.0;
  const bool use_sub = sel->kind == GECUT_SEL_PHYSICAL || sel->kind == GECUT_SEL_CUTONLY || sel->kind == GECUT_SEL_BOUNDARY;
  // one level set: the sums of the fill kernel answer without a pass (per-entity values still need the gather)
  const bool from_stats = res->has_l1_stats && !values_dev &&
                          ((sel->sub_lazy && sel->kind != GECUT_SEL_BOUNDARY) || (sel->sub_identity && sel->kind == GECUT_SEL_BOUNDARY));
  if (use_sub && n > 0 && from_stats) {
    total += sel->kind == GECUT_SEL_BOUNDARY ? res->l1_stats[2] : res->l1_stats[sel->lazy_side == GECUT_IN ? 0 : 1];
  } else if (use_sub && n > 0) {
    GC_CHECK(gc_materialize_sub_ids(ctx, const_cast<gecut_selection*>(sel)));
    GC_CHECK(gc_ensure_measures(ctx, res));  // simplex cells, one level set: not written at cut time
    const int NB = (int)std::min<int64_t>(1184, gc_div_up(n, 1024));
    GcScratch scratch(ctx);
    double* partial = nullptr;  // NB block partials, the total, the completion ticket
    GC_CHECK(scratch.alloc(&partial, sizeof(double) * (NB + 2)));
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + NB + 1);
    GC_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), ctx->stream));
    GC_LAUNCH(ctx, "k_measures", k_measures<<<NB, 256, 0, ctx->stream>>>(
        sel->sub_ids, n, sel->kind == GECUT_SEL_BOUNDARY ? lay.sf_meas : lay.sc_meas, values_dev, partial, ticket,
        partial + NB));
    double* h = (double*)ctx->h_pinned;
    GC_CUDA(cudaMemcpyAsync(h, partial + NB, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    total += h[0];
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
  const int check_bg = res->L > 1 ? 1 : 0;
  if (g.simplexified) {
    GcGradDots Gd{};
    for (int t = 0; t < g.cpc; ++t) host_bgcell_graddots(g, t, Gd.v + (size_t)t * nn);
    const unsigned nb1 = (unsigned)gc_div_up(ncut, 128);
    const GcProgram& pp = sel->prog;
    const bool leaf1 = (pp.len == 1 && pp.tok[0] >= 0) || (pp.len == 2 && pp.tok[0] >= 0 && pp.tok[1] == GECUT_OP_COMPLEMENT);
    if (res->cell_vio && res->L == 1 && leaf1) {
      const int eff = pp.len == 1 ? sel->side : -sel->side;
      if (g.D == 2)
        GC_LAUNCH(ctx, "k_cutcell_laplacian", k_cutcell_laplacian_p1_vio<2><<<nb1, 128, 0, ctx->stream>>>(
            res->lay, g.cpc, res->cell_vio, ncut, eff, Gd, sel->K_cut));
      else
        GC_LAUNCH(ctx, "k_cutcell_laplacian", k_cutcell_laplacian_p1_vio<3><<<nb1, 128, 0, ctx->stream>>>(
            res->lay, g.cpc, res->cell_vio, ncut, eff, Gd, sel->K_cut));
      return gc_launch_check(ctx, "k_cutcell_laplacian_p1_vio");
    }
    GC_CHECK(gc_ensure_measures(ctx, res));
    if (g.D == 2)
      GC_LAUNCH(ctx, "k_cutcell_laplacian", k_cutcell_laplacian_p1<2><<<nb1, 128, 0, ctx->stream>>>(
          g, res->lay, res->cut_sc_ptr, ncut, sel->prog, sel->side, check_bg, Gd, sel->K_cut));
    else
      GC_LAUNCH(ctx, "k_cutcell_laplacian", k_cutcell_laplacian_p1<3><<<nb1, 128, 0, ctx->stream>>>(
          g, res->lay, res->cut_sc_ptr, ncut, sel->prog, sel->side, check_bg, Gd, sel->K_cut));
    return gc_launch_check(ctx, "k_cutcell_laplacian_p1");
  }
  GC_CHECK(gc_ensure_measures(ctx, res));
  const int T = LAPQ_T;
  const unsigned nb = (unsigned)gc_div_up(ncut, T / LAPQ_G);
  if (g.D == 2)
    GC_LAUNCH(ctx, "k_cutcell_laplacian", k_cutcell_laplacian_q1<2><<<nb, T, 0, ctx->stream>>>(
        g, res->lay, res->cut_sc_ptr, ncut, sel->prog, sel->side, check_bg, sel->K_cut));
  else
    GC_LAUNCH(ctx, "k_cutcell_laplacian", k_cutcell_laplacian_q1<3><<<nb, T, 0, ctx->stream>>>(
        g, res->lay, res->cut_sc_ptr, ncut, sel->prog, sel->side, check_bg, sel->K_cut));
  return gc_launch_check(ctx, "k_cutcell_laplacian");
}

int gc_cell_measure(gecut_ctx* ctx, const gecut_selection* sel, double* values_dev) {
  const gecut_result* res = sel->res;
  const GcGrid& g = res->grid;
  const int64_t nc = res->sizes.num_bgcells, ncut = res->sizes.num_cutcells;
  if (nc == 0) return GECUT_OK;
  const bool use_bg = sel->kind == GECUT_SEL_PHYSICAL || sel->kind == GECUT_SEL_BGONLY;
  const bool use_sub = sel->kind == GECUT_SEL_PHYSICAL || sel->kind == GECUT_SEL_CUTONLY;
  double* tm = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&tm, sizeof(double) * 8));
  double* hm = reinterpret_cast<double*>((char*)ctx->h_pinned + 2048);
  for (int t = 0; t < 8; ++t) hm[t] = t < g.cpc ? host_bgcell_measure(g, t) : 0.0;
  GC_CUDA(cudaMemcpyAsync(tm, hm, sizeof(double) * 8, cudaMemcpyHostToDevice, ctx->stream));
  const int T = 256;
  GC_LAUNCH(ctx, "k_bgcell_measure", k_bgcell_measure<<<(unsigned)gc_div_up(nc, T), T, 0, ctx->stream>>>(
      res->lay, nc, sel->prog, sel->side, g.cpc, use_bg ? 1 : 0, tm, values_dev));
  if (use_sub && ncut > 0 && res->cut_sc_ptr) GC_CHECK(gc_ensure_measures(ctx, res));
  if (use_sub && ncut > 0 && res->cut_sc_ptr)
    GC_LAUNCH(ctx, "k_cutcell_measure", k_cutcell_measure<<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(
        res->lay, res->cut_sc_ptr, ncut, sel->prog, sel->side, values_dev));
  int st = gc_launch_check(ctx, "cell_measure");
  gc_free(ctx, tm);
  return st;
}
