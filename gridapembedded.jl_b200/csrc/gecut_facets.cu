// gecut_facets.cu -- cut_facets(LevelSetCutter(), bgmodel, geo): the cut of the background FACET grid.
//
// Reference path replaced:
//   _cut_ls_facets, compute_bgfacet_to_inoutcut            src/LevelSetCutters/LevelSetCutters.jl:107-142,159-173
//   initial_sub_triangulation on Grid(ReferenceFE{D-1},m)  src/LevelSetCutters/CutTriangulations.jl:443-560
//   _ensure_positive_jacobians_facets!                     :642-683
//   cut_sub_triangulation_several_levelsets                :294-307  (no boundary)
//   EmbeddedFacetDiscretization + selectors                src/Interfaces/EmbeddedFacetDiscretizations.jl:28-35,112-248
//
// B200 design.  The reference materialises the facet grid (connectivity of Nf facets through Gridap's topology) and
// the L nodal value vectors, then re-emits the sub-mesh once per level set.  Here the facet grid of a Cartesian model
// is never stored: Gridap's first-encounter facet numbering has a closed form (facet id <-> (cell, local facet)), so a
// thread decodes its facet id into node multi-indices with three integer divisions, evaluates the leaves at the 2^(D-1)
// nodes and writes the Int8 states lane-contiguously; "CUT by any level set" goes into a bit mask by warp ballot, block
// totals are scanned and the cut facet list is emitted in ascending id order (= the reference's findall).  The cut
// facets are then walked depth-first, one thread per stage-0 simplex (2 per QUAD facet, 1 per SEGMENT): the final arrays
// of the reference's stage-major loop are exactly the depth-first order of each simplex' refinement tree, so only the
// final layout is written (count walk -> exclusive scan -> fill walk).
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "gecut_internal.cuh"

// ------------------------------------------------------------------------------------------------
// objects behind the opaque handles
// ------------------------------------------------------------------------------------------------
struct GcFacetLayout {
  int8_t* bg_state;   // L x Nf (pitched)
  int64_t bg_pitch;
  int32_t* cutfacets; // Ncutf, 1-based ascending
  int32_t* c2p;       // Nsf * D
  int32_t* c2bg;      // Nsf (bg facet ids, 1-based)
  double* X;          // Nsfp * D
  double* R;          // Nsfp * (D-1)
  int8_t* state;      // L x Nsf (pitched)
  int64_t pitch;
  double* meas;       // Nsf: integral of 1 with the degree-2 rule
};

struct gecut_facet_result {
  gecut_ctx* ctx = nullptr;
  GcGrid grid{};
  int L = 0;
  GcLeaves leaves{};
  gecut_facet_sizes sizes{};
  GcFacetLayout lay{};
  uint32_t* cutmask = nullptr;
  bool owns_nodal[GC_LMAX] = {false};
  double* nodal_dev[GC_LMAX] = {nullptr};
};

struct gecut_facet_selection {
  const gecut_facet_result* res = nullptr;
  int which = 0, kind = 0, side = 0;
  int64_t n_sub = 0, n_bg = 0;
  int32_t* sub_ids = nullptr;
  int32_t* bg_ids = nullptr;
  int64_t bg_dir_count[3] = {0, 0, 0};  // selected whole facets per normal direction
};

namespace {

// ------------------------------------------------------------------------------------------------
// Facet numbering of a Cartesian model (Gridap generate_cell_to_faces: first encounter over (cell, local facet)).
// Local facets, highest direction first: lf = 2*(D-1-d) + side (side 0: low face, 1: high face); the low face of
// direction d is new only for cells with index 0 along d.
// ------------------------------------------------------------------------------------------------
struct FacetIdx {
  int ijk[3];  // cell
  int d;       // direction the facet is normal to
  int side;    // 0 low, 1 high
};

template <int D>
__host__ __device__ inline int64_t facets_per_row(const GcGrid& g, int j, int k) {
  // cells of a row: D new facets each (high faces) + low faces of the higher directions; +1 for the x-low face of cell 0
  return (int64_t)g.n[0] * (D + (j == 0 ? 1 : 0) + ((D == 3 && k == 0) ? 1 : 0)) + 1;
}
__host__ __device__ inline int64_t facets_per_layer(const GcGrid& g, int k) {
  return facets_per_row<3>(g, 0, k) + (int64_t)(g.n[1] - 1) * facets_per_row<3>(g, 1, k);
}
template <int D>
__host__ __device__ inline int64_t num_facets(const GcGrid& g) {
  if (D == 2) return facets_per_row<2>(g, 0, 0) + (int64_t)(g.n[1] - 1) * facets_per_row<2>(g, 1, 0);
  return facets_per_layer(g, 0) + (int64_t)(g.n[2] - 1) * facets_per_layer(g, 1);
}

template <int D>
__device__ __forceinline__ FacetIdx facet_decode(const GcGrid& g, uint32_t f) {
  FacetIdx r;
  uint32_t rem = f;
  int k = 0;
  if (D == 3) {
    const uint32_t L0 = (uint32_t)facets_per_layer(g, 0);
    if (rem >= L0) {
      const uint32_t L1 = (uint32_t)facets_per_layer(g, 1);
      rem -= L0;
      k = 1 + (int)(rem / L1);
      rem = rem % L1;
    }
  }
  int j = 0;
  const uint32_t R0 = (uint32_t)facets_per_row<D>(g, 0, k);
  if (rem >= R0) {
    const uint32_t R1 = (uint32_t)facets_per_row<D>(g, 1, k);
    rem -= R0;
    j = 1 + (int)(rem / R1);
    rem = rem % R1;
  }
  const uint32_t per = (uint32_t)(D + (j == 0 ? 1 : 0) + ((D == 3 && k == 0) ? 1 : 0));
  int i = 0;
  if (rem >= per + 1) {
    rem -= per + 1;
    i = 1 + (int)(rem / per);
    rem = rem % per;
  }
  r.ijk[0] = i;
  r.ijk[1] = j;
  r.ijk[2] = k;
  // rem-th NEW facet of the cell, local order: [z-low], z-high, [y-low], y-high, [x-low], x-high
  int t = (int)rem;
  r.d = 0;
  r.side = 1;
#pragma unroll
  for (int d = D - 1; d >= 0; --d) {
    const int idx = d == 0 ? i : (d == 1 ? j : k);
    if (idx == 0) {
      if (t == 0) {
        r.d = d;
        r.side = 0;
        return r;
      }
      t--;
    }
    if (t == 0) {
      r.d = d;
      r.side = 1;
      return r;
    }
    t--;
  }
  return r;
}

// multi-index of local node m (0 .. 2^(D-1)-1) of the facet: the free directions ascending, lower one fastest
template <int D>
__device__ __forceinline__ void facet_node(const FacetIdx& fi, int m, int* nijk) {
  nijk[0] = fi.ijk[0];
  nijk[1] = fi.ijk[1];
  nijk[2] = D == 3 ? fi.ijk[2] : 0;
  nijk[fi.d] += fi.side;
  int bit = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (d == fi.d) continue;
    nijk[d] += (m >> bit) & 1;
    bit++;
  }
}

template <int D>
__device__ __forceinline__ bool facet_is_boundary(const GcGrid& g, const FacetIdx& fi) {
  return fi.side == 0 || fi.ijk[fi.d] == g.n[fi.d] - 1;
}

template <int D>
__device__ __forceinline__ double leaf_at_node(const GcGrid& g, const GcLeaves& lv, int ls, const int* nijk) {
  if (lv.leaf[ls].kind == GECUT_LEAF_NODAL) {
    const int64_t node = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                                  : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
    return __ldg(lv.nodal[ls] + node);
  }
  const double x[3] = {gc_coord(g, 0, nijk[0]), gc_coord(g, 1, nijk[1]), D == 3 ? gc_coord(g, 2, nijk[2]) : 0.0};
  return gc_eval_leaf(lv.leaf[ls], x, D);
}

// ------------------------------------------------------------------------------------------------
// K1: states of every facet per level set (_compute_in_out_or_cut on the facet grid, CutTriangulations.jl:619-640),
//     cut mask + cut facets per block
// ------------------------------------------------------------------------------------------------
constexpr int FC_T = 256;

template <int D>
__global__ void __launch_bounds__(FC_T)
k_facet_classify(const GcGrid g, const GcLeaves lv, int64_t nf, int8_t* __restrict__ state, int64_t pitch,
                 uint32_t* __restrict__ cutmask, uint32_t* __restrict__ block_counts) {
  constexpr int NVF = 1 << (D - 1);
  __shared__ uint32_t s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int64_t f = (int64_t)blockIdx.x * FC_T + threadIdx.x;
  bool anycut = false;
  if (f < nf) {
    const FacetIdx fi = facet_decode<D>(g, (uint32_t)f);
    int nijk[NVF][3];
#pragma unroll
    for (int m = 0; m < NVF; ++m) facet_node<D>(fi, m, nijk[m]);
    for (int ls = 0; ls < lv.L; ++ls) {
      bool any = false, all = true;
#pragma unroll
      for (int m = 0; m < NVF; ++m) {
        const bool o = leaf_at_node<D>(g, lv, ls, nijk[m]) > 0.0;  // isout (LookupTables.jl:40-42)
        any |= o;
        all &= o;
      }
      const int8_t st = all ? (int8_t)GC_OUT : (any ? (int8_t)GC_CUT : (int8_t)GC_IN);
      state[(int64_t)ls * pitch + f] = st;
      anycut |= (st == GC_CUT);
    }
  }
  const uint32_t word = __ballot_sync(0xffffffffu, anycut);
  if ((threadIdx.x & 31) == 0) {
    if (f < nf) cutmask[f >> 5] = word;
    if (word) atomicAdd(&s_cnt, (uint32_t)__popc(word));
  }
  __syncthreads();
  if (threadIdx.x == 0) block_counts[blockIdx.x] = s_cnt;
}

// K2: ascending list of the cut facets (_find_cut_cells, CutTriangulations.jl:481-493)
__global__ void __launch_bounds__(FC_T)
k_facet_emit(int64_t nf, const uint32_t* __restrict__ cutmask, const uint32_t* __restrict__ block_offs,
             int32_t* __restrict__ cutfacets) {
  const int64_t f = (int64_t)blockIdx.x * FC_T + threadIdx.x;
  if (f >= nf) return;
  const int64_t w0 = ((int64_t)blockIdx.x * FC_T) >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t word = cutmask[w0 + warp];
  if (!((word >> lane) & 1u)) return;
  uint32_t pos = block_offs[blockIdx.x];
  for (int w = 0; w < warp; ++w) pos += (uint32_t)__popc(cutmask[w0 + w]);
  pos += (uint32_t)__popc(word & ((1u << lane) - 1u));
  cutfacets[pos] = (int32_t)(f + 1);
}

// ------------------------------------------------------------------------------------------------
// K3: depth-first walk of the refinement tree of every stage-0 simplex of the cut facets
// ------------------------------------------------------------------------------------------------
template <int D, int NW>
struct FPt {
  double x[D];
  double r[D - 1];
  double w[NW];
};

constexpr int FW_T = 128;

template <int D, int NW, bool FILL>
__global__ void __launch_bounds__(FW_T)
k_facet_walk(const GcGrid g, const GcLeaves lv, const GcSimplexTable* __restrict__ tables,
             const int32_t* __restrict__ cutfacets, int64_t nsimp, uint32_t* __restrict__ cnt, GcFacetLayout lay) {
  constexpr int DS = D - 1, NP = DS + 1, CE = DS, NPMAX = NP + CE, NVF = 1 << DS, NT0 = DS == 2 ? 2 : 1;
  __shared__ GcSimplexTable s_tab;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables + (DS - 1));
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
    for (int i = threadIdx.x; i < (int)(sizeof(GcSimplexTable) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const GcSimplexTable& T = s_tab;
  const int64_t s = (int64_t)blockIdx.x * FW_T + threadIdx.x;
  if (s >= nsimp) return;
  const int L = lv.L;
  const int32_t facet1 = cutfacets[s / NT0];
  const int t = (int)(s % NT0);
  const FacetIdx fi = facet_decode<D>(g, (uint32_t)(facet1 - 1));

  FPt<D, NW> pool[NP + CE * NW];
  uint8_t v[NW][NP], e[NW][NPMAX];
  int8_t child[NW], cs[NW], ptop_at[NW], st[NW];

  // ---- stage 0 (_simplexify :562-617 on the QUAD / SEGMENT facet, ltcell_to_lpoints made positive in reference space)
  double q0[3] = {0, 0, 0}, qa[DS][3];
#pragma unroll
  for (int m = 0; m < NP; ++m) {
    const int lvid = DS == 2 ? gc_simplexify_pos(0, t, m) : m;
    int nijk[3];
    facet_node<D>(fi, lvid, nijk);
    const double x[3] = {gc_coord(g, 0, nijk[0]), gc_coord(g, 1, nijk[1]), D == 3 ? gc_coord(g, 2, nijk[2]) : 0.0};
#pragma unroll
    for (int d = 0; d < D; ++d) pool[m].x[d] = x[d];
#pragma unroll
    for (int d = 0; d < DS; ++d) pool[m].r[d] = (double)((lvid >> d) & 1);
    for (int ls = 0; ls < L; ++ls) pool[m].w[ls] = leaf_at_node<D>(g, lv, ls, nijk);
  }
  // _ensure_positive_jacobians_facets! (:642-683): inner product of the Jacobians of the sub-simplex and of the bg
  // facet map at the origin of the reference facet; first two ids swapped when negative
  {
    int n0[3];
    facet_node<D>(fi, 0, n0);
    q0[0] = gc_coord(g, 0, n0[0]);
    q0[1] = gc_coord(g, 1, n0[1]);
    q0[2] = D == 3 ? gc_coord(g, 2, n0[2]) : 0.0;
    double dot = 0.0;
#pragma unroll
    for (int a = 0; a < DS; ++a) {
      int na[3];
      facet_node<D>(fi, 1 << a, na);
      qa[a][0] = gc_coord(g, 0, na[0]);
      qa[a][1] = gc_coord(g, 1, na[1]);
      qa[a][2] = D == 3 ? gc_coord(g, 2, na[2]) : 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d)
        dot = __dadd_rn(dot, __dmul_rn(__dsub_rn(pool[a + 1].x[d], pool[0].x[d]), __dsub_rn(qa[a][d], q0[d])));
    }
    if (dot < 0.0) {
      FPt<D, NW> tmp = pool[0];
      pool[0] = pool[1];
      pool[1] = tmp;
    }
  }
  uint32_t cell_off = 0, pt_off = 0;
  if (FILL) {
    cell_off = cnt[s];
    pt_off = cnt[nsimp + s];
  }
  int ptop = NP;
#pragma unroll
  for (int i = 0; i < NP; ++i) v[0][i] = (uint8_t)i;
  for (int k = 0; k < NW; ++k) st[k] = 0;

  // ---- cut_sub_triangulation_several_levelsets (:294-307): level sets popped last to first; depth d cuts ls = L-1-d
  int depth = 0;
  child[0] = -1;
  while (depth >= 0) {
    const int ls = L - 1 - depth;
    if (child[depth] < 0) {
      // _cut_cell! (:179-222): case, vertex copies, one new point per edge with a sign change
      int c = 0;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        e[depth][i] = v[depth][i];
        if (pool[v[depth][i]].w[ls] > 0.0) c |= 1 << i;
      }
      cs[depth] = (int8_t)c;
      ptop_at[depth] = (int8_t)ptop;
      int np = NP;
      if (T.inoutcut[c] == GC_CUT) {
        for (int ed = 0; ed < T.nedges; ++ed) {
          const FPt<D, NW>& a = pool[e[depth][T.edges[ed][0]]];
          const FPt<D, NW>& b = pool[e[depth][T.edges[ed][1]]];
          if ((a.w[ls] > 0.0) != (b.w[ls] > 0.0)) {
            FPt<D, NW>& o = pool[ptop];
            const double w1 = fabs(a.w[ls]), w2 = fabs(b.w[ls]);
            const double coeff = __ddiv_rn(w1, __dadd_rn(w1, w2));
            if (FILL) {
#pragma unroll
              for (int d = 0; d < D; ++d) o.x[d] = gc_lerp(a.x[d], b.x[d], coeff);
#pragma unroll
              for (int d = 0; d < DS; ++d) o.r[d] = gc_lerp(a.r[d], b.r[d], coeff);
            }
            for (int j = 0; j < ls; ++j) o.w[j] = gc_lerp(a.w[j], b.w[j], coeff);
            e[depth][np++] = (uint8_t)ptop++;
          }
        }
      }
      if (ls == 0) {
        // last stage: this parent's point block and children are final
        const int nsub = T.nsub[c];
        if (FILL) {
          for (int p = 0; p < np; ++p) {
            const FPt<D, NW>& P = pool[e[depth][p]];
#pragma unroll
            for (int d = 0; d < D; ++d) lay.X[(int64_t)(pt_off + p) * D + d] = P.x[d];
#pragma unroll
            for (int d = 0; d < DS; ++d) lay.R[(int64_t)(pt_off + p) * DS + d] = P.r[d];
          }
          for (int j = 0; j < nsub; ++j) {
            const uint32_t cid = cell_off + j;
            double pc[NP][3];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              lay.c2p[(int64_t)cid * NP + i] = (int32_t)(pt_off + T.sub_pts[c][j][i] + 1);
#pragma unroll
              for (int d = 0; d < D; ++d) pc[i][d] = pool[e[depth][T.sub_pts[c][j][i]]].x[d];
            }
            lay.c2bg[cid] = facet1;
            lay.meas[cid] = integrate_one<DS>(simplex_meas<D, DS>(pc));
            lay.state[cid] = T.sub_inout[c][j];
            for (int kk = 1; kk < L; ++kk) lay.state[(int64_t)kk * lay.pitch + cid] = st[kk];
          }
        }
        cell_off += nsub;
        pt_off += np;
        ptop = ptop_at[depth];
        depth--;
        continue;
      }
      child[depth] = 0;
    }
    const int c = cs[depth];
    if (child[depth] < T.nsub[c]) {
      const int j = child[depth]++;
      st[ls] = T.sub_inout[c][j];
#pragma unroll
      for (int i = 0; i < NP; ++i) v[depth + 1][i] = e[depth][T.sub_pts[c][j][i]];
      child[depth + 1] = -1;
      depth++;
    } else {
      ptop = ptop_at[depth];
      depth--;
    }
  }
  if (!FILL) {
    cnt[s] = cell_off;
    cnt[nsimp + s] = pt_off;
  }
}

// ------------------------------------------------------------------------------------------------
// selectors (EmbeddedFacetDiscretizations.jl:147-201) and measures
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void k_facet_flag_bg(const GcGrid g, const int8_t* __restrict__ rows, int64_t pitch, int64_t nf,
                                const GcProgram prog, int which, int mode, int side, uint32_t* __restrict__ flags,
                                unsigned long long* __restrict__ dir_counts) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  int dir = 0;
  if (f < nf) {
    const FacetIdx fi = facet_decode<D>(g, (uint32_t)f);
    const bool in_trian = facet_is_boundary<D>(g, fi) == (which == GECUT_FACETS_BOUNDARY);
    const int a = gc_eval_inoutcut(prog, rows, pitch, f);
    ok = in_trian && (mode == 1 ? a == side : (a == GC_CUT || a == side));
    flags[f] = ok ? 1u : 0u;
    dir = fi.d;
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const uint32_t m = __ballot_sync(0xffffffffu, ok && dir == d);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&dir_counts[d], (unsigned long long)__popc(m));
  }
}

template <int D>
__global__ void k_facet_flag_sub(const GcGrid g, const GcFacetLayout lay, int64_t nsf, const GcProgram prog, int which,
                                 int side, uint32_t* __restrict__ flags) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nsf) return;
  const int64_t f = (int64_t)lay.c2bg[e] - 1;
  const FacetIdx fi = facet_decode<D>(g, (uint32_t)f);
  const bool in_trian = facet_is_boundary<D>(g, fi) == (which == GECUT_FACETS_BOUNDARY);
  const int a = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, f);
  const int b = gc_eval_inoutcut(prog, lay.state, lay.pitch, e);
  flags[e] = (in_trian && a == GC_CUT && b == side) ? 1u : 0u;
}

// deterministic sum of meas[ids[i]-1]: block partials in a fixed order, the last block adds them up
constexpr int SUM_T = 256, SUM_PER = 8;
__global__ void __launch_bounds__(SUM_T)
k_facet_gather_sum(const int32_t* __restrict__ ids, int64_t n, const double* __restrict__ meas,
                   double* __restrict__ values, double* __restrict__ partial) {
  __shared__ double s[SUM_T];
  const int64_t base = (int64_t)blockIdx.x * SUM_T * SUM_PER;
  double acc = 0.0;
  for (int k = 0; k < SUM_PER; ++k) {
    const int64_t i = base + (int64_t)k * SUM_T + threadIdx.x;
    if (i < n) {
      const double m = meas[(int64_t)ids[i] - 1];
      if (values) values[i] = m;
      acc += m;
    }
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = SUM_T / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
__global__ void k_facet_final_sum(const double* __restrict__ partial, int64_t nb, double* __restrict__ out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < nb; i += 256) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

template <int D>
__global__ void k_facet_topology(const GcGrid g, int64_t nf, int32_t* __restrict__ nodes, int8_t* __restrict__ isb) {
  constexpr int NVF = 1 << (D - 1);
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const FacetIdx fi = facet_decode<D>(g, (uint32_t)f);
  if (isb) isb[f] = facet_is_boundary<D>(g, fi) ? 1 : 0;
  if (nodes) {
#pragma unroll
    for (int m = 0; m < NVF; ++m) {
      int nijk[3];
      facet_node<D>(fi, m, nijk);
      const int64_t node = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                                    : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
      nodes[f * NVF + m] = (int32_t)(node + 1);
    }
  }
}

inline int64_t pad_pitch(int64_t n) { return ((n + 255) / 256) * 256; }

void free_facet_arrays(gecut_facet_result* r) {
  gecut_ctx* ctx = r->ctx;
  GcFacetLayout& l = r->lay;
  gc_free(ctx, l.bg_state);
  gc_free(ctx, l.cutfacets);
  gc_free(ctx, l.c2p);
  gc_free(ctx, l.c2bg);
  gc_free(ctx, l.X);
  gc_free(ctx, l.R);
  gc_free(ctx, l.state);
  gc_free(ctx, l.meas);
  gc_free(ctx, r->cutmask);
  for (int i = 0; i < GC_LMAX; ++i)
    if (r->owns_nodal[i]) gc_free(ctx, r->nodal_dev[i]);
  l = GcFacetLayout{};
  r->cutmask = nullptr;
}

int read_u64(gecut_ctx* ctx, const uint64_t* dev, int n, uint64_t* out) {
  uint64_t* h = (uint64_t*)ctx->h_pinned;
  GC_CUDA(cudaMemcpyAsync(h, dev, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) out[i] = h[i];
  return GECUT_OK;
}

template <int D>
int classify_facets(gecut_ctx* ctx, gecut_facet_result* r) {
  const GcGrid& g = r->grid;
  const int64_t nf = r->sizes.num_bgfacets;
  const int64_t nblocks = gc_div_up(nf, FC_T);
  uint32_t* block_counts = nullptr;
  uint64_t* total_dev = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&block_counts, sizeof(uint32_t) * nblocks));
  GC_CHECK(gc_malloc(ctx, (void**)&total_dev, sizeof(uint64_t)));
  GC_CHECK(gc_malloc(ctx, (void**)&r->cutmask, sizeof(uint32_t) * gc_div_up(nf, 32)));
  GC_LAUNCH(ctx, "k_facet_classify",
            k_facet_classify<D><<<(unsigned)nblocks, FC_T, 0, ctx->stream>>>(g, r->leaves, nf, r->lay.bg_state, r->lay.bg_pitch,
                                                                            r->cutmask, block_counts));
  GC_CHECK(gc_launch_check(ctx, "k_facet_classify"));
  GC_CHECK(gc_exclusive_scan_u32(ctx, block_counts, block_counts, nblocks, total_dev));
  uint64_t total = 0;
  GC_CHECK(read_u64(ctx, total_dev, 1, &total));
  r->sizes.num_cutfacets = (int64_t)total;
  if (total > 0) {
    GC_CHECK(gc_malloc(ctx, (void**)&r->lay.cutfacets, sizeof(int32_t) * total));
    GC_LAUNCH(ctx, "k_facet_emit",
              k_facet_emit<<<(unsigned)nblocks, FC_T, 0, ctx->stream>>>(nf, r->cutmask, block_counts, r->lay.cutfacets));
    GC_CHECK(gc_launch_check(ctx, "k_facet_emit"));
  }
  gc_free(ctx, block_counts);
  gc_free(ctx, total_dev);
  return GECUT_OK;
}

template <int D, int NW>
int walk_facets_nw(gecut_ctx* ctx, gecut_facet_result* r) {
  constexpr int NT0 = D == 3 ? 2 : 1;
  const int64_t nsimp = r->sizes.num_cutfacets * NT0;
  const int L = r->L;
  uint32_t* cnt = nullptr;
  uint64_t* totals_dev = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&cnt, sizeof(uint32_t) * 2 * nsimp));
  GC_CHECK(gc_malloc(ctx, (void**)&totals_dev, sizeof(uint64_t) * 2));
  const unsigned nb = (unsigned)gc_div_up(nsimp, FW_T);
  GC_LAUNCH(ctx, "k_facet_walk_count",
            (k_facet_walk<D, NW, false><<<nb, FW_T, 0, ctx->stream>>>(r->grid, r->leaves, ctx->d_tables, r->lay.cutfacets, nsimp,
                                                                     cnt, r->lay)));
  GC_CHECK(gc_launch_check(ctx, "k_facet_walk_count"));
  GC_CHECK(gc_exclusive_scan_u32_multi(ctx, cnt, cnt, nsimp, 2, totals_dev));
  uint64_t tot[2] = {0, 0};
  GC_CHECK(read_u64(ctx, totals_dev, 2, tot));
  if (tot[0] >= 2147483647ull || tot[1] >= 2147483647ull) {
    ctx->err = "sub-facet ids overflow Int32";
    gc_free(ctx, cnt);
    gc_free(ctx, totals_dev);
    return GECUT_ERR_OVERFLOW;
  }
  r->sizes.num_subfacets = (int64_t)tot[0];
  r->sizes.num_subfacet_points = (int64_t)tot[1];
  GcFacetLayout& l = r->lay;
  l.pitch = pad_pitch((int64_t)tot[0]);
  GC_CHECK(gc_malloc(ctx, (void**)&l.c2p, sizeof(int32_t) * tot[0] * D));
  GC_CHECK(gc_malloc(ctx, (void**)&l.c2bg, sizeof(int32_t) * tot[0]));
  GC_CHECK(gc_malloc(ctx, (void**)&l.meas, sizeof(double) * tot[0]));
  GC_CHECK(gc_malloc(ctx, (void**)&l.state, (size_t)l.pitch * L));
  GC_CHECK(gc_malloc(ctx, (void**)&l.X, sizeof(double) * tot[1] * D));
  GC_CHECK(gc_malloc(ctx, (void**)&l.R, sizeof(double) * tot[1] * (D - 1)));
  GC_LAUNCH(ctx, "k_facet_walk_fill",
            (k_facet_walk<D, NW, true><<<nb, FW_T, 0, ctx->stream>>>(r->grid, r->leaves, ctx->d_tables, r->lay.cutfacets, nsimp,
                                                                    cnt, r->lay)));
  GC_CHECK(gc_launch_check(ctx, "k_facet_walk_fill"));
  gc_free(ctx, cnt);
  gc_free(ctx, totals_dev);
  return GECUT_OK;
}

template <int D>
int walk_facets(gecut_ctx* ctx, gecut_facet_result* r) {
  if (r->sizes.num_cutfacets == 0) return GECUT_OK;
  const int L = r->L;
  if (L <= 1) return walk_facets_nw<D, 1>(ctx, r);
  if (L <= 2) return walk_facets_nw<D, 2>(ctx, r);
  if (L <= 4) return walk_facets_nw<D, 4>(ctx, r);
  if (L <= 8) return walk_facets_nw<D, 8>(ctx, r);
  return walk_facets_nw<D, GC_LMAX>(ctx, r);
}

template <class T>
int to_host(gecut_ctx* ctx, T* host, const T* dev, int64_t count) {
  if (!host || count == 0 || !dev) return GECUT_OK;
  GC_CUDA(cudaMemcpyAsync(host, dev, sizeof(T) * count, cudaMemcpyDeviceToHost, ctx->stream));
  return GECUT_OK;
}

int rows_to_host(gecut_ctx* ctx, int8_t* host, const int8_t* dev, int64_t pitch, int64_t n, int L) {
  if (!host || n == 0 || !dev) return GECUT_OK;
  GC_CUDA(cudaMemcpy2DAsync(host, (size_t)n, dev, (size_t)pitch, (size_t)n, (size_t)L, cudaMemcpyDeviceToHost, ctx->stream));
  return GECUT_OK;
}

int eval_rows_to_host(gecut_ctx* ctx, const int8_t* rows, int64_t pitch, int64_t n, int L, const int32_t* program, int len,
                      int8_t* host_out) {
  if (n == 0) return GECUT_OK;
  GcProgram p;
  GC_CHECK(gc_make_program(ctx, program, len, L, &p));
  int8_t* d = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&d, (size_t)pad_pitch(n)));
  int st = gc_eval_rows(ctx, rows, pitch, n, p, d, true);
  if (st == GECUT_OK) st = to_host(ctx, host_out, d, n);
  if (st == GECUT_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "facet state evaluation failed";
    st = GECUT_ERR_CUDA;
  }
  gc_free(ctx, d);
  return st;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int gecut_cut_facets(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaves, int L,
                     const double* const* nodal_values, int nodal_on_device, int classify_only,
                     gecut_facet_result** out) {
  if (!ctx) return GECUT_ERR_INVALID;
  if (!out || !leaves || !grid) {
    ctx->err = "NULL argument";
    return GECUT_ERR_INVALID;
  }
  *out = nullptr;
  if (L < 1 || L > GC_LMAX) {
    ctx->err = "number of level sets must be in 1..GECUT_MAX_LEVELSETS";
    return L < 1 ? GECUT_ERR_INVALID : GECUT_ERR_NOTIMPLEMENTED;
  }
  if (grid->simplexified) {
    ctx->err = "cut_facets is implemented for n-cube Cartesian models only (not simplexified)";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  if (!(grid->slab_begin == 0 && grid->slab_end == 0)) {
    ctx->err = "cut_facets works on the whole model (no slab)";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  DeviceGuard guard(ctx->device);
  gecut_facet_result* r = new (std::nothrow) gecut_facet_result();
  if (!r) return GECUT_ERR_OOM;
  r->ctx = ctx;
  r->L = L;
  int st = gc_make_grid(ctx, grid, &r->grid);
  if (st != GECUT_OK) {
    delete r;
    return st;
  }
  const GcGrid& g = r->grid;
  auto fail = [&](int code) {
    free_facet_arrays(r);
    cudaStreamSynchronize(ctx->stream);
    delete r;
    return code;
  };
  const int D = g.D;
  const int64_t nf = D == 2 ? num_facets<2>(g) : num_facets<3>(g);
  if (nf >= 2147483647ll) {
    ctx->err = "facet ids overflow Int32";
    return fail(GECUT_ERR_OVERFLOW);
  }
  r->leaves.L = L;
  r->leaves.nodal_base = 0;
  const int64_t nn_total = g.layer_nodes * g.nn[D - 1];
  for (int i = 0; i < L; ++i) {
    r->leaves.leaf[i] = leaves[i];
    r->leaves.nodal[i] = nullptr;
    const int kind = leaves[i].kind;
    if (kind < GECUT_LEAF_SPHERE || kind > GECUT_LEAF_NODAL) {
      ctx->err = "unknown leaf kind";
      return fail(GECUT_ERR_INVALID);
    }
    if (kind == GECUT_LEAF_DOUGHNUT && D != 3) {
      ctx->err = "doughnut leaves need a 3-D model";
      return fail(GECUT_ERR_INVALID);
    }
    if (kind == GECUT_LEAF_NODAL) {
      if (!nodal_values || !nodal_values[i]) {
        ctx->err = "NODAL leaf without nodal values";
        return fail(GECUT_ERR_INVALID);
      }
      if (nodal_on_device) {
        r->leaves.nodal[i] = nodal_values[i];
      } else {
        double* d = nullptr;
        st = gc_malloc(ctx, (void**)&d, sizeof(double) * nn_total);
        if (st != GECUT_OK) return fail(st);
        r->nodal_dev[i] = d;
        r->owns_nodal[i] = true;
        if (cudaMemcpyAsync(d, nodal_values[i], sizeof(double) * nn_total, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
          ctx->err = "H2D copy of nodal values failed";
          return fail(GECUT_ERR_CUDA);
        }
        r->leaves.nodal[i] = d;
      }
    }
  }
  gecut_facet_sizes& sz = r->sizes;
  memset(&sz, 0, sizeof(sz));
  sz.num_bgfacets = nf;
  sz.num_levelsets = L;
  sz.D = D;
  sz.num_vertices_per_bgfacet = 1 << (D - 1);
  // facets with one cell around: the two faces of the box per direction
  if (D == 2)
    sz.num_boundary_facets = 2ll * (g.n[0] + g.n[1]);
  else
    sz.num_boundary_facets = 2ll * ((int64_t)g.n[0] * g.n[1] + (int64_t)g.n[0] * g.n[2] + (int64_t)g.n[1] * g.n[2]);
  r->lay.bg_pitch = pad_pitch(nf);
  st = gc_malloc(ctx, (void**)&r->lay.bg_state, (size_t)r->lay.bg_pitch * L);
  if (st != GECUT_OK) return fail(st);
  st = D == 2 ? classify_facets<2>(ctx, r) : classify_facets<3>(ctx, r);
  if (st != GECUT_OK) return fail(st);
  if (!classify_only) {
    st = D == 2 ? walk_facets<2>(ctx, r) : walk_facets<3>(ctx, r);
    if (st != GECUT_OK) return fail(st);
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = std::string("cut_facets: ") + cudaGetErrorString(cudaGetLastError());
    return fail(GECUT_ERR_CUDA);
  }
  *out = r;
  return GECUT_OK;
}

int gecut_facet_result_free(gecut_facet_result* res) {
  if (!res) return GECUT_OK;
  DeviceGuard guard(res->ctx->device);
  free_facet_arrays(res);
  delete res;
  return GECUT_OK;
}

int gecut_facet_result_sizes(const gecut_facet_result* res, gecut_facet_sizes* sizes) {
  if (!res || !sizes) return GECUT_ERR_INVALID;
  *sizes = res->sizes;
  return GECUT_OK;
}

int gecut_facet_result_copy(const gecut_facet_result* res, const gecut_facet_buffers* host) {
  if (!res || !host) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = res->ctx;
  DeviceGuard guard(ctx->device);
  const gecut_facet_sizes& s = res->sizes;
  const GcFacetLayout& l = res->lay;
  const int D = s.D, L = s.num_levelsets;
  GC_CHECK(rows_to_host(ctx, host->ls_to_facet_to_inoutcut, l.bg_state, l.bg_pitch, s.num_bgfacets, L));
  GC_CHECK(to_host(ctx, host->cutfacet_to_facet, l.cutfacets, s.num_cutfacets));
  GC_CHECK(to_host(ctx, host->sf_cell_to_points, l.c2p, s.num_subfacets * D));
  GC_CHECK(to_host(ctx, host->sf_cell_to_bgcell, l.c2bg, s.num_subfacets));
  GC_CHECK(to_host(ctx, host->sf_point_to_coords, l.X, s.num_subfacet_points * D));
  GC_CHECK(to_host(ctx, host->sf_point_to_rcoords, l.R, s.num_subfacet_points * (D - 1)));
  GC_CHECK(rows_to_host(ctx, host->ls_to_subfacet_to_inout, l.state, l.pitch, s.num_subfacets, L));
  int32_t* nodes = nullptr;
  int8_t* isb = nullptr;
  if (host->facet_to_nodes || host->facet_is_boundary) {
    const int64_t nf = s.num_bgfacets;
    const int nvf = s.num_vertices_per_bgfacet;
    if (host->facet_to_nodes) GC_CHECK(gc_malloc(ctx, (void**)&nodes, sizeof(int32_t) * nf * nvf));
    if (host->facet_is_boundary) GC_CHECK(gc_malloc(ctx, (void**)&isb, (size_t)nf));
    const int T = 256;
    if (D == 2)
      GC_LAUNCH(ctx, "k_facet_topology", k_facet_topology<2><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, nf, nodes, isb));
    else
      GC_LAUNCH(ctx, "k_facet_topology", k_facet_topology<3><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, nf, nodes, isb));
    GC_CHECK(gc_launch_check(ctx, "k_facet_topology"));
    GC_CHECK(to_host(ctx, host->facet_to_nodes, nodes, nf * nvf));
    GC_CHECK(to_host(ctx, host->facet_is_boundary, isb, nf));
  }
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  gc_free(ctx, nodes);
  gc_free(ctx, isb);
  return GECUT_OK;
}

int gecut_bgfacet_to_inoutcut(const gecut_facet_result* res, const int32_t* program, int len, int8_t* host_out) {
  if (!res || !host_out) return GECUT_ERR_INVALID;
  DeviceGuard guard(res->ctx->device);
  return eval_rows_to_host(res->ctx, res->lay.bg_state, res->lay.bg_pitch, res->sizes.num_bgfacets, res->L, program, len,
                           host_out);
}

int gecut_subfacet_to_inout(const gecut_facet_result* res, const int32_t* program, int len, int8_t* host_out) {
  if (!res || !host_out) return GECUT_ERR_INVALID;
  DeviceGuard guard(res->ctx->device);
  return eval_rows_to_host(res->ctx, res->lay.state, res->lay.pitch, res->sizes.num_subfacets, res->L, program, len,
                           host_out);
}

int gecut_select_facets(const gecut_facet_result* res, int which, int kind, const int32_t* program, int len, int side,
                        gecut_facet_selection** out) {
  if (!res || !out) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = res->ctx;
  *out = nullptr;
  if ((which != GECUT_FACETS_BOUNDARY && which != GECUT_FACETS_SKELETON) || (side != GECUT_IN && side != GECUT_OUT) ||
      (kind != GECUT_SEL_PHYSICAL && kind != GECUT_SEL_ACTIVE && kind != GECUT_SEL_CUTONLY && kind != GECUT_SEL_BGONLY)) {
    ctx->err = "invalid facet selection";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  GcProgram p;
  GC_CHECK(gc_make_program(ctx, program, len, res->L, &p));
  gecut_facet_selection* sel = new (std::nothrow) gecut_facet_selection();
  if (!sel) return GECUT_ERR_OOM;
  sel->res = res;
  sel->which = which;
  sel->kind = kind;
  sel->side = side;
  auto fail = [&](int code) {
    gc_free(ctx, sel->sub_ids);
    gc_free(ctx, sel->bg_ids);
    delete sel;
    return code;
  };
  const int D = res->sizes.D;
  const int T = 256;
  int st = GECUT_OK;
  const int64_t nsf = res->sizes.num_subfacets, nf = res->sizes.num_bgfacets;
  if ((kind == GECUT_SEL_PHYSICAL || kind == GECUT_SEL_CUTONLY) && nsf > 0) {
    uint32_t* flags = nullptr;
    st = gc_malloc(ctx, (void**)&flags, sizeof(uint32_t) * nsf);
    if (st != GECUT_OK) return fail(st);
    if (D == 2)
      GC_LAUNCH(ctx, "k_facet_flag_sub", k_facet_flag_sub<2><<<(unsigned)gc_div_up(nsf, T), T, 0, ctx->stream>>>(res->grid, res->lay, nsf, p, which, side, flags));
    else
      GC_LAUNCH(ctx, "k_facet_flag_sub", k_facet_flag_sub<3><<<(unsigned)gc_div_up(nsf, T), T, 0, ctx->stream>>>(res->grid, res->lay, nsf, p, which, side, flags));
    st = gc_launch_check(ctx, "k_facet_flag_sub");
    if (st == GECUT_OK) st = gc_compact_flags(ctx, flags, nsf, &sel->sub_ids, &sel->n_sub, nullptr, nullptr);
    gc_free(ctx, flags);
    if (st != GECUT_OK) return fail(st);
  }
  if (kind != GECUT_SEL_CUTONLY && nf > 0) {
    uint32_t* flags = nullptr;
    unsigned long long* dirc = nullptr;
    st = gc_malloc(ctx, (void**)&flags, sizeof(uint32_t) * nf);
    if (st != GECUT_OK) return fail(st);
    st = gc_malloc(ctx, (void**)&dirc, sizeof(unsigned long long) * 3);
    if (st != GECUT_OK) {
      gc_free(ctx, flags);
      return fail(st);
    }
    cudaMemsetAsync(dirc, 0, sizeof(unsigned long long) * 3, ctx->stream);
    const int mode = kind == GECUT_SEL_ACTIVE ? 2 : 1;
    if (D == 2)
      GC_LAUNCH(ctx, "k_facet_flag_bg", k_facet_flag_bg<2><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, res->lay.bg_state, res->lay.bg_pitch, nf, p, which, mode, side, flags, dirc));
    else
      GC_LAUNCH(ctx, "k_facet_flag_bg", k_facet_flag_bg<3><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, res->lay.bg_state, res->lay.bg_pitch, nf, p, which, mode, side, flags, dirc));
    st = gc_launch_check(ctx, "k_facet_flag_bg");
    if (st == GECUT_OK) st = gc_compact_flags(ctx, flags, nf, &sel->bg_ids, &sel->n_bg, nullptr, nullptr);
    if (st == GECUT_OK) {
      uint64_t h[3];
      st = read_u64(ctx, (const uint64_t*)dirc, 3, h);
      for (int d = 0; d < 3; ++d) sel->bg_dir_count[d] = (int64_t)h[d];
    }
    gc_free(ctx, flags);
    gc_free(ctx, dirc);
    if (st != GECUT_OK) return fail(st);
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "facet selection failed";
    return fail(GECUT_ERR_CUDA);
  }
  *out = sel;
  return GECUT_OK;
}

int gecut_facet_selection_free(gecut_facet_selection* sel) {
  if (!sel) return GECUT_OK;
  gecut_ctx* ctx = sel->res->ctx;
  DeviceGuard guard(ctx->device);
  gc_free(ctx, sel->sub_ids);
  gc_free(ctx, sel->bg_ids);
  delete sel;
  return GECUT_OK;
}

int gecut_facet_selection_sizes(const gecut_facet_selection* sel, int64_t* n_sub, int64_t* n_bg) {
  if (!sel) return GECUT_ERR_INVALID;
  if (n_sub) *n_sub = sel->n_sub;
  if (n_bg) *n_bg = sel->n_bg;
  return GECUT_OK;
}

int gecut_facet_selection_copy(const gecut_facet_selection* sel, int32_t* sub_ids, int32_t* bg_ids) {
  if (!sel) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = sel->res->ctx;
  DeviceGuard guard(ctx->device);
  GC_CHECK(to_host(ctx, sub_ids, sel->sub_ids, sel->n_sub));
  GC_CHECK(to_host(ctx, bg_ids, sel->bg_ids, sel->n_bg));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  return GECUT_OK;
}

int gecut_facet_selection_measure(const gecut_facet_selection* sel, double* sub_values, double* total) {
  if (!sel) return GECUT_ERR_INVALID;
  const gecut_facet_result* res = sel->res;
  gecut_ctx* ctx = res->ctx;
  DeviceGuard guard(ctx->device);
  double sum = 0.0;
  if (sel->n_sub > 0) {
    const int64_t n = sel->n_sub;
    const int64_t nb = gc_div_up(n, (int64_t)SUM_T * SUM_PER);
    double *partial = nullptr, *vals = nullptr;
    GC_CHECK(gc_malloc(ctx, (void**)&partial, sizeof(double) * (nb + 1)));
    if (sub_values) GC_CHECK(gc_malloc(ctx, (void**)&vals, sizeof(double) * n));
    GC_LAUNCH(ctx, "k_facet_gather_sum",
              k_facet_gather_sum<<<(unsigned)nb, SUM_T, 0, ctx->stream>>>(sel->sub_ids, n, res->lay.meas, vals, partial));
    GC_CHECK(gc_launch_check(ctx, "k_facet_gather_sum"));
    GC_LAUNCH(ctx, "k_facet_final_sum", k_facet_final_sum<<<1, 256, 0, ctx->stream>>>(partial, nb, partial + nb));
    GC_CHECK(gc_launch_check(ctx, "k_facet_final_sum"));
    double* h = (double*)ctx->h_pinned;
    GC_CUDA(cudaMemcpyAsync(h, partial + nb, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GC_CHECK(to_host(ctx, sub_values, vals, n));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    sum = h[0];
    gc_free(ctx, partial);
    gc_free(ctx, vals);
  }
  // whole facets: tensor Gauss rule of degree 2 on the (D-1)-cube, sum_q w_q * meas (the facet map is affine)
  const GcGrid& g = res->grid;
  const int D = g.D;
  for (int d = 0; d < D; ++d) {
    if (sel->bg_dir_count[d] == 0) continue;
    double meas = 1.0;
    if (D == 2) {
      const double e = g.dx[1 - d];
      meas = sqrt(e * e);
    } else {
      const int a = d == 0 ? 1 : 0, b = d == 2 ? 1 : 2;
      const double n = g.dx[a] * g.dx[b];
      meas = sqrt(n * n);
    }
    const int nq = 1 << (D - 1);
    const double w = D == 2 ? 0.5 : 0.25;
    double one = 0.0;
    for (int q = 0; q < nq; ++q) one += w * meas;
    sum += one * (double)sel->bg_dir_count[d];
  }
  if (total) *total = sum;
  return GECUT_OK;
}

}  // extern "C"
