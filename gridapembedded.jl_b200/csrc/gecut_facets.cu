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
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "gecut_internal.cuh"
#include "gecut_facets.cuh"

// Slot table of the facet grid of a simplexified model (gecut_facets.cuh), from the simplexify table of the n-cube.
bool gc_build_sx_facets(int D, GcSxFacets* t) {
  static const uint8_t LF2[3][3] = {{0, 1, 0}, {0, 2, 0}, {1, 2, 0}};
  static const uint8_t LF3[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
  memset(t, 0, sizeof(*t));
  t->D = D;
  t->nt = D == 2 ? 2 : 6;
  t->nlf = D + 1;
  t->nslots = t->nt * t->nlf;
  const int ns = t->nslots, nvf = D;
  auto key_of = [&](const uint8_t* v, int flip) {  // the node set as a bit mask over the 2^D cube vertices
    unsigned m = 0;
    for (int i = 0; i < nvf; ++i) m |= 1u << (v[i] ^ flip);
    return m;
  };
  for (int s = 0; s < ns; ++s) {
    const int tt = s / t->nlf, lf = s % t->nlf;
    for (int m = 0; m < nvf; ++m) t->vtx[s][m] = GC_SIMPLEXIFY_RAW[D - 2][tt][D == 2 ? LF2[lf][m] : LF3[lf][m]];
    t->face[s] = -1;
    for (int d = 0; d < D; ++d) {
      int ones = 0;
      for (int m = 0; m < nvf; ++m) ones += (t->vtx[s][m] >> d) & 1;
      if (ones == 0) t->face[s] = (int8_t)(d << 1);
      if (ones == nvf) t->face[s] = (int8_t)((d << 1) | 1);
    }
  }
  for (int s = 0; s < ns; ++s) {
    const unsigned key = key_of(t->vtx[s], t->face[s] < 0 ? 0 : 1 << (t->face[s] >> 1));
    int found = -1;
    for (int s2 = 0; s2 < ns; ++s2)
      if (s2 != s && key_of(t->vtx[s2], 0) == key) found = s2;
    if (found < 0) return false;  // the simplices of neighbouring cubes do not share this face: not a conforming tiling
    t->partner[s] = (uint8_t)found;
  }
  int nown[8];
  for (int cls = 0; cls < (1 << D); ++cls) {
    int n = 0;
    for (int s = 0; s < ns; ++s) {
      const int fc = t->face[s];
      const bool own = fc < 0 ? s < t->partner[s] : ((fc & 1) ? true : ((cls >> (fc >> 1)) & 1) != 0);
      t->rank[cls][s] = own ? (uint8_t)n : (uint8_t)0xff;
      if (own) t->order[cls][n++] = (uint8_t)s;
    }
    nown[cls] = n;
  }
  t->C0 = nown[0];
  t->cd = nown[1] - nown[0];
  for (int cls = 0; cls < (1 << D); ++cls) {
    int z = 0;
    for (int d = 0; d < D; ++d) z += (cls >> d) & 1;
    if (nown[cls] != t->C0 + t->cd * z) return false;
  }
  return true;
}

namespace {

// Node coordinate for the facet kernels.  On a z-slab (layers [k0, k1) of the last direction) the facet grid is the LOCAL
// model of the rank -- partition (.., k1 - k0), local facet numbering, as every rank of the reference's distributed mode
// cuts its own local model (src/Distributed/DistributedDiscretizations.jl:42-53) -- but its nodes keep the GLOBAL
// coordinates pmin + (k0 + k) * dx, so a slab's facets are bit-identical to the same facets of the whole model.
template <int D>
__device__ __forceinline__ double fc_coord(const GcGrid& g, int d, int i) {
  return gc_coord(g, d, d == D - 1 ? i + g.k0 : i);
}

template <int D>
__device__ __forceinline__ double leaf_at_node(const GcGrid& g, const GcLeaves& lv, int ls, const int* nijk) {
  if (lv.leaf[ls].kind == GECUT_LEAF_NODAL) {
    const int64_t node = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                                  : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
    return __ldg(lv.nodal[ls] + node);
  }
  const double x[3] = {fc_coord<D>(g, 0, nijk[0]), fc_coord<D>(g, 1, nijk[1]), D == 3 ? fc_coord<D>(g, 2, nijk[2]) : 0.0};
  return gc_eval_leaf(lv.leaf[ls], x, D);
}

// ------------------------------------------------------------------------------------------------
// K1: sign bit of every level set at every node (discretize + isout, DiscreteGeometries.jl:48-63, LookupTables.jl:40-42):
//     one evaluation per node, 32 x-consecutive nodes -> one ballot word.  A warp owns a word column and marches through
//     NS_KCH planes of the last direction; the in-plane partial sums of the closed-form leaves are hoisted out of the
//     march with the association of the reference ((w1*w1 + w2*w2) + w3*w3, gc_dot3), so a node costs 3-6 FP64
//     operations.  The bit planes (Nn/8 bytes per level set, 17 MB at 513^3) stay in L2 for the next pass; node VALUES
//     never touch HBM.
// ------------------------------------------------------------------------------------------------
constexpr int NS_KCH = 32;

template <int D>
__global__ void __launch_bounds__(256)
k_node_signs(const GcGrid g, const GcLeaves lv, int wxt, uint32_t* __restrict__ signs, int64_t ls_stride) {
  constexpr int M = D - 1;  // marching direction
  const int wx = blockIdx.x;
  int j = 0, kbeg;
  if (D == 3) {
    j = blockIdx.y * 8 + threadIdx.y;
    if (j >= g.nn[1]) return;  // a warp is one threadIdx.y: it leaves as a whole
    kbeg = blockIdx.z * NS_KCH;
  } else {
    kbeg = (blockIdx.y * 8 + threadIdx.y) * NS_KCH;
    if (kbeg >= g.nn[1]) return;
  }
  const int kend = min(kbeg + NS_KCH, g.nn[M]);
  const int i = wx * 32 + threadIdx.x;
  const bool valid = i < g.nn[0];
  const int ic = valid ? i : g.nn[0] - 1;
  const double x = fc_coord<D>(g, 0, ic);
  const double y = D == 3 ? fc_coord<D>(g, 1, j) : 0.0;
  for (int ls = 0; ls < lv.L; ++ls) {
    const gecut_leaf& lf = lv.leaf[ls];
    const int kind = lf.kind;
    uint32_t* out = signs + (int64_t)ls * ls_stride + wx;
    if (kind == GECUT_LEAF_NODAL) {
      for (int k = kbeg; k < kend; ++k) {
        const int64_t nrow = D == 3 ? (int64_t)k * g.nn[1] + j : (int64_t)k;
        const double val = __ldg(lv.nodal[ls] + nrow * g.nn[0] + ic);
        const uint32_t word = __ballot_sync(0xffffffffu, valid && val > 0.0);
        if (threadIdx.x == 0) out[nrow * wxt] = word;
      }
      continue;
    }
    const double w0 = __dsub_rn(x, lf.x0[0]);
    const double w1 = D == 3 ? __dsub_rn(y, lf.x0[1]) : 0.0;
    double ww = __dmul_rn(w0, w0), wv = __dmul_rn(w0, lf.v[0]);
    if (D == 3) {
      ww = __dadd_rn(ww, __dmul_rn(w1, w1));
      wv = __dadd_rn(wv, __dmul_rn(w1, lf.v[1]));
    }
    const double R2 = __dmul_rn(lf.R, lf.R), r2 = __dmul_rn(lf.r, lf.r);
    const double tt = kind == GECUT_LEAF_DOUGHNUT ? __dsub_rn(lf.R, __dsqrt_rn(ww)) : 0.0;
    const double t2 = __dmul_rn(tt, tt);
    for (int k = kbeg; k < kend; ++k) {
      const double wl = __dsub_rn(fc_coord<D>(g, M, k), lf.x0[M]);
      double val;
      if (kind == GECUT_LEAF_SPHERE) {
        val = __dsub_rn(__dadd_rn(ww, __dmul_rn(wl, wl)), R2);
      } else if (kind == GECUT_LEAF_PLANE) {
        val = __dadd_rn(wv, __dmul_rn(wl, lf.v[M]));
      } else if (kind == GECUT_LEAF_CYLINDER) {
        const double A = __dadd_rn(wv, __dmul_rn(wl, lf.v[M]));
        const double H2 = __dadd_rn(ww, __dmul_rn(wl, wl));
        val = __dsub_rn(__dsub_rn(H2, __dmul_rn(A, A)), R2);
      } else {  // DOUGHNUT (3-D only)
        val = __dsub_rn(__dadd_rn(t2, __dmul_rn(wl, wl)), r2);
      }
      const uint32_t word = __ballot_sync(0xffffffffu, valid && val > 0.0);
      const int64_t nrow = D == 3 ? (int64_t)k * g.nn[1] + j : (int64_t)k;
      if (threadIdx.x == 0) out[nrow * wxt] = word;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2: states of every facet per level set (_compute_in_out_or_cut on the facet grid, CutTriangulations.jl:619-640) and
//     the ascending list of the cut facets (_find_cut_cells :481-493).  The facets a row of cells meets first are a
//     contiguous id range (closed form), and consecutive rows are consecutive ranges, so the model's facets are ONE byte
//     stream ordered by (row, x word).  A lane owns a word = 32 cells of a row and works bit-parallel: the "all OUT" /
//     "any OUT" masks of a facet kind are AND / OR of the node sign words of the rows it touches (x+1 by funnel shift),
//     3-5 kinds per word.  A warp owns 32 consecutive words = one contiguous piece of the stream.
//       EMIT = false: pieces with one sign on all nodes (all but the O(n^2) pieces a surface crosses) are written as
//       constant 16-byte stores; otherwise lanes expand their masks into the warp's shared-memory slice (constant fill for
//       lanes with one sign) and the piece leaves as aligned 16-byte stores.  Tallies (IN/OUT x direction x boundary) are
//       popcounts; the number of cut facets of every piece is recorded.
//       EMIT = true (after the scan over pieces): the cut facets of the piece are written in id order.
// ------------------------------------------------------------------------------------------------
constexpr int FR_SEG = 1024;  // cells per warp segment (k_facet_rows_count)
constexpr int FW_ITER = 4;    // pieces per warp

// tallies of the whole facets per level set, fused into K2: index ((s*3 + d)*2 + isb), s = 0 IN / 1 OUT, d = normal
// direction, isb = facet on the boundary of the model.  A selection over a single-leaf geometry reads its whole-facet
// counts from here (no pass over the facets at all).
constexpr int FT_N = 12;

template <int D>
__host__ __device__ constexpr int fw_stage_bytes() { return 32 * (32 * (D + 2) + 1) + 32; }

// `len` copies of byte b at dst (shared memory, any alignment): bytes up to the first 4-byte boundary, then words
__device__ __forceinline__ void fw_fill(uint8_t* dst, int len, uint8_t b) {
  int p = 0;
  while (p < len && ((reinterpret_cast<uintptr_t>(dst + p)) & 3u)) dst[p++] = b;
  const uint32_t b4 = (uint32_t)b * 0x01010101u;
  for (; p + 4 <= len; p += 4) *reinterpret_cast<uint32_t*>(dst + p) = b4;
  for (; p < len; ++p) dst[p] = b;
}

template <int D, bool EMIT>
__global__ void __launch_bounds__(256)
k_facet_words(const GcGrid g, int L, int wxt, int wpr, const uint32_t* __restrict__ signs, int64_t ls_stride,
              int8_t* __restrict__ state, int64_t pitch, uint32_t* __restrict__ piece_counts,
              int32_t* __restrict__ cutfacets, int64_t nitems, int64_t npieces, unsigned long long* __restrict__ tallies) {
  constexpr int NR = 1 << (D - 1);
  // fixed slots of the facet kinds in local order: (low, high) of the directions D-1 .. 1, then x-high; the low slots
  // exist only in rows with index 0 along that direction, the x-low facet only for cell 0 of a row (handled apart)
  constexpr int NKS = 2 * (D - 1) + 1;
  __shared__ __align__(16) uint8_t s_stage[EMIT ? 1 : 8][EMIT ? 16 : ((fw_stage_bytes<D>() + 15) & ~15)];
  __shared__ uint32_t s_tally[EMIT ? 1 : GC_LMAX * FT_N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!EMIT) {
    for (int c = threadIdx.x; c < GC_LMAX * FT_N; c += blockDim.x) s_tally[c] = 0u;
    __syncthreads();
  }
  const int64_t gw = (int64_t)blockIdx.x * 8 + warp;
  uint8_t* stage = s_stage[EMIT ? 0 : warp];
  const int nx = g.n[0];
  for (int it = 0; it < FW_ITER; ++it) {
    const int64_t piece = gw * FW_ITER + it;
    if (piece >= npieces) break;  // warp-uniform
    const int64_t t = piece * 32 + lane;
    const bool valid = t < nitems;
    // ---- this lane's word: row (j,k), word w, cells, position in the facet stream
    const uint32_t tt = valid ? (uint32_t)t : (uint32_t)(nitems - 1);
    const uint32_t row = tt / (uint32_t)wpr;
    const int w = (int)(tt - row * (uint32_t)wpr);
    const int j = D == 3 ? (int)(row % (uint32_t)g.n[1]) : (int)row;
    const int k = D == 3 ? (int)(row / (uint32_t)g.n[1]) : 0;
    const bool lowp[3] = {false, j == 0, D == 3 && k == 0};  // low facets present, by direction
    const bool hib[3] = {false, j == g.n[1] - 1, D == 3 && k == g.n[2] - 1};  // high facets of the row on the boundary
    const int per = D + (lowp[1] ? 1 : 0) + (lowp[2] ? 1 : 0);
    const int ncell = valid ? min(32, nx - 32 * w) : 0;
    const uint32_t cmask = ncell == 32 ? 0xffffffffu : ((1u << ncell) - 1u);
    const int64_t start = facet_row_base<D>(g, j, k) + (int64_t)32 * w * per + (w > 0 ? 1 : 0);
    const int len = valid ? ncell * per + (w == 0 ? 1 : 0) : 0;
    const int64_t start0 = __shfl_sync(0xffffffffu, start, 0);
    // lanes are consecutive words of the stream: the piece is [start0, start of the last valid lane + its length)
    const int last = (int)min((int64_t)31, nitems - 1 - piece * 32);
    const int total_len = (int)(__shfl_sync(0xffffffffu, start, last) - start0) + __shfl_sync(0xffffffffu, len, last);
    const int align = (int)(start0 & 15);
    const int off = (int)(start - start0) + align;  // of this lane's first byte in the staging slice
    const bool has_xl = valid && w == 0;
    const bool lastword = valid && 32 * w + ncell == nx;  // the x-high facet of its last cell is a boundary facet
    int64_t nrow[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) nrow[r] = ((int64_t)(k + (r >> 1)) * g.nn[1] + (j + (r & 1))) * wxt + w;
    uint32_t cut[NKS];  // cut facets of this word per slot, over all level sets
#pragma unroll
    for (int q = 0; q < NKS; ++q) cut[q] = 0u;
    bool xl_cut = false;
    for (int ls = 0; ls < L; ++ls) {
      const uint32_t* sg = signs + (int64_t)ls * ls_stride;
      uint32_t S[NR], X[NR];
      uint32_t n_or = 0u, n_and = 0xffffffffu;  // over the (ncell + 1) x NR nodes of the word
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        S[r] = valid ? __ldg(sg + nrow[r]) : 0u;
        const uint32_t nxt = (valid && w + 1 < wxt) ? __ldg(sg + nrow[r] + 1) : 0u;
        X[r] = __funnelshift_r(S[r], nxt, 1);  // bit c = sign of node 32w + c + 1
        n_or |= (S[r] | X[r]) & cmask;
        n_and &= (S[r] & X[r]) | ~cmask;
      }
      const bool u_in = n_or == 0u, u_out = n_and == 0xffffffffu;
      const bool w_in = __all_sync(0xffffffffu, u_in || !valid), w_out = __all_sync(0xffffffffu, u_out || !valid);
      uint32_t all[NKS], any[NKS];
#pragma unroll
      for (int q = 0; q < NKS; ++q) {
        uint32_t a = 0xffffffffu, o = 0u;
        if (q < 2 * (D - 1)) {
          const int d = D - 1 - q / 2, sd = q & 1;  // node rows on side sd of direction d: bit (d-1) of r
#pragma unroll
          for (int r = 0; r < NR; ++r)
            if (((r >> (d - 1)) & 1) == sd) {
              a &= S[r] & X[r];
              o |= S[r] | X[r];
            }
        } else {
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            a &= X[r];
            o |= X[r];
          }
        }
        all[q] = a & cmask;
        any[q] = o & cmask;
      }
      uint32_t xl_all = 1u, xl_any = 0u;  // x-low facet of cell 0 (word 0 only): nodes i = 0 of all rows
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        xl_all &= S[r] & 1u;
        xl_any |= S[r] & 1u;
      }
#pragma unroll
      for (int q = 0; q < NKS; ++q) {
        const bool present = q >= 2 * (D - 1) || (q & 1) || lowp[D - 1 - q / 2];
        if (present) cut[q] |= any[q] & ~all[q];
      }
      if (has_xl && xl_any && !xl_all) xl_cut = true;
      if (!EMIT) {
        // ---- tallies of this level set: popcounts, summed over the warp
        uint32_t tl[FT_N];
#pragma unroll
        for (int c = 0; c < FT_N; ++c) tl[c] = 0u;
#pragma unroll
        for (int q = 0; q < NKS; ++q) {
          const uint32_t nin = (uint32_t)__popc(~any[q] & cmask), nout = (uint32_t)__popc(all[q]);
          if (q < 2 * (D - 1)) {
            const int d = D - 1 - q / 2;
            const bool low = (q & 1) == 0;
            if (low && !lowp[d]) continue;
            const bool isb = low || hib[d];
            tl[(0 * 3 + d) * 2 + 0] += isb ? 0u : nin;
            tl[(0 * 3 + d) * 2 + 1] += isb ? nin : 0u;
            tl[(1 * 3 + d) * 2 + 0] += isb ? 0u : nout;
            tl[(1 * 3 + d) * 2 + 1] += isb ? nout : 0u;
          } else {
            const uint32_t lb = lastword ? (1u << (ncell - 1)) : 0u;  // the boundary x-high facet
            const uint32_t bin = (~any[q] & lb) ? 1u : 0u, bout = (all[q] & lb) ? 1u : 0u;
            tl[(0 * 3 + 0) * 2 + 0] += nin - bin;
            tl[(0 * 3 + 0) * 2 + 1] += bin;
            tl[(1 * 3 + 0) * 2 + 0] += nout - bout;
            tl[(1 * 3 + 0) * 2 + 1] += bout;
          }
        }
        if (has_xl) {
          if (!xl_any) tl[(0 * 3 + 0) * 2 + 1] += 1u;
          if (xl_all) tl[(1 * 3 + 0) * 2 + 1] += 1u;
        }
#pragma unroll
        for (int c = 0; c < FT_N; ++c) {
          if (D == 2 && ((c >> 1) % 3) == 2) continue;  // no z direction
          const uint32_t v = __reduce_add_sync(0xffffffffu, tl[c]);
          if (lane == 0 && v) atomicAdd(&s_tally[ls * FT_N + c], v);
        }
        // ---- states
        int8_t* g0 = state + (int64_t)ls * pitch + start0 - align;
        const int end = align + total_len;
        if (w_in || w_out) {
          const uint32_t b4 = w_in ? 0xffffffffu : 0x01010101u;
          const uint4 cv = make_uint4(b4, b4, b4, b4);
          for (int p0 = lane * 16; p0 + 16 <= end; p0 += 512)
            if (p0 >= align) *reinterpret_cast<uint4*>(g0 + p0) = cv;
          if (align != 0 && lane >= align && lane < 16 && lane < end) g0[lane] = (int8_t)(b4 & 0xffu);
          const int tail0 = end & ~15;
          const int p = tail0 + (lane & 15);
          if (lane >= 16 && p < end && p >= align && (tail0 > 0 || align == 0)) g0[p] = (int8_t)(b4 & 0xffu);
        } else {
          if (valid) {
            if (u_in || u_out) {
              fw_fill(stage + off, len, u_in ? (uint8_t)0xff : (uint8_t)0x01);
            } else {
              int p = off;
              for (int c = 0; c < ncell; ++c) {
#pragma unroll
                for (int q = 0; q < NKS; ++q) {
                  const bool present = q >= 2 * (D - 1) || (q & 1) || lowp[D - 1 - q / 2];
                  if (!present) continue;
                  if (q == NKS - 1 && c == 0 && has_xl)  // x-low of cell 0 sits before its x-high
                    stage[p++] = xl_all ? (uint8_t)0x01 : (xl_any ? (uint8_t)0x00 : (uint8_t)0xff);
                  const uint32_t ab = (all[q] >> c) & 1u, ob = (any[q] >> c) & 1u;
                  stage[p++] = ab ? (uint8_t)0x01 : (ob ? (uint8_t)0x00 : (uint8_t)0xff);
                }
              }
            }
          }
          __syncwarp();
          for (int p0 = lane * 16; p0 + 16 <= end; p0 += 512)
            if (p0 >= align) *reinterpret_cast<uint4*>(g0 + p0) = *reinterpret_cast<const uint4*>(stage + p0);
          if (align != 0 && lane >= align && lane < 16 && lane < end) g0[lane] = (int8_t)stage[lane];
          const int tail0 = end & ~15;
          const int p = tail0 + (lane & 15);
          if (lane >= 16 && p < end && p >= align && (tail0 > 0 || align == 0)) g0[p] = (int8_t)stage[p];
          __syncwarp();
        }
      }
    }
    // ---- cut facets of the piece
    uint32_t cnt = xl_cut ? 1u : 0u;
#pragma unroll
    for (int q = 0; q < NKS; ++q) cnt += (uint32_t)__popc(cut[q]);
    if (!EMIT) {
      const uint32_t tot = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0) piece_counts[piece] = tot;
    } else {
      uint32_t inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
      }
      if (cnt) {
        uint32_t pos = piece_counts[piece] + inc - cnt;
        uint32_t m = xl_cut ? 1u : 0u;
#pragma unroll
        for (int q = 0; q < NKS; ++q) m |= cut[q];
        while (m) {
          const int c = __ffs(m) - 1;
          m &= m - 1u;
          int64_t f = start + (int64_t)c * per + ((w == 0 && c > 0) ? 1 : 0);
#pragma unroll
          for (int q = 0; q < NKS; ++q) {
            const bool present = q >= 2 * (D - 1) || (q & 1) || lowp[D - 1 - q / 2];
            if (!present) continue;
            if (q == NKS - 1 && c == 0 && w == 0) {  // x-low of cell 0
              if (xl_cut) cutfacets[pos++] = (int32_t)(f + 1);
              f++;
            }
            if ((cut[q] >> c) & 1u) cutfacets[pos++] = (int32_t)(f + 1);
            f++;
          }
        }
      }
    }
  }
  if (!EMIT) {
    __syncthreads();
    for (int c = threadIdx.x; c < L * FT_N; c += blockDim.x)
      if (s_tally[c]) atomicAdd(&tallies[c], (unsigned long long)s_tally[c]);
  }
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

// SX: simplexified model -- the bg facet is itself a (D-1)-simplex (one stage-0 simplex, nodes and boundary flag from the
// slot table `sx`, reference vertices of SEGMENT / TRI; the sub-simplex map IS the facet map, so no Jacobian swap).
template <int D, int NW, bool FILL, bool SX>
__global__ void __launch_bounds__(FW_T)
k_facet_walk(const GcGrid g, const GcLeaves lv, const GcSimplexTable* __restrict__ tables,
             const int32_t* __restrict__ cutfacets, int64_t nsimp, uint32_t* __restrict__ cnt, GcFacetLayout lay,
             const GcSxFacets sx) {
  constexpr int DS = D - 1, NP = DS + 1, CE = DS, NPMAX = NP + CE, NT0 = (DS == 2 && !SX) ? 2 : 1;
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
  FacetIdx fi{};
  SxFacetIdx sfi{};
  if (SX)
    sfi = sx_facet_decode<D>(g, sx, (uint32_t)(facet1 - 1));
  else
    fi = facet_decode<D>(g, (uint32_t)(facet1 - 1));
  const bool on_boundary = SX ? sx_facet_is_boundary<D>(g, sfi, sx) : facet_is_boundary<D>(g, fi);

  FPt<D, NW> pool[NP + CE * NW];
  uint8_t v[NW][NP], e[NW][NPMAX];
  int8_t child[NW], cs[NW], ptop_at[NW], st[NW];

  // ---- stage 0 (_simplexify :562-617 on the QUAD / SEGMENT facet, ltcell_to_lpoints made positive in reference space)
  double q0[3] = {0, 0, 0}, qa[DS][3];
#pragma unroll
  for (int m = 0; m < NP; ++m) {
    const int lvid = (DS == 2 && !SX) ? gc_simplexify_pos(0, t, m) : m;
    int nijk[3];
    if (SX)
      sx_facet_node<D>(sfi, sx, m, nijk);
    else
      facet_node<D>(fi, lvid, nijk);
    const double x[3] = {fc_coord<D>(g, 0, nijk[0]), fc_coord<D>(g, 1, nijk[1]), D == 3 ? fc_coord<D>(g, 2, nijk[2]) : 0.0};
#pragma unroll
    for (int d = 0; d < D; ++d) pool[m].x[d] = x[d];
#pragma unroll
    for (int d = 0; d < DS; ++d) pool[m].r[d] = SX ? (m == d + 1 ? 1.0 : 0.0) : (double)((lvid >> d) & 1);
    for (int ls = 0; ls < L; ++ls) pool[m].w[ls] = leaf_at_node<D>(g, lv, ls, nijk);
  }
  // _ensure_positive_jacobians_facets! (:642-683): inner product of the Jacobians of the sub-simplex and of the bg
  // facet map at the origin of the reference facet; first two ids swapped when negative
  if (!SX) {
    int n0[3];
    facet_node<D>(fi, 0, n0);
    q0[0] = fc_coord<D>(g, 0, n0[0]);
    q0[1] = fc_coord<D>(g, 1, n0[1]);
    q0[2] = D == 3 ? fc_coord<D>(g, 2, n0[2]) : 0.0;
    double dot = 0.0;
#pragma unroll
    for (int a = 0; a < DS; ++a) {
      int na[3];
      facet_node<D>(fi, 1 << a, na);
      qa[a][0] = fc_coord<D>(g, 0, na[0]);
      qa[a][1] = fc_coord<D>(g, 1, na[1]);
      qa[a][2] = D == 3 ? fc_coord<D>(g, 2, na[2]) : 0.0;
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
            lay.isb[cid] = on_boundary ? 1 : 0;
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
// whole facets of a selection, counted per normal direction (the measure needs nothing else).  `row` is ONE Int8 state
// row (a level set's row, or the memoised row of a multi-operator program); a warp owns a segment of a row of cells and
// reads its facets' states as aligned 16-byte vectors, the direction and boundary flag of a byte follow from its offset
// in the row (division by the warp-uniform facets-per-cell through constant divisors)
template <int PER>
__device__ __forceinline__ void divmod_per(uint32_t u, uint32_t& q, uint32_t& r) {
  q = u / (uint32_t)PER;
  r = u - q * (uint32_t)PER;
}

template <int D>
__global__ void __launch_bounds__(256)
k_facet_rows_count(const GcGrid g, const int8_t* __restrict__ row_state, int which, int mode, int side,
                   unsigned long long* __restrict__ dir_counts, int nseg_per_row, int64_t nsegs) {
  __shared__ uint32_t s_cnt[3];
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * 8 + warp;
  uint32_t cnt[3] = {0u, 0u, 0u};
  if (gw < nsegs) {
    const int row = (int)(gw / nseg_per_row), seg = (int)(gw % nseg_per_row);
    const int j = D == 3 ? row % g.n[1] : row;
    const int k = D == 3 ? row / g.n[1] : 0;
    const int per = D + (j == 0 ? 1 : 0) + ((D == 3 && k == 0) ? 1 : 0);
    const int64_t rowbase = facet_row_base<D>(g, j, k);
    const int nx = g.n[0];
    const int i_beg = seg * FR_SEG, i_end = min(nx, i_beg + FR_SEG);
    // full local sequence of a cell: [z-low], z-high, [y-low], y-high, x-low, x-high (per + 1 entries); 3 bits each:
    // direction (2) | side (1)
    uint32_t code = 0u;
    {
      int t = 0;
      for (int d = D - 1; d >= 0; --d) {
        const int idx = d == 0 ? 0 : (d == 1 ? j : k);
        if (idx == 0) code |= (uint32_t)(d << 1) << (3 * t++);
        code |= (uint32_t)((d << 1) | 1) << (3 * t++);
      }
    }
    const bool row_b1 = j == g.n[1] - 1, row_b2 = D == 3 && k == g.n[2] - 1;
    const int64_t t_beg = (int64_t)i_beg * per + (i_beg > 0 ? 1 : 0), t_end = (int64_t)i_end * per + 1;
    const int64_t p_beg = rowbase + t_beg, p_end = rowbase + t_end;
    for (int64_t blk = (p_beg & ~(int64_t)15) + lane * 16; blk < p_end; blk += 512) {
      const uint4 v = *reinterpret_cast<const uint4*>(row_state + blk);
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        const int64_t p = blk + b;
        if (p < p_beg || p >= p_end) continue;
        const int a = (int)(int8_t)((wv[b >> 2] >> (8 * (b & 3))) & 0xffu);
        if (!(mode == 1 ? a == side : (a == GC_CUT || a == side))) continue;
        const uint32_t t = (uint32_t)(p - rowbase);
        uint32_t i = 0u, sq = t;  // cell, index in the full sequence
        if (t > (uint32_t)per) {
          uint32_t loc;
          switch (per) {
            case 2: divmod_per<2>(t - 1u, i, loc); break;
            case 3: divmod_per<3>(t - 1u, i, loc); break;
            case 4: divmod_per<4>(t - 1u, i, loc); break;
            default: divmod_per<5>(t - 1u, i, loc); break;
          }
          sq = loc < (uint32_t)per - 1u ? loc : (uint32_t)per;  // cells i > 0 have no x-low facet
        }
        const uint32_t c = (code >> (3u * sq)) & 7u;
        const int d = (int)(c >> 1);
        const bool isb = (c & 1u) == 0u || (d == 0 ? (int)i == nx - 1 : (d == 1 ? row_b1 : row_b2));
        if (isb == (which == GECUT_FACETS_BOUNDARY)) cnt[d]++;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    uint32_t v = cnt[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(&s_cnt[d], v);
  }
  __syncthreads();
  if (threadIdx.x < D && s_cnt[threadIdx.x]) atomicAdd(&dir_counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// ids of the selected whole facets, materialised only when a caller asks for them
template <int D>
__global__ void k_facet_flag_bg(const GcGrid g, const int8_t* __restrict__ rows, int64_t pitch, int64_t nf,
                                const GcProgram prog, int which, int mode, int side, uint32_t* __restrict__ flags) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const FacetIdx fi = facet_decode<D>(g, (uint32_t)f);
  const bool in_trian = facet_is_boundary<D>(g, fi) == (which == GECUT_FACETS_BOUNDARY);
  const int a = gc_eval_inoutcut(prog, rows, pitch, f);
  flags[f] = (in_trian && (mode == 1 ? a == side : (a == GC_CUT || a == side))) ? 1u : 0u;
}

// sub-facet e belongs to BoundaryTriangulation / SkeletonTriangulation(cut_facets, CutInOrOut(side), geo)
// (EmbeddedFacetDiscretizations.jl:167-190): bg facet CUT, sub-facet on `side`, bg facet in the triangulation
__device__ __forceinline__ bool facet_sub_selected(const GcFacetLayout& lay, int L, const GcProgram& prog, int which, int side,
                                                   int64_t e) {
  if ((lay.isb[e] != 0) != (which == GECUT_FACETS_BOUNDARY)) return false;
  if (gc_eval_inoutcut(prog, lay.state, lay.pitch, e) != side) return false;
  // one level set: a sub-facet only exists in a facet that level set cuts
  return L == 1 || gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, (int64_t)lay.c2bg[e] - 1) == GC_CUT;
}

__global__ void k_facet_flag_sub(const GcFacetLayout lay, int L, int64_t nsf, const GcProgram prog, int which, int side,
                                 uint32_t* __restrict__ flags) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nsf) return;
  flags[e] = facet_sub_selected(lay, L, prog, which, side, e) ? 1u : 0u;
}

// number and measure of the selected sub-facets in one pass (the id list is not needed for either): fixed element ->
// thread -> tree order inside a block, the block that finishes last adds the block partials in index order
__global__ void __launch_bounds__(256)
k_facet_sub_count_sum(const GcFacetLayout lay, int L, int64_t nsf, const GcProgram prog, int which, int side,
                      double* __restrict__ partial, unsigned long long* __restrict__ pcount, unsigned int* __restrict__ ticket,
                      double* __restrict__ out_sum, unsigned long long* __restrict__ out_count) {
  __shared__ double sh[256];
  __shared__ unsigned long long shc[256];
  __shared__ bool s_last;
  double a = 0.0;
  unsigned long long c = 0ull;
  const int64_t per = (nsf + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi = lo + per < nsf ? lo + per : nsf;
  for (int64_t e = lo + threadIdx.x; e < hi; e += blockDim.x)
    if (facet_sub_selected(lay, L, prog, which, side, e)) {
      a += lay.meas[e];
      c++;
    }
  sh[threadIdx.x] = a;
  shc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[threadIdx.x] += sh[threadIdx.x + o];
      shc[threadIdx.x] += shc[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = sh[0];
    pcount[blockIdx.x] = shc[0];
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  a = 0.0;
  c = 0ull;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    a += reinterpret_cast<volatile double*>(partial)[i];
    c += reinterpret_cast<volatile unsigned long long*>(pcount)[i];
  }
  sh[threadIdx.x] = a;
  shc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[threadIdx.x] += sh[threadIdx.x + o];
      shc[threadIdx.x] += shc[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *out_sum = sh[0];
    *out_count = shc[0];
  }
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

// ------------------------------------------------------------------------------------------------
// simplexified models: a thread per facet id (the slot table decodes it), lanes = consecutive ids
// ------------------------------------------------------------------------------------------------
// states of every facet per level set from the node sign planes of K1 + "cut by any level set"; EMIT = false records
// the number of cut facets of every 32-facet piece, EMIT = true (after the scan) writes their ids in ascending order
template <int D, bool EMIT>
__global__ void __launch_bounds__(256)
k_sx_facet_states(const GcGrid g, const GcSxFacets sx, int L, int wxt, const uint32_t* __restrict__ signs, int64_t ls_stride,
                  int8_t* __restrict__ state, int64_t pitch, uint32_t* __restrict__ piece_counts,
                  int32_t* __restrict__ cutfacets, int64_t nf) {
  const int64_t f = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = f < nf;
  bool cut = false;
  if (valid) {
    const SxFacetIdx fi = sx_facet_decode<D>(g, sx, (uint32_t)f);
    int64_t word[D];
    int bit[D];
#pragma unroll
    for (int m = 0; m < D; ++m) {
      int nijk[3];
      sx_facet_node<D>(fi, sx, m, nijk);
      const int64_t nrow = D == 3 ? (int64_t)nijk[2] * g.nn[1] + nijk[1] : (int64_t)nijk[1];
      word[m] = nrow * wxt + (nijk[0] >> 5);
      bit[m] = nijk[0] & 31;
    }
    for (int ls = 0; ls < L; ++ls) {
      const uint32_t* sg = signs + (int64_t)ls * ls_stride;
      int nout = 0;
#pragma unroll
      for (int m = 0; m < D; ++m) nout += (int)((__ldg(sg + word[m]) >> bit[m]) & 1u);
      const int8_t st = nout == 0 ? (int8_t)GC_IN : (nout == D ? (int8_t)GC_OUT : (int8_t)GC_CUT);
      if (!EMIT) state[(int64_t)ls * pitch + f] = st;
      cut = cut || st == GC_CUT;
    }
  }
  const uint32_t ballot = __ballot_sync(0xffffffffu, cut);
  const int64_t piece = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (!EMIT) {
    if (lane == 0) piece_counts[piece] = (uint32_t)__popc(ballot);
  } else if (cut) {
    cutfacets[piece_counts[piece] + (uint32_t)__popc(ballot & ((1u << lane) - 1u))] = (int32_t)(f + 1);
  }
}

template <int D>
__global__ void k_sx_facet_topology(const GcGrid g, const GcSxFacets sx, int64_t nf, int32_t* __restrict__ nodes,
                                    int8_t* __restrict__ isb) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const SxFacetIdx fi = sx_facet_decode<D>(g, sx, (uint32_t)f);
  if (isb) isb[f] = sx_facet_is_boundary<D>(g, fi, sx) ? 1 : 0;
  if (nodes) {
#pragma unroll
    for (int m = 0; m < D; ++m) {
      int nijk[3];
      sx_facet_node<D>(fi, sx, m, nijk);
      const int64_t node = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                                    : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
      nodes[f * D + m] = (int32_t)(node + 1);
    }
  }
}

// whole facets of a selection, counted per slot of the cube (facets of one slot are congruent: one measure per slot);
// `row` is ONE Int8 state row.  flags != nullptr: the per-facet flags for the id list instead of the counts.
template <int D>
__global__ void __launch_bounds__(256)
k_sx_facet_select(const GcGrid g, const GcSxFacets sx, const int8_t* __restrict__ row, int64_t nf, int which, int mode,
                  int side, unsigned long long* __restrict__ slot_counts, uint32_t* __restrict__ flags) {
  __shared__ uint32_t s_cnt[24];
  if (threadIdx.x < 24) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t f = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (f < nf) {
    const SxFacetIdx fi = sx_facet_decode<D>(g, sx, (uint32_t)f);
    const bool in_trian = sx_facet_is_boundary<D>(g, fi, sx) == (which == GECUT_FACETS_BOUNDARY);
    const int a = row[f];
    const bool sel = in_trian && (mode == 1 ? a == side : (a == GC_CUT || a == side));
    if (flags) flags[f] = sel ? 1u : 0u;
    if (sel && slot_counts) atomicAdd(&s_cnt[fi.slot], 1u);
  }
  __syncthreads();
  if (slot_counts && threadIdx.x < 24 && s_cnt[threadIdx.x]) atomicAdd(&slot_counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
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
  gc_free(ctx, l.isb);
  for (int i = 0; i < GC_LMAX; ++i)
    if (r->owns_nodal[i]) gc_free(ctx, r->nodal_dev[i]);
  l = GcFacetLayout{};
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
  const int L = r->L;
  const int wxt = (g.nn[0] + 31) / 32;
  const int64_t node_rows = D == 3 ? (int64_t)g.nn[1] * g.nn[2] : (int64_t)g.nn[1];
  const int64_t ls_stride = node_rows * wxt;
  const int64_t nrows = D == 3 ? (int64_t)g.n[1] * g.n[2] : (int64_t)g.n[1];
  const int wpr = (int)gc_div_up(g.n[0], 32);  // cell words per row
  const int64_t nitems = nrows * wpr;
  const int64_t npieces = gc_div_up(nitems, 32);
  if (nitems >= 4294967295ll) {
    ctx->err = "cut_facets: model too large";
    return GECUT_ERR_OVERFLOW;
  }
  dim3 sgrid((unsigned)wxt, 1u, 1u), sblock(32, 8);
  if (D == 3) {
    sgrid.y = (unsigned)gc_div_up(g.nn[1], 8);
    sgrid.z = (unsigned)gc_div_up(g.nn[2], NS_KCH);
  } else {
    sgrid.y = (unsigned)gc_div_up(gc_div_up(g.nn[1], NS_KCH), 8);
  }
  if (sgrid.y > 65535u || sgrid.z > 65535u) {
    ctx->err = "cut_facets: model too large for the node-sign grid";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  GcScratch scratch(ctx);  // released on every exit path
  uint32_t* signs = nullptr;
  uint32_t* seg_counts = nullptr;
  uint64_t* total_dev = nullptr;
  GC_CHECK(scratch.alloc(&signs, sizeof(uint32_t) * ls_stride * L));
  GC_CHECK(scratch.alloc(&seg_counts, sizeof(uint32_t) * npieces));
  GC_CHECK(scratch.alloc(&total_dev, sizeof(uint64_t) * (1 + GC_LMAX * FT_N)));
  GC_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(uint64_t) * (1 + GC_LMAX * FT_N), ctx->stream));
  GC_LAUNCH(ctx, "k_node_signs", k_node_signs<D><<<sgrid, sblock, 0, ctx->stream>>>(g, r->leaves, wxt, signs, ls_stride));
  GC_CHECK(gc_launch_check(ctx, "k_node_signs"));
  if (g.simplexified) {
    const int64_t nf = r->sizes.num_bgfacets;
    const unsigned nbs = (unsigned)gc_div_up(nf, 256);
    uint32_t* pieces = nullptr;
    GC_CHECK(scratch.alloc(&pieces, sizeof(uint32_t) * (size_t)nbs * 8));
    GC_LAUNCH(ctx, "k_sx_facet_states",
              (k_sx_facet_states<D, false><<<nbs, 256, 0, ctx->stream>>>(g, r->sx, L, wxt, signs, ls_stride, r->lay.bg_state,
                                                                        r->lay.bg_pitch, pieces, nullptr, nf)));
    GC_CHECK(gc_launch_check(ctx, "k_sx_facet_states"));
    GC_CHECK(gc_exclusive_scan_u32(ctx, pieces, pieces, (int64_t)nbs * 8, total_dev));
    uint64_t total = 0;
    GC_CHECK(read_u64(ctx, total_dev, 1, &total));
    r->sizes.num_cutfacets = (int64_t)total;
    if (total > 0) {
      GC_CHECK(gc_malloc(ctx, (void**)&r->lay.cutfacets, sizeof(int32_t) * total));
      GC_LAUNCH(ctx, "k_sx_facet_states_emit",
                (k_sx_facet_states<D, true><<<nbs, 256, 0, ctx->stream>>>(g, r->sx, L, wxt, signs, ls_stride, nullptr, 0, pieces,
                                                                         r->lay.cutfacets, nf)));
      GC_CHECK(gc_launch_check(ctx, "k_sx_facet_states_emit"));
    }
    return GECUT_OK;
  }
  const unsigned nb = (unsigned)gc_div_up(gc_div_up(npieces, FW_ITER), 8);
  GC_LAUNCH(ctx, "k_facet_words",
            (k_facet_words<D, false><<<nb, 256, 0, ctx->stream>>>(g, L, wxt, wpr, signs, ls_stride, r->lay.bg_state,
                                                                r->lay.bg_pitch, seg_counts, nullptr, nitems, npieces,
                                                                (unsigned long long*)(total_dev + 1))));
  GC_CHECK(gc_launch_check(ctx, "k_facet_words"));
  GC_CHECK(gc_exclusive_scan_u32(ctx, seg_counts, seg_counts, npieces, total_dev));
  uint64_t hbuf[1 + GC_LMAX * FT_N];
  GC_CHECK(read_u64(ctx, total_dev, 1 + L * FT_N, hbuf));
  const uint64_t total = hbuf[0];
  for (int c = 0; c < L * FT_N; ++c) r->tallies[c] = (int64_t)hbuf[1 + c];
  r->sizes.num_cutfacets = (int64_t)total;
  if (total > 0) {
    GC_CHECK(gc_malloc(ctx, (void**)&r->lay.cutfacets, sizeof(int32_t) * total));
    GC_LAUNCH(ctx, "k_facet_words_emit",
              (k_facet_words<D, true><<<nb, 256, 0, ctx->stream>>>(g, L, wxt, wpr, signs, ls_stride, nullptr, 0, seg_counts,
                                                                 r->lay.cutfacets, nitems, npieces, nullptr)));
    GC_CHECK(gc_launch_check(ctx, "k_facet_words_emit"));
  }
  return GECUT_OK;
}

template <int D, int NW, bool SX>
int walk_facets_nw(gecut_ctx* ctx, gecut_facet_result* r) {
  constexpr int NT0 = (D == 3 && !SX) ? 2 : 1;
  const int64_t nsimp = r->sizes.num_cutfacets * NT0;
  const int L = r->L;
  GcScratch scratch(ctx);
  uint32_t* cnt = nullptr;
  uint64_t* totals_dev = nullptr;
  GC_CHECK(scratch.alloc(&cnt, sizeof(uint32_t) * 2 * nsimp));
  GC_CHECK(scratch.alloc(&totals_dev, sizeof(uint64_t) * 2));
  const unsigned nb = (unsigned)gc_div_up(nsimp, FW_T);
  GC_LAUNCH(ctx, "k_facet_walk_count",
            (k_facet_walk<D, NW, false, SX><<<nb, FW_T, 0, ctx->stream>>>(r->grid, r->leaves, ctx->d_tables, r->lay.cutfacets,
                                                                         nsimp, cnt, r->lay, r->sx)));
  GC_CHECK(gc_launch_check(ctx, "k_facet_walk_count"));
  GC_CHECK(gc_exclusive_scan_u32_multi(ctx, cnt, cnt, nsimp, 2, totals_dev));
  uint64_t tot[2] = {0, 0};
  GC_CHECK(read_u64(ctx, totals_dev, 2, tot));
  if (tot[0] >= 2147483647ull || tot[1] >= 2147483647ull) {
    ctx->err = "sub-facet ids overflow Int32";
    return GECUT_ERR_OVERFLOW;
  }
  r->sizes.num_subfacets = (int64_t)tot[0];
  r->sizes.num_subfacet_points = (int64_t)tot[1];
  GcFacetLayout& l = r->lay;
  l.pitch = pad_pitch((int64_t)tot[0]);
  GC_CHECK(gc_malloc(ctx, (void**)&l.c2p, sizeof(int32_t) * tot[0] * D));
  GC_CHECK(gc_malloc(ctx, (void**)&l.c2bg, sizeof(int32_t) * tot[0]));
  GC_CHECK(gc_malloc(ctx, (void**)&l.meas, sizeof(double) * tot[0]));
  GC_CHECK(gc_malloc(ctx, (void**)&l.isb, (size_t)tot[0]));
  GC_CHECK(gc_malloc(ctx, (void**)&l.state, (size_t)l.pitch * L));
  GC_CHECK(gc_malloc(ctx, (void**)&l.X, sizeof(double) * tot[1] * D));
  GC_CHECK(gc_malloc(ctx, (void**)&l.R, sizeof(double) * tot[1] * (D - 1)));
  GC_LAUNCH(ctx, "k_facet_walk_fill",
            (k_facet_walk<D, NW, true, SX><<<nb, FW_T, 0, ctx->stream>>>(r->grid, r->leaves, ctx->d_tables, r->lay.cutfacets,
                                                                        nsimp, cnt, r->lay, r->sx)));
  GC_CHECK(gc_launch_check(ctx, "k_facet_walk_fill"));
  return GECUT_OK;
}

template <int D, bool SX>
int walk_facets_sx(gecut_ctx* ctx, gecut_facet_result* r) {
  const int L = r->L;
  if (L <= 1) return walk_facets_nw<D, 1, SX>(ctx, r);
  if (L <= 2) return walk_facets_nw<D, 2, SX>(ctx, r);
  if (L <= 4) return walk_facets_nw<D, 4, SX>(ctx, r);
  if (L <= 8) return walk_facets_nw<D, 8, SX>(ctx, r);
  return walk_facets_nw<D, GC_LMAX, SX>(ctx, r);
}
template <int D>
int walk_facets(gecut_ctx* ctx, gecut_facet_result* r) {
  if (r->sizes.num_cutfacets == 0) return GECUT_OK;
  return r->grid.simplexified ? walk_facets_sx<D, true>(ctx, r) : walk_facets_sx<D, false>(ctx, r);
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
  if (r->grid.simplexified && !gc_build_sx_facets(r->grid.D, &r->sx)) {
    ctx->err = "cut_facets: the simplexify table of the n-cube does not tile conformingly";
    delete r;
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  {
    // z-slab: the facet grid of the rank's LOCAL model (layers [k0, k1) of the last direction as a Cartesian model of its own:
    // local numbering, the slab's lower and upper planes are boundary facets of the local model); k0 stays as the shift of
    // the node coordinates (fc_coord)
    GcGrid& gg = r->grid;
    const int Dg = gg.D, nloc = gg.k1 - gg.k0;
    gg.n[Dg - 1] = nloc;
    gg.nn[Dg - 1] = nloc + 1;
    gg.k1 = gg.k0 + nloc;
  }
  const GcGrid& g = r->grid;
  auto fail = [&](int code) {
    free_facet_arrays(r);
    cudaStreamSynchronize(ctx->stream);
    delete r;
    return code;
  };
  const int D = g.D;
  const int64_t nf = g.simplexified ? (D == 2 ? sx_num_facets<2>(g, r->sx) : sx_num_facets<3>(g, r->sx))
                                    : (D == 2 ? num_facets<2>(g) : num_facets<3>(g));
  if (nf >= 2147483647ll) {
    ctx->err = "facet ids overflow Int32";
    return fail(GECUT_ERR_OVERFLOW);
  }
  r->leaves.L = L;
  r->leaves.nodal_base = 0;
  const int64_t nn_total = g.layer_nodes * g.nn[D - 1];  // nodes of the (local) model
  // NODAL vectors: of the whole model (skip the planes below the slab) or slab-local (they start at the slab's first plane)
  const int64_t nodal_skip = grid->nodal_slab_local ? 0 : (int64_t)g.k0 * g.layer_nodes;
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
        r->leaves.nodal[i] = nodal_values[i] + nodal_skip;
      } else {
        double* d = nullptr;
        st = gc_malloc(ctx, (void**)&d, sizeof(double) * nn_total);
        if (st != GECUT_OK) return fail(st);
        r->nodal_dev[i] = d;
        r->owns_nodal[i] = true;
        if (cudaMemcpyAsync(d, nodal_values[i] + nodal_skip, sizeof(double) * nn_total, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
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
  sz.num_vertices_per_bgfacet = g.simplexified ? D : 1 << (D - 1);
  // facets with one cell around: the two faces of the box per direction (cd simplex facets per cube face when simplexified)
  if (D == 2)
    sz.num_boundary_facets = 2ll * (g.n[0] + g.n[1]);
  else
    sz.num_boundary_facets = 2ll * ((int64_t)g.n[0] * g.n[1] + (int64_t)g.n[0] * g.n[2] + (int64_t)g.n[1] * g.n[2]);
  if (g.simplexified) sz.num_boundary_facets *= r->sx.cd;
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

// Facet numbering on the host: the index arithmetic of the kernels (facet_decode / sx_facet_decode are __host__ __device__),
// no device work.  For host programs that need the nodes of a few facets, and for the CPU tests that pin the closed forms
// to the oracle's first-encounter numbering.
int gecut_facet_topology_host(const gecut_grid* grid, int64_t first, int64_t count, int32_t* facet_to_nodes,
                              int8_t* is_boundary_out, int64_t* num_facets_out, int32_t* nodes_per_facet) {
  if (!grid || (grid->D != 2 && grid->D != 3)) return GECUT_ERR_INVALID;
  const int D = grid->D;
  GcGrid g{};
  g.D = D;
  g.simplexified = grid->simplexified ? 1 : 0;
  for (int d = 0; d < 3; ++d) {
    g.n[d] = d < D ? grid->partition[d] : 1;
    if (g.n[d] < 1) return GECUT_ERR_INVALID;
  }
  if (grid->slab_begin != 0 || grid->slab_end != 0) {  // the local model of the slab
    if (grid->slab_begin < 0 || grid->slab_end <= grid->slab_begin || grid->slab_end > grid->partition[D - 1]) return GECUT_ERR_INVALID;
    g.n[D - 1] = grid->slab_end - grid->slab_begin;
  }
  for (int d = 0; d < 3; ++d) g.nn[d] = d < D ? g.n[d] + 1 : 1;
  GcSxFacets sx{};
  if (g.simplexified && !gc_build_sx_facets(D, &sx)) return GECUT_ERR_NOTIMPLEMENTED;
  const int64_t nf = g.simplexified ? (D == 2 ? sx_num_facets<2>(g, sx) : sx_num_facets<3>(g, sx))
                                    : (D == 2 ? num_facets<2>(g) : num_facets<3>(g));
  const int nvf = g.simplexified ? D : 1 << (D - 1);
  if (num_facets_out) *num_facets_out = nf;
  if (nodes_per_facet) *nodes_per_facet = nvf;
  if (count == 0) return GECUT_OK;
  if (first < 0 || count < 0 || first + count > nf || nf >= 2147483647ll) return GECUT_ERR_INVALID;
  for (int64_t q = 0; q < count; ++q) {
    const uint32_t f = (uint32_t)(first + q);
    bool isb;
    int nijk[4][3];
    if (g.simplexified) {
      const SxFacetIdx fi = D == 2 ? sx_facet_decode<2>(g, sx, f) : sx_facet_decode<3>(g, sx, f);
      isb = D == 2 ? sx_facet_is_boundary<2>(g, fi, sx) : sx_facet_is_boundary<3>(g, fi, sx);
      for (int m = 0; m < nvf; ++m) {
        if (D == 2) sx_facet_node<2>(fi, sx, m, nijk[m]); else sx_facet_node<3>(fi, sx, m, nijk[m]);
      }
    } else {
      const FacetIdx fi = D == 2 ? facet_decode<2>(g, f) : facet_decode<3>(g, f);
      isb = D == 2 ? facet_is_boundary<2>(g, fi) : facet_is_boundary<3>(g, fi);
      for (int m = 0; m < nvf; ++m) {
        if (D == 2) facet_node<2>(fi, m, nijk[m]); else facet_node<3>(fi, m, nijk[m]);
      }
    }
    if (is_boundary_out) is_boundary_out[q] = isb ? 1 : 0;
    if (facet_to_nodes)
      for (int m = 0; m < nvf; ++m) {
        const int64_t node = (int64_t)nijk[m][0] + (int64_t)g.nn[0] * ((int64_t)nijk[m][1] + (D == 3 ? (int64_t)g.nn[1] * nijk[m][2] : 0));
        facet_to_nodes[q * nvf + m] = (int32_t)(node + 1);
      }
  }
  return GECUT_OK;
}

int gecut_facet_result_free(gecut_facet_result* res) {
  if (!res) return GECUT_OK;
  if (res->live_selections > 0) {  // deferred: the last selection of this result releases it
    res->free_requested = true;
    return GECUT_OK;
  }
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
  GcScratch scratch(ctx);  // released on every exit path
  int32_t* nodes = nullptr;
  int8_t* isb = nullptr;
  if (host->facet_to_nodes || host->facet_is_boundary) {
    const int64_t nf = s.num_bgfacets;
    const int nvf = s.num_vertices_per_bgfacet;
    if (host->facet_to_nodes) GC_CHECK(scratch.alloc(&nodes, sizeof(int32_t) * nf * nvf));
    if (host->facet_is_boundary) GC_CHECK(scratch.alloc(&isb, (size_t)nf));
    const int T = 256;
    if (res->grid.simplexified) {
      if (D == 2)
        GC_LAUNCH(ctx, "k_facet_topology", k_sx_facet_topology<2><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, res->sx, nf, nodes, isb));
      else
        GC_LAUNCH(ctx, "k_facet_topology", k_sx_facet_topology<3><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, res->sx, nf, nodes, isb));
    } else if (D == 2)
      GC_LAUNCH(ctx, "k_facet_topology", k_facet_topology<2><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, nf, nodes, isb));
    else
      GC_LAUNCH(ctx, "k_facet_topology", k_facet_topology<3><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, nf, nodes, isb));
    GC_CHECK(gc_launch_check(ctx, "k_facet_topology"));
    GC_CHECK(to_host(ctx, host->facet_to_nodes, nodes, nf * nvf));
    GC_CHECK(to_host(ctx, host->facet_is_boundary, isb, nf));
  }
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
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
  if (res->free_requested) {
    ctx->err = "the facet result has been freed";
    return GECUT_ERR_INVALID;
  }
  gecut_facet_selection* sel = new (std::nothrow) gecut_facet_selection();
  if (!sel) return GECUT_ERR_OOM;
  sel->res = res;
  res->live_selections++;
  sel->which = which;
  sel->kind = kind;
  sel->side = side;
  sel->prog = p;
  auto fail = [&](int code) {
    gc_free(ctx, sel->sub_ids);
    gc_free(ctx, sel->bg_ids);
    delete sel;
    res->live_selections--;
    return code;
  };
  const int D = res->sizes.D;
  const int T = 256;
  int st = GECUT_OK;
  const int64_t nsf = res->sizes.num_subfacets, nf = res->sizes.num_bgfacets;
  if ((kind == GECUT_SEL_PHYSICAL || kind == GECUT_SEL_CUTONLY) && nsf > 0) {
    // count and measure in one pass; the id list is materialised by gecut_facet_selection_copy / per-entity values
    const int NB = (int)std::min<int64_t>(1184, gc_div_up(nsf, 1024));
    double* partial = nullptr;
    st = gc_malloc(ctx, (void**)&partial, sizeof(double) * (2 * NB + 3));
    if (st != GECUT_OK) return fail(st);
    unsigned long long* pcount = reinterpret_cast<unsigned long long*>(partial + NB);
    double* out_sum = partial + 2 * NB;
    unsigned long long* out_count = reinterpret_cast<unsigned long long*>(partial + 2 * NB + 1);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + 2 * NB + 2);
    cudaMemsetAsync(ticket, 0, sizeof(unsigned int), ctx->stream);
    GC_LAUNCH(ctx, "k_facet_sub_count_sum", k_facet_sub_count_sum<<<NB, 256, 0, ctx->stream>>>(res->lay, res->L, nsf, p, which, side, partial, pcount, ticket, out_sum, out_count));
    st = gc_launch_check(ctx, "k_facet_sub_count_sum");
    if (st == GECUT_OK) {
      uint64_t h[2];
      st = read_u64(ctx, (const uint64_t*)out_sum, 2, h);
      memcpy(&sel->sub_measure, &h[0], sizeof(double));
      sel->n_sub = (int64_t)h[1];
      sel->sub_lazy = true;
    }
    gc_free(ctx, partial);
    if (st != GECUT_OK) return fail(st);
  }
  if (kind != GECUT_SEL_CUTONLY && nf > 0) {
    const int mode = kind == GECUT_SEL_ACTIVE ? 2 : 1;
    const GcGrid& g = res->grid;
    const bool single = p.len == 1 || (p.len == 2 && p.tok[1] == GECUT_OP_COMPLEMENT);
    if (g.simplexified) {
      // the program's Int8 row, then one pass: selected whole facets per slot of the cube
      GcScratch scratch(ctx);
      unsigned long long* slotc = nullptr;
      int8_t* tmp = nullptr;
      st = scratch.alloc(&slotc, sizeof(unsigned long long) * 24);
      if (st == GECUT_OK) st = scratch.alloc(&tmp, (size_t)pad_pitch(nf));
      if (st == GECUT_OK) st = gc_eval_rows(ctx, res->lay.bg_state, res->lay.bg_pitch, nf, p, tmp, true);
      if (st != GECUT_OK) return fail(st);
      cudaMemsetAsync(slotc, 0, sizeof(unsigned long long) * 24, ctx->stream);
      const unsigned nbs = (unsigned)gc_div_up(nf, 256);
      if (D == 2)
        GC_LAUNCH(ctx, "k_sx_facet_select", k_sx_facet_select<2><<<nbs, 256, 0, ctx->stream>>>(g, res->sx, tmp, nf, which, mode, side, slotc, nullptr));
      else
        GC_LAUNCH(ctx, "k_sx_facet_select", k_sx_facet_select<3><<<nbs, 256, 0, ctx->stream>>>(g, res->sx, tmp, nf, which, mode, side, slotc, nullptr));
      st = gc_launch_check(ctx, "k_sx_facet_select");
      uint64_t h[24];
      if (st == GECUT_OK) st = read_u64(ctx, (const uint64_t*)slotc, 24, h);
      if (st != GECUT_OK) return fail(st);
      for (int q = 0; q < 24; ++q) {
        sel->bg_slot_count[q] = (int64_t)h[q];
        sel->n_bg += (int64_t)h[q];
      }
    } else if (single) {
      // single leaf (or its complement = the other side): the tallies of the classification pass answer directly
      const int ls = p.tok[0];
      const int eff_side = p.len == 1 ? side : -side;
      const int isb = which == GECUT_FACETS_BOUNDARY ? 1 : 0;
      const int64_t* tl = res->tallies + (size_t)ls * 12;
      for (int d = 0; d < D; ++d) {
        int64_t others = 1;
        for (int e = 0; e < D; ++e)
          if (e != d) others *= g.n[e];
        const int64_t all = isb ? 2 * others : (int64_t)(g.n[d] - 1) * others;
        const int s_same = eff_side == GECUT_IN ? 0 : 1;
        const int64_t c = mode == 1 ? tl[(s_same * 3 + d) * 2 + isb] : all - tl[((1 - s_same) * 3 + d) * 2 + isb];
        sel->bg_dir_count[d] = c;
        sel->n_bg += c;
      }
    } else {
      unsigned long long* dirc = nullptr;
      st = gc_malloc(ctx, (void**)&dirc, sizeof(unsigned long long) * 3);
      if (st != GECUT_OK) return fail(st);
      cudaMemsetAsync(dirc, 0, sizeof(unsigned long long) * 3, ctx->stream);
      const int64_t nrows = D == 3 ? (int64_t)g.n[1] * g.n[2] : (int64_t)g.n[1];
      const int nseg_per_row = (int)gc_div_up(g.n[0], FR_SEG);
      const int64_t nsegs = nrows * nseg_per_row;
      // the program's own Int8 row first (truth-table gather over the level-set rows), then one pass over that row
      int8_t* tmp = nullptr;
      st = gc_malloc(ctx, (void**)&tmp, (size_t)pad_pitch(nf));
      if (st == GECUT_OK) st = gc_eval_rows(ctx, res->lay.bg_state, res->lay.bg_pitch, nf, p, tmp, true);
      if (st != GECUT_OK) {
        gc_free(ctx, tmp);
        gc_free(ctx, dirc);
        return fail(st);
      }
      const unsigned nb = (unsigned)gc_div_up(nsegs, 8);
      if (D == 2)
        GC_LAUNCH(ctx, "k_facet_rows_count", k_facet_rows_count<2><<<nb, 256, 0, ctx->stream>>>(g, tmp, which, mode, side, dirc, nseg_per_row, nsegs));
      else
        GC_LAUNCH(ctx, "k_facet_rows_count", k_facet_rows_count<3><<<nb, 256, 0, ctx->stream>>>(g, tmp, which, mode, side, dirc, nseg_per_row, nsegs));
      gc_free(ctx, tmp);
      st = gc_launch_check(ctx, "k_facet_rows_count");
      if (st == GECUT_OK) {
        uint64_t h[3];
        st = read_u64(ctx, (const uint64_t*)dirc, 3, h);
        for (int d = 0; d < 3; ++d) {
          sel->bg_dir_count[d] = (int64_t)h[d];
          sel->n_bg += (int64_t)h[d];
        }
      }
      gc_free(ctx, dirc);
      if (st != GECUT_OK) return fail(st);
    }
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
  gecut_facet_result* res = const_cast<gecut_facet_result*>(sel->res);
  delete sel;
  if (--res->live_selections == 0 && res->free_requested) return gecut_facet_result_free(res);
  return GECUT_OK;
}

int gecut_facet_selection_sizes(const gecut_facet_selection* sel, int64_t* n_sub, int64_t* n_bg) {
  if (!sel) return GECUT_ERR_INVALID;
  if (n_sub) *n_sub = sel->n_sub;
  if (n_bg) *n_bg = sel->n_bg;
  return GECUT_OK;
}

static int facet_materialize_sub_ids(gecut_ctx* ctx, gecut_facet_selection* sel) {
  if (!sel->sub_lazy || sel->sub_ids || sel->n_sub <= 0) return GECUT_OK;
  const gecut_facet_result* res = sel->res;
  const int64_t nsf = res->sizes.num_subfacets;
  const int T = 256;
  uint32_t* flags = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&flags, sizeof(uint32_t) * nsf));
  GC_LAUNCH(ctx, "k_facet_flag_sub", k_facet_flag_sub<<<(unsigned)gc_div_up(nsf, T), T, 0, ctx->stream>>>(res->lay, res->L, nsf, sel->prog, sel->which, sel->side, flags));
  int st = gc_launch_check(ctx, "k_facet_flag_sub");
  int64_t n = 0;
  if (st == GECUT_OK) st = gc_compact_flags(ctx, flags, nsf, &sel->sub_ids, &n, nullptr, nullptr);
  gc_free(ctx, flags);
  GC_CHECK(st);
  if (n != sel->n_sub) {
    ctx->err = "facet selection: sub-facet id list and count disagree";
    return GECUT_ERR_CUDA;
  }
  return GECUT_OK;
}

int gecut_facet_selection_copy(const gecut_facet_selection* sel_c, int32_t* sub_ids, int32_t* bg_ids) {
  if (!sel_c) return GECUT_ERR_INVALID;
  gecut_facet_selection* sel = const_cast<gecut_facet_selection*>(sel_c);
  const gecut_facet_result* res = sel->res;
  gecut_ctx* ctx = res->ctx;
  DeviceGuard guard(ctx->device);
  if (bg_ids && sel->n_bg > 0 && !sel->bg_ids) {
    // ids of the whole facets: flags over all facets -> scan -> scatter (only when somebody wants the list)
    const int64_t nf = res->sizes.num_bgfacets;
    const int mode = sel->kind == GECUT_SEL_ACTIVE ? 2 : 1;
    const int T = 256;
    uint32_t* flags = nullptr;
    GC_CHECK(gc_malloc(ctx, (void**)&flags, sizeof(uint32_t) * nf));
    if (res->grid.simplexified) {
      int8_t* tmp = nullptr;
      int st0 = gc_malloc(ctx, (void**)&tmp, (size_t)pad_pitch(nf));
      if (st0 == GECUT_OK) st0 = gc_eval_rows(ctx, res->lay.bg_state, res->lay.bg_pitch, nf, sel->prog, tmp, true);
      if (st0 == GECUT_OK) {
        const unsigned nbs = (unsigned)gc_div_up(nf, 256);
        if (res->sizes.D == 2)
          GC_LAUNCH(ctx, "k_sx_facet_select", k_sx_facet_select<2><<<nbs, 256, 0, ctx->stream>>>(res->grid, res->sx, tmp, nf, sel->which, mode, sel->side, nullptr, flags));
        else
          GC_LAUNCH(ctx, "k_sx_facet_select", k_sx_facet_select<3><<<nbs, 256, 0, ctx->stream>>>(res->grid, res->sx, tmp, nf, sel->which, mode, sel->side, nullptr, flags));
      }
      gc_free(ctx, tmp);
      if (st0 != GECUT_OK) {
        gc_free(ctx, flags);
        return st0;
      }
    } else if (res->sizes.D == 2)
      GC_LAUNCH(ctx, "k_facet_flag_bg", k_facet_flag_bg<2><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, res->lay.bg_state, res->lay.bg_pitch, nf, sel->prog, sel->which, mode, sel->side, flags));
    else
      GC_LAUNCH(ctx, "k_facet_flag_bg", k_facet_flag_bg<3><<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(res->grid, res->lay.bg_state, res->lay.bg_pitch, nf, sel->prog, sel->which, mode, sel->side, flags));
    int st = gc_launch_check(ctx, "k_facet_flag_bg");
    int64_t n = 0;
    if (st == GECUT_OK) st = gc_compact_flags(ctx, flags, nf, &sel->bg_ids, &n, nullptr, nullptr);
    gc_free(ctx, flags);
    GC_CHECK(st);
    if (n != sel->n_bg) {
      ctx->err = "facet selection: id list and count disagree";
      return GECUT_ERR_CUDA;
    }
  }
  if (sub_ids) GC_CHECK(facet_materialize_sub_ids(ctx, sel));
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
  if (sel->n_sub > 0 && sel->sub_lazy && !sub_values) {
    sum = sel->sub_measure;  // summed by the selection pass
  } else if (sel->n_sub > 0) {
    GC_CHECK(facet_materialize_sub_ids(ctx, const_cast<gecut_facet_selection*>(sel)));
    const int64_t n = sel->n_sub;
    const int64_t nb = gc_div_up(n, (int64_t)SUM_T * SUM_PER);
    GcScratch scratch(ctx);
    double *partial = nullptr, *vals = nullptr;
    GC_CHECK(scratch.alloc(&partial, sizeof(double) * (nb + 1)));
    if (sub_values) GC_CHECK(scratch.alloc(&vals, sizeof(double) * n));
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
  }
  // whole facets: tensor Gauss rule of degree 2 on the (D-1)-cube, sum_q w_q * meas (the facet map is affine)
  const GcGrid& g = res->grid;
  const int D = g.D;
  if (g.simplexified) {
    // facets of one slot are translates of each other: measure of the SEGMENT / TRI from the cube-local vertex offsets,
    // degree-2 rule of the facet polytope (2-point Gauss on the segment, 3-point rule on the triangle)
    const GcSxFacets& t = res->sx;
    for (int q = 0; q < t.nslots; ++q) {
      if (sel->bg_slot_count[q] == 0) continue;
      double x[3][3];
      for (int m = 0; m < D; ++m)
        for (int d = 0; d < 3; ++d) x[m][d] = d < D ? (double)((t.vtx[q][m] >> d) & 1) * g.dx[d] : 0.0;
      double meas;
      if (D == 2) {
        const double a = x[1][0] - x[0][0], b = x[1][1] - x[0][1];
        meas = sqrt(a * a + b * b);
      } else {
        const double a[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
        const double b[3] = {x[2][0] - x[0][0], x[2][1] - x[0][1], x[2][2] - x[0][2]};
        const double n1 = a[1] * b[2] - a[2] * b[1], n2 = a[2] * b[0] - a[0] * b[2], n3 = a[0] * b[1] - a[1] * b[0];
        meas = sqrt(n1 * n1 + n2 * n2 + n3 * n3);
      }
      const int nq = D == 2 ? 2 : 3;
      const double w = D == 2 ? 0.5 : 1.0 / 6;
      double one = 0.0;
      for (int i = 0; i < nq; ++i) one += w * meas;
      sum += one * (double)sel->bg_slot_count[q];
    }
    if (total) *total = sum;
    return GECUT_OK;
  }
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
