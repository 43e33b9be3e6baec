// gecut_classify.cu -- K1+K2 fused: level-set evaluation at the background nodes and IN/OUT/CUT classification
// of every background cell, plus the stable compaction of the cut-cell list.
//
// Reference path replaced:
//   discretize(geo, grid)                 src/LevelSetCutters/DiscreteGeometries.jl:48-63   (L x Nn FP64 evaluations)
//   _compute_in_out_or_cut per level set  src/LevelSetCutters/CutTriangulations.jl:619-640  (compute_case, LookupTables.jl:19-33)
//   _find_cut_cells                       src/LevelSetCutters/CutTriangulations.jl:481-493  (ascending ids)
//
// B200 design: only the SIGN of a node value decides the classification (`isout(v) = v > 0`, LookupTables.jl:40-42),
// so node values never go to HBM.  A warp evaluates 32 consecutive nodes of an x-row and `__ballot_sync` packs their
// signs into one 32-bit word kept in shared memory; a block marches through the layers of the last direction keeping
// two node planes of sign words, and one thread then classifies 32 cells at once with bit-parallel AND/OR over the
// vertex words.  HBM traffic: L bytes per cell written (+ 8 L bytes per node read for NODAL leaves) + 1 bit per cell.
#include <algorithm>

#include "gecut_internal.cuh"

// ------------------------------------------------------------------------------------------------
// exclusive scan of uint32 (3 phases, 2048 items per block)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan across the block of one value per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  uint32_t inc = warp_incl_scan(v);
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = lane < nw ? wsum[lane] : 0;
    uint32_t si = warp_incl_scan(s);
    wsum[lane] = si - s;
    if (lane == 31) wsum[32] = si;
  }
  __syncthreads();
  uint32_t excl = wsum[wid] + inc - v;
  *total = wsum[32];
  __syncthreads();
  return excl;
}

// Single-pass exclusive scan (decoupled look-back): one launch, every element read once and written once.  Tiles are
// handed out in launch order by a ticket counter, so a tile only ever waits on tiles that are already running; tile t
// publishes its aggregate, warp 0 walks back over the 32 closest predecessors at a time until it meets an inclusive
// prefix, then publishes its own.  state word = value (62 bits) | flag (0: not ready, 1: aggregate, 2: inclusive prefix);
// value and flag travel in one 64-bit store, so no fence is needed.  `nch` arrays of n elements stored back to back
// (blockIdx.y = channel).  in == out is allowed.
#define SCAN_AGG (1ull << 62)
#define SCAN_PRE (2ull << 62)
#define SCAN_VAL ((1ull << 62) - 1ull)

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_lookback(const uint32_t* in, uint32_t* out, int64_t n, int64_t nb, unsigned long long* state, uint32_t* ticket,
                uint64_t* __restrict__ totals) {
  const int ch = blockIdx.y;
  in += (int64_t)ch * n;
  out += (int64_t)ch * n;
  volatile unsigned long long* st = state + (int64_t)ch * nb;
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_excl;
  if (threadIdx.x == 0) s_tile = atomicAdd(&ticket[ch], 1u);
  __syncthreads();
  const int64_t tile = s_tile;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile so that the order is preserved
  const int64_t base = tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
  if (base + SCAN_ITEMS <= n && ((reinterpret_cast<uintptr_t>(in + base) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i += 4) {
      const uint4 q = *reinterpret_cast<const uint4*>(in + base + i);
      v[i] = q.x;
      v[i + 1] = q.y;
      v[i + 2] = q.z;
      v[i + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = (base + i < n) ? in[base + i] : 0u;
  }
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) s += v[i];
  uint32_t tot;
  const uint32_t excl_in_tile = block_excl_scan(s, &tot);
  if (wid == 0) {
    if (lane == 0) st[tile] = (tile == 0 ? SCAN_PRE : SCAN_AGG) | (unsigned long long)tot;
    unsigned long long excl = 0ull;
    if (tile > 0) {
      int64_t j = tile - 1 - lane;
      while (true) {
        unsigned long long sv = j >= 0 ? st[j] : SCAN_PRE;  // before the first tile: inclusive prefix 0
        while (__any_sync(0xffffffffu, (sv >> 62) == 0ull))
          if ((sv >> 62) == 0ull) sv = st[j];
        const uint32_t pre = __ballot_sync(0xffffffffu, (sv >> 62) == 2ull);
        const int first = pre ? __ffs(pre) - 1 : 31;  // the closest predecessor that already knows its prefix
        unsigned long long val = lane <= first ? (sv & SCAN_VAL) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (pre) break;
        j -= 32;
      }
      if (lane == 0) st[tile] = SCAN_PRE | (excl + (unsigned long long)tot);
    }
    if (lane == 0) {
      s_excl = excl;
      if (tile == nb - 1) totals[ch] = excl + (unsigned long long)tot;
    }
  }
  __syncthreads();
  uint32_t run = (uint32_t)s_excl + excl_in_tile;
  if (base + SCAN_ITEMS <= n && ((reinterpret_cast<uintptr_t>(out + base) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i += 4) {
      uint4 q;
      q.x = run;
      q.y = q.x + v[i];
      q.z = q.y + v[i + 1];
      q.w = q.z + v[i + 2];
      run = q.w + v[i + 3];
      *reinterpret_cast<uint4*>(out + base + i) = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      if (base + i < n) out[base + i] = run;
      run += v[i];
    }
  }
}

}  // namespace

int gc_exclusive_scan_u32_multi(gecut_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int nch, uint64_t* totals_dev) {
  if (n <= 0) {
    GC_CUDA(cudaMemsetAsync(totals_dev, 0, sizeof(uint64_t) * nch, ctx->stream));
    return GECUT_OK;
  }
  const int64_t nb = gc_div_up(n, SCAN_TILE);
  unsigned long long* state = nullptr;  // [nch][nb] tile states, then nch tickets
  const size_t bytes = sizeof(unsigned long long) * (nb * nch + nch);
  GC_CHECK(gc_malloc(ctx, (void**)&state, bytes));
  GC_CUDA(cudaMemsetAsync(state, 0, bytes, ctx->stream));
  uint32_t* ticket = reinterpret_cast<uint32_t*>(state + nb * nch);
  dim3 grid((unsigned)nb, (unsigned)nch);
  GC_LAUNCH(ctx, "k_scan_lookback", k_scan_lookback<<<grid, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, nb, state, ticket, totals_dev));
  int st = gc_launch_check(ctx, "scan");
  gc_free(ctx, state);
  return st;
}

int gc_exclusive_scan_u32(gecut_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, uint64_t* total_dev) {
  return gc_exclusive_scan_u32_multi(ctx, in, out, n, 1, total_dev);
}

// ------------------------------------------------------------------------------------------------
// fused node evaluation + classification
// ------------------------------------------------------------------------------------------------
namespace {

// Warp-autonomous marching: every warp owns a tile of WXT x-words (32 cubes each) x TY cell rows and marches through
// ZS layers of the last direction.  No block barrier in the loop (only __syncwarp); two node planes of sign words per
// level set live in a warp-private slice of shared memory.

// ---- box culling -----------------------------------------------------------------------------------------------------
// Conservative sign of a closed-form level set over a whole box of nodes (x in [xlo, xhi], y in [ylo, yhi], a range of the
// marching coordinate).  SPHERE / CYLINDER:  phi = |w|^2 - (w.v)^2 - R^2 = sum_d w_d^2 (1 - v_d^2) - 2 sum_{d<e} w_d w_e
// v_d v_e - R^2 (an identity for ANY v; v = 0 for the sphere), PLANE: phi = w.v; every term is bounded by interval
// arithmetic on the per-direction ranges.  The box is declared one-signed only when the bound clears zero by 1e-9 of the
// magnitude of the terms -- six orders of magnitude above the rounding error of the reference formula (a few ulp of the
// same magnitude), so the exact evaluation it replaces would have given the same sign at every node: results stay bit
// identical, nodes near a surface are always evaluated exactly.
constexpr int CLS_CBN = 10;
constexpr int CLS_CZ = 8;  // layers per culling chunk

__device__ __forceinline__ void cls_range_mul(double k, double lo, double hi, double* rlo, double* rhi) {
  const double a = k * lo, b = k * hi;
  *rlo = fmin(a, b);
  *rhi = fmax(a, b);
}

template <int D>
__device__ __forceinline__ void cls_tile_bounds(const gecut_leaf& lf, double xlo, double xhi, double ylo, double yhi,
                                                double* cb) {
  const double wxl = xlo - lf.x0[0], wxh = xhi - lf.x0[0];
  const double wyl = (D == 3) ? ylo - lf.x0[1] : 0.0, wyh = (D == 3) ? yhi - lf.x0[1] : 0.0;
  if (lf.kind == GECUT_LEAF_PLANE) {
    cls_range_mul(lf.v[0], wxl, wxh, cb + 0, cb + 1);
    if (D == 3) cls_range_mul(lf.v[1], wyl, wyh, cb + 2, cb + 3);
    else cb[2] = cb[3] = 0.0;
    return;
  }
  const bool sph = lf.kind == GECUT_LEAF_SPHERE;
  const double v0 = sph ? 0.0 : lf.v[0], v1 = (sph || D == 2) ? 0.0 : lf.v[1];
  const double sxl = (wxl <= 0.0 && wxh >= 0.0) ? 0.0 : fmin(wxl * wxl, wxh * wxh), sxh = fmax(wxl * wxl, wxh * wxh);
  const double syl = (wyl <= 0.0 && wyh >= 0.0) ? 0.0 : fmin(wyl * wyl, wyh * wyh), syh = fmax(wyl * wyl, wyh * wyh);
  cls_range_mul(1.0 - v0 * v0, sxl, sxh, cb + 0, cb + 1);
  cls_range_mul(1.0 - v1 * v1, syl, syh, cb + 2, cb + 3);
  if (D == 2) cb[2] = cb[3] = 0.0;
  // cross term -2 v0 v1 wx wy: interval product of the two ranges
  const double k = -2.0 * v0 * v1;
  const double p1 = k * wxl * wyl, p2 = k * wxl * wyh, p3 = k * wxh * wyl, p4 = k * wxh * wyh;
  cb[4] = fmin(fmin(p1, p2), fmin(p3, p4));
  cb[5] = fmax(fmax(p1, p2), fmax(p3, p4));
  cb[6] = wxl;
  cb[7] = wxh;
  cb[8] = wyl;
  cb[9] = wyh;
}

// 0: mixed or too close to call, 1: every node of the box has phi <= 0 (IN), 2: every node has phi > 0 (OUT); the box is
// the tile's node patch x the node planes with marching coordinate in [zlo, zhi]
template <int D>
__device__ __forceinline__ uint32_t cls_box_flag(const gecut_leaf& lf, const double* cb, double zlo, double zhi) {
  const double wzl = zlo - lf.x0[D - 1], wzh = zhi - lf.x0[D - 1];
  double lo, hi, mag;
  if (lf.kind == GECUT_LEAF_PLANE) {
    double tl, th;
    cls_range_mul(lf.v[D - 1], wzl, wzh, &tl, &th);
    lo = cb[0] + cb[2] + tl;
    hi = cb[1] + cb[3] + th;
    mag = fmax(fabs(cb[0]), fabs(cb[1])) + fmax(fabs(cb[2]), fabs(cb[3])) + fmax(fabs(tl), fabs(th));
  } else {
    const bool sph = lf.kind == GECUT_LEAF_SPHERE;
    const double v0 = sph ? 0.0 : lf.v[0], v1 = (sph || D == 2) ? 0.0 : lf.v[1], vz = sph ? 0.0 : lf.v[D - 1];
    const double szl = (wzl <= 0.0 && wzh >= 0.0) ? 0.0 : fmin(wzl * wzl, wzh * wzh), szh = fmax(wzl * wzl, wzh * wzh);
    double ql, qh;
    cls_range_mul(1.0 - vz * vz, szl, szh, &ql, &qh);
    lo = cb[0] + cb[2] + ql + cb[4];
    hi = cb[1] + cb[3] + qh + cb[5];
    mag = fabs(cb[1]) + fabs(cb[3]) + fabs(qh) + fmax(fabs(cb[4]), fabs(cb[5]));
    if (!sph) {  // cross terms with the marching direction: -2 v_d v_z w_d w_z, interval products
      const double kx = -2.0 * v0 * vz, ky = -2.0 * v1 * vz;
      const double a1 = kx * cb[6] * wzl, a2 = kx * cb[6] * wzh, a3 = kx * cb[7] * wzl, a4 = kx * cb[7] * wzh;
      const double b1 = ky * cb[8] * wzl, b2 = ky * cb[8] * wzh, b3 = ky * cb[9] * wzl, b4 = ky * cb[9] * wzh;
      const double xl = fmin(fmin(a1, a2), fmin(a3, a4)), xh = fmax(fmax(a1, a2), fmax(a3, a4));
      const double yl = fmin(fmin(b1, b2), fmin(b3, b4)), yh = fmax(fmax(b1, b2), fmax(b3, b4));
      lo += xl + yl;
      hi += xh + yh;
      mag += fmax(fabs(xl), fabs(xh)) + fmax(fabs(yl), fabs(yh));
    }
    const double c = lf.R * lf.R;
    lo -= c;
    hi -= c;
    mag += c;
  }
  const double tol = 1e-9 * mag;
  return lo > tol ? 2u : (hi < -tol ? 1u : 0u);
}

#ifndef CLS_WXT_TET
#define CLS_WXT_TET 2
#endif
#ifndef CLS_WXT_HEX
#define CLS_WXT_HEX 2
#endif
template <int D, bool SIMPLEX>
struct ClsCfg {
  static constexpr int NVC = 1 << D;                             // vertices of the n-cube
  static constexpr int CPC = SIMPLEX ? (D == 3 ? 6 : 2) : 1;     // bg cells per n-cube
  static constexpr int LPW = (D == 3) ? (SIMPLEX ? 2 : 1) : 4;   // lanes sharing one 32-cube word
  static constexpr int CPL = 32 / LPW;                           // cubes per lane and layer
  static constexpr int NTASK = 32 / LPW;                         // words per warp and layer
  static constexpr int WXT = (D == 3) ? (SIMPLEX ? CLS_WXT_TET : CLS_WXT_HEX) : 8;   // x words of the warp tile
  static constexpr int TY = NTASK / WXT;                         // cell rows of the warp tile
  static constexpr int RY = (D == 3) ? TY + 1 : 1;               // node rows per plane
  static constexpr int PW = RY * (WXT + 1);                      // sign words per plane and level set (+ x halo bit)
  static constexpr int OUTB = CPL * CPC;                         // state bytes per lane and layer: 32 / 96 / 8 / 16
  static constexpr int NQ = OUTB / 8;                            // ... in 64-bit words
  static constexpr int WARPS = 4;
};

#ifndef CLS_MIN_BLOCKS
#define CLS_MIN_BLOCKS 4
#endif
// words per x-row of the cut mask
__host__ __device__ inline int gc_words_per_row(int nx) { return (nx + 31) / 32; }

// NQ 64-bit words to 8*NQ contiguous bytes with the widest store the size allows (32-byte stores: whole sectors)
template <int NQ>
__device__ __forceinline__ void cls_store(int8_t* out, const unsigned long long* e) {
  if constexpr (NQ % 4 == 0) {
#pragma unroll
    for (int q = 0; q < NQ; q += 4)
      asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(out + 8 * q), "l"(e[q]), "l"(e[q + 1]), "l"(e[q + 2]),
                   "l"(e[q + 3])
                   : "memory");
  } else if constexpr (NQ == 2) {
    *reinterpret_cast<ulonglong2*>(out) = make_ulonglong2(e[0], e[1]);
  } else {
#pragma unroll
    for (int q = 0; q < NQ; ++q) reinterpret_cast<unsigned long long*>(out)[q] = e[q];
  }
}

// scalar fall-back when rows are not 32-byte aligned (nx not a multiple of 32): the first nbytes bytes of e[]
template <int NQ>
__device__ __forceinline__ void cls_store_bytes(int8_t* out, const unsigned long long* e, int nbytes) {
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (q * 8 + k < nbytes) out[q * 8 + k] = (int8_t)((e[q] >> (8 * k)) & 0xFFull);
}

// the first nbytes bytes of e[] to `out` with the widest stores the address allows: whole 32-byte sectors when the row
// segment is aligned (grids whose x size is a multiple of 32 always are), else 8- or 4-byte words and a byte tail
template <int NQ>
__device__ __forceinline__ void cls_store_any(int8_t* out, const unsigned long long* e, int nbytes) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(out);
  constexpr uintptr_t VMASK = (NQ % 4 == 0) ? 31 : (NQ == 2 ? 15 : 7);
  if (nbytes == 8 * NQ && (a & VMASK) == 0) {
    cls_store<NQ>(out, e);
  } else if ((a & 7) == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (8 * (q + 1) <= nbytes) {
        reinterpret_cast<unsigned long long*>(out)[q] = e[q];
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (q * 8 + k < nbytes) out[q * 8 + k] = (int8_t)((e[q] >> (8 * k)) & 0xFFull);
      }
    }
  } else if ((a & 3) == 0) {
#pragma unroll
    for (int w = 0; w < 2 * NQ; ++w) {
      const uint32_t v = (uint32_t)(e[w >> 1] >> (32 * (w & 1)));
      if (4 * (w + 1) <= nbytes) {
        reinterpret_cast<uint32_t*>(out)[w] = v;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (w * 4 + k < nbytes) out[w * 4 + k] = (int8_t)((v >> (8 * k)) & 0xFFu);
      }
    }
  } else {
    cls_store_bytes<NQ>(out, e, nbytes);
  }
}

// LH > 0: "hoisted" evaluation for models with at most LH level sets.  SPHERE / PLANE / CYLINDER split into parts that
// depend on one coordinate only: the x parts live in registers (per x word of the tile), the y parts in shared memory
// (per node row), the marching coordinate enters once per plane.  The association of the reference formulas
// (gc_eval_leaf) is preserved exactly -- same operations, same rounding -- only loop-invariant ones move.
template <int D, bool SIMPLEX, int LH, bool SG /* some level set of the launch comes as precomputed sign words */>
__global__ void __launch_bounds__(ClsCfg<D, SIMPLEX>::WARPS * 32, CLS_MIN_BLOCKS)
k_classify(const GcGrid g, const GcLeaves lv, int8_t* __restrict__ states, int64_t pitch,
           uint32_t* __restrict__ cutmask, uint32_t* __restrict__ cutcount,
           uint32_t* __restrict__ cut22count /* TET cells split 2-2 per group (one level set), or nullptr */,
           uint32_t* __restrict__ cut22mask /* ... and which ones, laid out like cutmask; written for crossed groups only */,
           uint32_t* __restrict__ rowcount /* [2][rows]: cut cells (2-2 TETs) per x-row of cubes, summed with atomics */,
           int64_t nrows,
           const uint8_t* __restrict__ simplexify_raw,
           unsigned long long* __restrict__ bg_counts /* [L][2][6]: IN / OUT cells per level set and cell type */,
           int ZS, int ntx, int nty, int64_t ntiles, int accumulate /* OR into the cut mask of earlier level-set chunks */,
           const uint32_t* __restrict__ worklist /* tiles k_classify_cull left to this kernel (+ their number), or nullptr */,
           const uint32_t* __restrict__ worklist_n) {
  using C = ClsCfg<D, SIMPLEX>;
  constexpr int NVC = C::NVC, CPC = C::CPC, LPW = C::LPW, CPL = C::CPL, WXT = C::WXT, TY = C::TY, RY = C::RY;
  constexpr int PW = C::PW, NQ = C::NQ, WARPS = C::WARPS;
  constexpr int LHS = LH > 0 ? LH : 1;
  constexpr int ROWD = 1 + 2 * (LH > 0 ? LH : 0);  // doubles per node row: y, then (wy^2, wy*vy) per hoisted level set
  constexpr uint32_t FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char cls_smem[];
  if (worklist != nullptr) {  // the grid is sized for every tile; blocks past the end of the list leave at once
    ntiles = (int64_t)*worklist_n;
    if ((int64_t)blockIdx.x * C::WARPS >= ntiles) return;
  }
  __shared__ unsigned long long s_lut[SIMPLEX ? (1 << NVC) : 1];  // sign byte of an n-cube -> states of its CPC simplices
  // per warp and level set: IN[6], OUT[6] per cell type, + IN/OUT of uniform groups (warp-private: no block barrier)
  __shared__ unsigned int s_cnt_all[WARPS][GC_LMAX * 14];
  const int L = lv.L;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned int* s_cnt = s_cnt_all[wid];
  for (int i = lane; i < L * 14; i += 32) s_cnt[i] = 0u;
  if (SIMPLEX) {
    for (int sb = threadIdx.x; sb < (1 << NVC); sb += WARPS * 32) {
      unsigned long long e = 0ull;
#pragma unroll
      for (int t = 0; t < CPC; ++t) {
        bool any = false, all = true;
#pragma unroll
        for (int m = 0; m <= D; ++m) {
          const bool o = (sb >> gc_simplexify_raw(D - 2, t, m)) & 1;
          any |= o;
          all &= o;
        }
        const unsigned long long st = all ? 0x01ull : (any ? 0x00ull : 0xFFull);  // OUT / CUT / IN
        e |= st << (8 * t);
      }
      s_lut[sb] = e;
    }
  }
  __syncthreads();

  const int64_t slot = (int64_t)blockIdx.x * WARPS + wid;
  if (slot < ntiles) {
    const int64_t tile = worklist != nullptr ? (int64_t)worklist[slot] : slot;
    const int nx = g.n[0];
    const int ny = (D == 3) ? g.n[1] : 1;
    const int wxtot = gc_words_per_row(nx);
    const int tx = (int)(tile % ntx);
    const int ty = (int)((tile / ntx) % nty);
    const int tz = (int)(tile / ((int64_t)ntx * nty));
    const int w0 = tx * WXT;
    const int j0 = ty * TY;
    const int kz0 = g.k0 + tz * ZS;
    const int kz1 = min(kz0 + ZS, g.k1);
    const bool vec_ok = (nx & 31) == 0;
    // warp-private shared memory: row doubles, then the sign planes [2][L][PW]
    const size_t per_warp = sizeof(double) * RY * ROWD + sizeof(uint32_t) * 2 * L * PW;
    double* rowd = reinterpret_cast<double*>(cls_smem + (size_t)wid * ((per_warp + 15) & ~(size_t)15));
    uint32_t* planes = reinterpret_cast<uint32_t*>(rowd + RY * ROWD);

    // ---- evaluation set-up: this lane's node column of every x word, the halo node column (lane = node row)
    double xw[WXT];
    uint32_t xmask = 0u;  // bit ww: this lane's node of word ww exists
#pragma unroll
    for (int ww = 0; ww < WXT; ++ww) {
      const int i = 32 * (w0 + ww) + lane;
      xw[ww] = gc_coord(g, 0, i);
      if (i < g.nn[0]) xmask |= 1u << ww;
    }
    const int ih = 32 * (w0 + WXT);
    const bool hvalid = (ih < g.nn[0]) && (lane < RY) && (D == 2 || j0 + lane < g.nn[1]);
    const double xh = gc_coord(g, 0, ih);
    const double yh = (D == 3) ? gc_coord(g, 1, j0 + lane) : 0.0;
    uint32_t rowmask = 0u;  // bit r: node row r exists
#pragma unroll
    for (int r = 0; r < RY; ++r)
      if (D == 2 || j0 + r < g.nn[1]) rowmask |= 1u << r;
    if (D == 3 && lane < RY) {
      rowd[lane * ROWD] = yh;
      if (LH > 0) {
#pragma unroll
        for (int ls = 0; ls < LH; ++ls) {
          const double wy = ls < L ? __dsub_rn(yh, lv.leaf[ls].x0[1]) : 0.0;
          rowd[lane * ROWD + 1 + 2 * ls] = __dmul_rn(wy, wy);
          rowd[lane * ROWD + 2 + 2 * ls] = ls < L ? __dmul_rn(wy, lv.leaf[ls].v[1]) : 0.0;
        }
      }
    }
    double hxa[LHS][WXT], hxb[LHS][WXT], ha[LHS], hb[LHS], hc[LHS];
    if (LH > 0) {
#pragma unroll
      for (int ls = 0; ls < LH; ++ls) {
        const bool on = ls < L;
        const double x00 = on ? lv.leaf[ls].x0[0] : 0.0, x01 = on ? lv.leaf[ls].x0[1] : 0.0;
        const double v0 = on ? lv.leaf[ls].v[0] : 0.0, v1 = on ? lv.leaf[ls].v[1] : 0.0;
        hc[ls] = on ? __dmul_rn(lv.leaf[ls].R, lv.leaf[ls].R) : 0.0;
#pragma unroll
        for (int ww = 0; ww < WXT; ++ww) {
          const double wx = __dsub_rn(xw[ww], x00);
          hxa[ls][ww] = __dmul_rn(wx, wx);
          hxb[ls][ww] = __dmul_rn(wx, v0);
        }
        const double wx = __dsub_rn(xh, x00);
        const double wy = __dsub_rn(yh, x01);
        // in-plane partial sums of the halo column, associated like gc_dot3: (a0*b0 + a1*b1) [+ a2*b2 per plane]
        ha[ls] = (D == 3) ? __dadd_rn(__dmul_rn(wx, wx), __dmul_rn(wy, wy)) : __dmul_rn(wx, wx);
        hb[ls] = (D == 3) ? __dadd_rn(__dmul_rn(wx, v0), __dmul_rn(wy, v1)) : __dmul_rn(wx, v0);
      }
    }
    __syncwarp();

    // ---- classification set-up: lane -> (word task, part of the word)
    const int task = lane / LPW, sub = lane % LPW;
    const int ww_c = task % WXT, r_c = task / WXT;
    const int c0 = sub * CPL;
    const int i0 = 32 * (w0 + ww_c);
    const int jc = j0 + r_c;
    const int ncw = max(0, min(32, nx - i0));                              // live cubes of the word
    const uint32_t livew = ncw == 32 ? FULL : ((1u << ncw) - 1u);
    const int nl = max(0, min(CPL, ncw - c0));                             // live cubes of this lane
    const uint32_t livel = nl == 32 ? FULL : ((1u << nl) - 1u);
    const bool lane_on = nl > 0 && jc < ny;
    const bool word_on = sub == 0 && ncw > 0 && jc < ny;
    const int pofs = r_c * (WXT + 1) + ww_c;
    const int64_t layer_rows = (D == 3) ? ny : 1;
    unsigned int cnt_in[CPC], cnt_out[CPC], uni_in = 0u, uni_out = 0u;
#pragma unroll
    for (int t = 0; t < CPC; ++t) cnt_in[t] = cnt_out[t] = 0u;
    bool slow_seen = false;

    auto flush_counts = [&](int ls) {
      const unsigned int a = __reduce_add_sync(FULL, uni_in), b = __reduce_add_sync(FULL, uni_out);
      if (lane == 0) {
        if (a) atomicAdd(&s_cnt[ls * 14 + 12], a);
        if (b) atomicAdd(&s_cnt[ls * 14 + 13], b);
      }
      uni_in = uni_out = 0u;
      if (slow_seen) {
#pragma unroll
        for (int t = 0; t < CPC; ++t) {
          const unsigned int ai = __reduce_add_sync(FULL, cnt_in[t]), ao = __reduce_add_sync(FULL, cnt_out[t]);
          if (lane == 0) {
            if (ai) atomicAdd(&s_cnt[ls * 14 + t], ai);
            if (ao) atomicAdd(&s_cnt[ls * 14 + 6 + t], ao);
          }
          cnt_in[t] = cnt_out[t] = 0u;
        }
        slow_seen = false;
      }
    };

    for (int kp = kz0; kp <= kz1; ++kp) {  // node planes kz0..kz1
      uint32_t* Pk = planes + ((kp - kz0) & 1) * L * PW;
      const double zc = gc_coord(g, D - 1, kp);
      // ================= signs of node plane kp: one ballot per (node row, x word) and level set
      for (int ls = 0; ls < L; ++ls) {
        uint32_t* P = Pk + ls * PW;
        const int kind = lv.leaf[ls].kind;
        if (SG && lv.signs[ls]) {
          // NODAL leaf: the sign words were written by k_nodal_signs at full streaming bandwidth; a lane fetches one
          // word of the plane (node row r, x word ww; the halo column only needs bit 0 of the next word)
          const int wpr = lv.signs_wpr;
          const uint32_t* S = lv.signs[ls] + (int64_t)(kp - g.k0) * ((D == 3) ? g.nn[1] : 1) * wpr;
          for (int idx = lane; idx < PW; idx += 32) {
            const int r = idx / (WXT + 1), ww = idx - r * (WXT + 1);
            const int j = (D == 3) ? j0 + r : 0, wx = w0 + ww;
            uint32_t word = 0u;
            if ((D == 2 || j < g.nn[1]) && wx < wpr) word = __ldg(S + (int64_t)j * wpr + wx);
            P[idx] = (ww == WXT) ? (word & 1u) : word;
          }
          continue;
        }
        // rows x words with a value functor f(ww2, wv, r, ww) (in-plane partial sums) -- fully unrolled
        auto sweep = [&](auto f) {
#pragma unroll
          for (int r = 0; r < RY; ++r) {
            double wy2 = 0.0, wyv = 0.0;
            if (D == 3 && LH > 0) {
              const int lsh = ls < LHS ? ls : 0;
              wy2 = rowd[r * ROWD + 1 + 2 * lsh];
              wyv = rowd[r * ROWD + 2 + 2 * lsh];
            }
#pragma unroll
            for (int ww = 0; ww < WXT; ++ww) {
              const int lsh = (LH > 0 && ls < LHS) ? ls : 0;
              double a = 0.0, b = 0.0;
              if (LH > 0) {
                // select the registers of level set ls without dynamic indexing
                a = hxa[0][ww];
                b = hxb[0][ww];
#pragma unroll
                for (int q = 1; q < LHS; ++q)
                  if (lsh == q) {
                    a = hxa[q][ww];
                    b = hxb[q][ww];
                  }
              }
              const double ww2 = (D == 3) ? __dadd_rn(a, wy2) : a;
              const double wv = (D == 3) ? __dadd_rn(b, wyv) : b;
              const double v = f(ww2, wv, r, ww, false);
              const bool ok = ((rowmask >> r) & 1u) && ((xmask >> ww) & 1u);
              const uint32_t word = __ballot_sync(FULL, ok && (v > 0.0));
              if (lane == 0) P[r * (WXT + 1) + ww] = word;
            }
          }
          {  // x halo: the node column right of the tile, lane = node row
            double a = 0.0, b = 0.0;
            if (LH > 0) {
              const int lsh = ls < LHS ? ls : 0;
              a = ha[0];
              b = hb[0];
#pragma unroll
              for (int q = 1; q < LHS; ++q)
                if (lsh == q) {
                  a = ha[q];
                  b = hb[q];
                }
            }
            const double v = f(a, b, lane, WXT, true);
            const uint32_t word = __ballot_sync(FULL, hvalid && (v > 0.0));
            if (lane < RY) P[lane * (WXT + 1) + WXT] = (word >> lane) & 1u;
          }
        };
        // generic value of one node: closed-form leaf or NODAL vector
        auto node_value = [&](int r, int ww, bool halo) -> double {
          const int i = halo ? ih : 32 * (w0 + ww) + lane;
          const int j = (D == 3) ? j0 + r : 0;
          if (kind == GECUT_LEAF_NODAL) {
            const bool ok = halo ? hvalid : (((rowmask >> (r & 31)) & 1u) && (i < g.nn[0]));
            const int64_t node = (D == 3) ? ((int64_t)i + (int64_t)g.nn[0] * ((int64_t)j + (int64_t)g.nn[1] * kp))
                                          : ((int64_t)i + (int64_t)g.nn[0] * kp);
            return ok ? __ldg(lv.nodal[ls] + (node - lv.nodal_base)) : -1.0;
          }
          double x[3];
          x[0] = halo ? xh : gc_coord(g, 0, i);
          if (D == 3) {
            x[1] = halo ? yh : rowd[(r < RY ? r : 0) * ROWD];
            x[2] = zc;
          } else {
            x[1] = zc;
            x[2] = 0.0;
          }
          return gc_eval_leaf(lv.leaf[ls], x, D);
        };
        if (LH > 0 && ls < LHS && (kind == GECUT_LEAF_SPHERE || kind == GECUT_LEAF_PLANE || kind == GECUT_LEAF_CYLINDER)) {
          const double wz = __dsub_rn(zc, lv.leaf[ls].x0[D - 1]);
          const double wz2 = __dmul_rn(wz, wz);
          const double wzv = __dmul_rn(wz, lv.leaf[ls].v[D - 1]);
          double c = hc[0];
#pragma unroll
          for (int q = 1; q < LHS; ++q)
            if (ls == q) c = hc[q];
          if (kind == GECUT_LEAF_SPHERE) {
            sweep([&](double ww2, double, int, int, bool) { return __dsub_rn(__dadd_rn(ww2, wz2), c); });
          } else if (kind == GECUT_LEAF_PLANE) {
            sweep([&](double, double wv, int, int, bool) { return __dadd_rn(wv, wzv); });
          } else {
            sweep([&](double ww2, double wv, int, int, bool) {
              const double A = __dadd_rn(wv, wzv);
              const double H2 = __dadd_rn(ww2, wz2);
              return __dsub_rn(__dsub_rn(H2, __dmul_rn(A, A)), c);
            });
          }
        } else {
          sweep([&](double, double, int r, int ww, bool halo) { return node_value(r, ww, halo); });
        }
      }
      __syncwarp();
      if (kp > kz0) {
        // ================= classify layer kp-1 from planes kp-1 [lo] and kp [hi]
        const int kc = kp - 1;  // global layer index
        const uint32_t* Q0 = planes + ((kp - 1 - kz0) & 1) * L * PW;
        const uint32_t* Q1 = Pk;
        const int64_t rowid = (int64_t)(kc - g.k0) * layer_rows + jc;
        const int64_t cube0 = rowid * nx + i0 + c0;
        const int64_t gidx = rowid * wxtot + (w0 + ww_c);
        uint32_t cutacc[CPC];
        uint32_t n22 = 0u;
        bool slow_word = false;  // some level set of this launch crosses the lane's 32-cube word
#pragma unroll
        for (int t = 0; t < CPC; ++t) cutacc[t] = 0u;
        uint32_t n_old = 0u;  // the group's count after the earlier level-set chunks
        if (accumulate && word_on) {
          n_old = cutcount[gidx];
          if (n_old != 0u) {  // mask words exist only where the group count is not zero
#pragma unroll
            for (int t = 0; t < CPC; ++t) cutacc[t] = cutmask[gidx * CPC + t];
          }
        }
        for (int ls = 0; ls < L; ++ls) {
          uint32_t V[NVC];
#pragma unroll
          for (int v = 0; v < NVC; ++v) {
            const int xo = v & 1;
            const int ro = (D == 3) ? ((v >> 1) & 1) : 0;
            const int po = (D == 3) ? ((v >> 2) & 1) : ((v >> 1) & 1);
            const uint32_t* row = (po ? Q1 : Q0) + ls * PW + pofs + ro * (WXT + 1);
            V[v] = xo ? __funnelshift_r(row[0], row[1], 1) : row[0];
          }
          uint32_t vor = 0u, vand = FULL;
#pragma unroll
          for (int v = 0; v < NVC; ++v) {
            vor |= V[v];
            vand &= V[v];
          }
          const uint32_t lor = (vor >> c0) & livel, land = (vand >> c0) & livel;
          const bool u_in = lor == 0u, u_out = land == livel;
          const bool w_uni = ((vor & livew) == 0u) || ((vand & livew) == livew);
          const bool fast = (!lane_on || u_in || u_out) && (!word_on || w_uni);
          int8_t* out = states + (int64_t)ls * pitch + cube0 * CPC;
          unsigned long long e[NQ];
          if (__all_sync(FULL, fast)) {
            // uniform everywhere in the warp tile (the common case away from the surfaces): constant states
            if (lane_on) {
              const unsigned long long pat = u_in ? 0xFFFFFFFFFFFFFFFFull : 0x0101010101010101ull;
#pragma unroll
              for (int q = 0; q < NQ; ++q) e[q] = pat;
              if (vec_ok) cls_store<NQ>(out, e);
              else cls_store_any<NQ>(out, e, nl * CPC);
              if (u_in) uni_in += (unsigned int)nl; else uni_out += (unsigned int)nl;
            }
          } else {
            slow_seen = true;
            slow_word = slow_word || !w_uni;
            if (SIMPLEX) {
              // per cell type: any / all over the simplex' vertices -> cut mask word and tallies
#pragma unroll
              for (int t = 0; t < CPC; ++t) {
                uint32_t any = 0u, all = FULL;
#pragma unroll
                for (int m = 0; m <= D; ++m) {
                  const uint32_t a = V[gc_simplexify_raw(D - 2, t, m)];  // compile-time index
                  any |= a;
                  all &= a;
                }
                cutacc[t] |= any & ~all & livew;
                if (D == 3 && cut22count != nullptr) {
                  // a cut TET has 1, 2 or 3 vertices OUT: even parity = the 2-2 split (6 subcells, 8 points, 2 facets)
                  const uint32_t par = V[gc_simplexify_raw(D - 2, t, 0)] ^ V[gc_simplexify_raw(D - 2, t, 1)] ^
                                       V[gc_simplexify_raw(D - 2, t, 2)] ^ V[gc_simplexify_raw(D - 2, t, D)];
                  const uint32_t s22 = any & ~all & livew & ~par;
                  n22 += __popc(s22);
                  if (word_on) cut22mask[gidx * CPC + t] = s22;  // read by k_cutmask_emit where the group count is not 0
                }
                if (lane_on) {
                  cnt_in[t] += __popc((~any >> c0) & livel);
                  cnt_out[t] += __popc((all >> c0) & livel);
                }
              }
              // state bytes: sign byte of every cube (bit v = vertex v OUT) -> LUT row
#pragma unroll
              for (int q4 = 0; q4 < CPL / 4; ++q4) {
                uint32_t acc = 0u;
#pragma unroll
                for (int v = 0; v < NVC; ++v) {
                  const uint32_t nib = (V[v] >> (c0 + 4 * q4)) & 0xFu;
                  acc += ((nib * 0x00204081u) & 0x01010101u) << v;
                }
                const unsigned long long e0 = s_lut[acc & 0xFFu], e1 = s_lut[(acc >> 8) & 0xFFu];
                const unsigned long long e2 = s_lut[(acc >> 16) & 0xFFu], e3 = s_lut[acc >> 24];
                if (CPC == 6) {
                  e[3 * q4 + 0] = e0 | (e1 << 48);
                  e[3 * q4 + 1] = (e1 >> 16) | (e2 << 32);
                  e[3 * q4 + 2] = (e2 >> 32) | (e3 << 16);
                } else {
                  e[q4] = e0 | (e1 << 16) | (e2 << 32) | (e3 << 48);
                }
              }
            } else {
              cutacc[0] |= vor & ~vand & livew;
              if (lane_on) {
                cnt_in[0] += __popc(~lor & livel);
                cnt_out[0] += __popc(land);
              }
              const uint32_t o = vand >> c0, n = ~vor >> c0;
#pragma unroll
              for (int q = 0; q < NQ; ++q) {
                unsigned long long w8 = 0ull;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint32_t o4 = (o >> (8 * q + 4 * h)) & 0xFu, n4 = (n >> (8 * q + 4 * h)) & 0xFu;
                  const uint32_t so = (o4 * 0x00204081u) & 0x01010101u;
                  const uint32_t si = ((n4 * 0x00204081u) & 0x01010101u) * 0xFFu;
                  w8 |= (unsigned long long)(so | si) << (32 * h);  // OUT(1) where all, IN(0xFF) where !any, CUT(0) otherwise
                }
                e[q] = w8;
              }
            }
            if (lane_on) {
              if (vec_ok) cls_store<NQ>(out, e);
              else cls_store_any<NQ>(out, e, nl * CPC);
            }
          }
          if (L > 1) flush_counts(ls);
        }
        if (word_on && slow_word) {
          // mask, count (and 2-2 count) were zeroed before the first launch: only the groups a surface crosses -- a few per
          // cent -- store anything (scattered 4-byte stores of zeros cost partial-sector read-modify-writes in DRAM)
          uint32_t* cm = cutmask + gidx * CPC;
          uint32_t n = 0u;
#pragma unroll
          for (int t = 0; t < CPC; ++t) {
            cm[t] = cutacc[t];
            n += __popc(cutacc[t]);
          }
          cutcount[gidx] = n;
          // per-row totals: the compaction scans the rows (16x fewer elements than groups at 512 cubes per row)
          if (n != n_old) atomicAdd(&rowcount[rowid], n - n_old);
          if (SIMPLEX && D == 3 && cut22count != nullptr) {
            cut22count[gidx] = n22;
            if (n22) atomicAdd(&rowcount[nrows + rowid], n22);
          }
        }
      }
      __syncwarp();  // the next plane overwrites the buffer this step has just read
    }
    if (L == 1) flush_counts(0);
  }
  __syncwarp();
  for (int i = lane; i < L * 12; i += 32) {
    const unsigned int v = s_cnt[(i / 12) * 14 + (i % 12)] + s_cnt[(i / 12) * 14 + 12 + (i % 12) / 6];
    // uniform groups count for every cell type of the n-cube
    const bool type_exists = (i % 6) < CPC;
    if (type_exists && v) atomicAdd(&bg_counts[i], (unsigned long long)v);
  }
}


// Kernel 1 of the closed-form classification: one warp per tile of WXT x-words x TY rows x CLS_CZ layers.  When every level
// set of the launch has one sign on all nodes of the tile's box (interval bounds, ~3/4 of the tiles of a 512^3 sphere) the
// tile's states are constants and it holds no cut cell: no node is evaluated, the bytes leave at store speed (few
// registers, full occupancy).  The other tiles go on a work list for the marching kernel, whose latency-bound work then
// comes in equal pieces and no longer shares its few resident warps with the streaming part.
template <int D, bool SIMPLEX, int LH>
__global__ void __launch_bounds__(256)
k_classify_cull(const GcGrid g, const GcLeaves lv, int8_t* __restrict__ states, int64_t pitch,
                unsigned long long* __restrict__ bg_counts, int ntx, int nty, int64_t ntiles,
                uint32_t* __restrict__ worklist, uint32_t* __restrict__ worklist_n) {
  using C = ClsCfg<D, SIMPLEX>;
  constexpr int CPC = C::CPC, LPW = C::LPW, CPL = C::CPL, WXT = C::WXT, TY = C::TY, NQ = C::NQ;
  constexpr uint32_t FULL = 0xffffffffu;
  __shared__ unsigned int s_uni[LH][2];
  if (threadIdx.x < LH * 2) s_uni[threadIdx.x >> 1][threadIdx.x & 1] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t tile = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (tile < ntiles) {
    const int nx = g.n[0];
    const int ny = (D == 3) ? g.n[1] : 1;
    const int tx = (int)(tile % ntx);
    const int ty = (int)((tile / ntx) % nty);
    const int tz = (int)(tile / ((int64_t)ntx * nty));
    const int w0 = tx * WXT, j0 = ty * TY;
    const int kz0 = g.k0 + tz * CLS_CZ, kz1 = min(kz0 + CLS_CZ, g.k1);
    // the box: node columns 32*w0 .. 32*(w0+WXT) (incl. the halo column), node rows j0 .. j0+TY, node planes kz0 .. kz1
    const double xlo = gc_coord(g, 0, 32 * w0), xhi = gc_coord(g, 0, min(32 * (w0 + WXT), g.nn[0] - 1));
    const double ylo = (D == 3) ? gc_coord(g, 1, j0) : 0.0;
    const double yhi = (D == 3) ? gc_coord(g, 1, min(j0 + TY, g.nn[1] - 1)) : 0.0;
    const double zlo = gc_coord(g, D - 1, kz0), zhi = gc_coord(g, D - 1, kz1);
    uint32_t uflags = 0u;
    bool all_uni = true;
#pragma unroll
    for (int ls = 0; ls < LH; ++ls) {
      double cb[CLS_CBN];
      cls_tile_bounds<D>(lv.leaf[ls], xlo, xhi, ylo, yhi, cb);
      const uint32_t uf = cls_box_flag<D>(lv.leaf[ls], cb, zlo, zhi);
      uflags |= uf << (2 * ls);
      all_uni = all_uni && uf != 0u;
    }
    if (!all_uni) {
      if (lane == 0) worklist[atomicAdd(worklist_n, 1u)] = (uint32_t)tile;
    } else {
      const int task = lane / LPW, sub = lane % LPW;
      const int ww_c = task % WXT, r_c = task / WXT;
      const int c0 = sub * CPL;
      const int i0 = 32 * (w0 + ww_c);
      const int jc = j0 + r_c;
      const int nl = max(0, min(CPL, min(32, nx - i0) - c0));  // live cubes of this lane
      const bool lane_on = nl > 0 && jc < ny;
      const bool vec_ok = (nx & 31) == 0;
      const int64_t layer_rows = (D == 3) ? ny : 1;
      const int nlay = kz1 - kz0;
      const int64_t rowid0 = (int64_t)(kz0 - g.k0) * layer_rows + jc;
      const int64_t step = layer_rows * nx * CPC;
#pragma unroll
      for (int ls = 0; ls < LH; ++ls) {
        const uint32_t uf = (uflags >> (2 * ls)) & 3u;
        unsigned int mine = 0u;
        if (lane_on) {
          unsigned long long e[NQ];
          const unsigned long long pat = uf == 1u ? 0xFFFFFFFFFFFFFFFFull : 0x0101010101010101ull;
#pragma unroll
          for (int q = 0; q < NQ; ++q) e[q] = pat;
          int8_t* out = states + (int64_t)ls * pitch + (rowid0 * nx + i0 + c0) * CPC;
          if (vec_ok) {
#pragma unroll
            for (int q = 0; q < CLS_CZ; ++q)
              if (q < nlay) cls_store<NQ>(out + q * step, e);
          } else {
            for (int q = 0; q < nlay; ++q, out += step) cls_store_any<NQ>(out, e, nl * CPC);
          }
          mine = (unsigned int)(nl * nlay);
        }
        const unsigned int tot = __reduce_add_sync(FULL, mine);
        if (lane == 0 && tot) atomicAdd(&s_uni[ls][uf == 1u ? 0 : 1], tot);
      }
    }
  }
  __syncthreads();
  // one-signed tiles count for every cell type of the n-cube
  if (threadIdx.x < LH * 2 * CPC) {
    const int ls = threadIdx.x / (2 * CPC), io = (threadIdx.x / CPC) & 1, t = threadIdx.x % CPC;
    const unsigned int v = s_uni[ls][io];
    if (v) atomicAdd(&bg_counts[ls * 12 + io * 6 + t], (unsigned long long)v);
  }
}


// ---- compaction of the cut mask -------------------------------------------------------------------
// One thread per 32-cube group.  The scan runs over the per-ROW totals (k_classify adds every crossed group to its row);
// a group that holds cut cells finds its first slot as row prefix + the counts of the groups before it in the row -- the
// counts and mask words of empty groups are never read beyond the group's own count.  Cells leave in (cube, cell type)
// order = ascending cell id.
// T22: one level set on TET cells -- also hand the fill kernel the base of every tile of 32 consecutive cut cells: the
// number of 2-2 splits before the tile (row prefix + groups before + the 2-2 bits of the cells of this group before it)
template <int CPC, bool T22>
__global__ void __launch_bounds__(256)
k_cutmask_emit(const uint32_t* __restrict__ cutmask, int64_t ngroups, const uint32_t* __restrict__ cutcount,
               const uint32_t* __restrict__ rowoffs, int64_t nrows, int nx, int wxtot, int32_t* __restrict__ cutcells,
               const uint32_t* __restrict__ cut22mask, const uint32_t* __restrict__ cut22count,
               uint32_t* __restrict__ tile_n22) {
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = gidx < ngroups;
  const uint32_t mine = in_range ? cutcount[gidx] : 0u;
  const uint32_t mine22 = (T22 && in_range) ? cut22count[gidx] : 0u;
  const int64_t rowid = in_range ? gidx / wxtot : 0;
  const int w = (int)(gidx - rowid * wxtot);
  uint32_t o = 0u, b = 0u;
  if (wxtot <= 32 && (32 % wxtot) == 0) {
    // the groups of a row are wxtot consecutive lanes (ngroups and the block size are multiples of wxtot): the counts before
    // the group in its row come from a segmented warp scan instead of w loads
    const int lane = threadIdx.x & 31, pos = lane & (wxtot - 1);
    uint32_t x = mine | (mine22 << 16);  // <= 192 cells per group, <= 32 groups: both halves stay below 2^16
    for (int d = 1; d < wxtot; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, x, d);
      if (pos >= d) x += up;
    }
    x -= mine | (mine22 << 16);
    if (!in_range || mine == 0u) return;
    o = rowoffs[rowid] + (x & 0xFFFFu);
    if (T22) b = rowoffs[nrows + rowid] + (x >> 16);
  } else {
    if (!in_range || mine == 0u) return;
    o = rowoffs[rowid];
    b = T22 ? rowoffs[nrows + rowid] : 0u;
    for (int q = 0; q < w; ++q) {
      o += cutcount[rowid * wxtot + q];
      if (T22) b += cut22count[rowid * wxtot + q];
    }
  }
  uint32_t m[CPC], s[T22 ? CPC : 1];
  uint32_t any = 0u;
#pragma unroll
  for (int t = 0; t < CPC; ++t) {
    m[t] = cutmask[gidx * CPC + t];
    any |= m[t];
  }
  if (T22) {
    const bool has = mine22 != 0u;  // k_classify wrote the bits of such groups only
#pragma unroll
    for (int t = 0; t < CPC; ++t) s[t] = has ? cut22mask[gidx * CPC + t] : 0u;
  }
  const int64_t cube0 = rowid * nx + 32 * (int64_t)w;
  while (any) {
    const int c = __ffs(any) - 1;
    any &= any - 1;
#pragma unroll
    for (int t = 0; t < CPC; ++t)
      if ((m[t] >> c) & 1u) {
        if (T22) {
          if ((o & 31u) == 0u) tile_n22[o >> 5] = b;
          b += (s[t] >> c) & 1u;
        }
        cutcells[o++] = (int32_t)((cube0 + c) * CPC + t + 1);
      }
  }
}

// Sign words of a NODAL level set (DiscreteGeometry): 8 bytes per node are read exactly once with eight independent
// 256-byte warp loads in flight per warp, one bit per node is written.  word (row, wx) covers the nodes 32*wx .. 32*wx+31
// of node row `row` (rows = the slab's node planes k0..k1 x the node rows of a plane); bits of nodes past the row end are 0.
constexpr int NSG_WPI = 8;    // words per warp iteration
constexpr int NSG_ITERS = 2;  // iterations per warp

__global__ void __launch_bounds__(256)
k_nodal_signs(const double* __restrict__ vals /* first node of the slab's first plane */, int nn0, int wpr, int64_t nwords,
              uint32_t* __restrict__ signs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t W0 = warp * (NSG_WPI * NSG_ITERS);
#pragma unroll 1
  for (int it = 0; it < NSG_ITERS; ++it, W0 += NSG_WPI) {
    if (W0 >= nwords) return;
    double v[NSG_WPI];
    int64_t row = W0 / wpr;  // one division per iteration, then incremental (row, wx)
    int wx = (int)(W0 - row * wpr);
    const double* rp = vals + row * nn0;
#pragma unroll
    for (int q = 0; q < NSG_WPI; ++q) {
      const int i = wx * 32 + lane;
      v[q] = (W0 + q < nwords && i < nn0) ? __ldg(rp + i) : -1.0;
      if (++wx == wpr) {
        wx = 0;
        rp += nn0;
      }
    }
    uint32_t mine = 0u;
#pragma unroll
    for (int q = 0; q < NSG_WPI; ++q) {
      const uint32_t word = __ballot_sync(0xffffffffu, v[q] > 0.0);
      if (lane == q) mine = word;
    }
    if (lane < NSG_WPI && W0 + lane < nwords) signs[W0 + lane] = mine;
  }
}

// `plane0`: index (last direction) of the first node plane written, `total` nodes from there on (whole model: 0 / all)
template <int D>
__global__ void k_eval_leaf_nodes(const GcGrid g, const gecut_leaf leaf, int plane0, int64_t total, double* __restrict__ out) {
  int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= total) return;
  int i = (int)(node % g.nn[0]);
  int j = (D == 3) ? (int)((node / g.nn[0]) % g.nn[1]) : (int)(node / g.nn[0]) + plane0;
  int k = (D == 3) ? (int)(node / ((int64_t)g.nn[0] * g.nn[1])) + plane0 : 0;
  double x[3] = {gc_coord(g, 0, i), gc_coord(g, 1, j), D == 3 ? gc_coord(g, 2, k) : 0.0};
  out[node] = gc_eval_leaf(leaf, x, D);
}

}  // namespace

// number of layers one warp tile marches: few enough tiles per resident warp leave SMs idle at the end (tail), short
// marches re-evaluate the shared node plane more often ((ZS+1)/ZS); pick the best product of the two efficiencies
static int cls_pick_zs(int64_t tiles_per_slab, int nlayers, int64_t resident_warps) {
  int best = 8;
  double best_score = -1.0;
  for (int zs = 4; zs <= 32; ++zs) {
    const int64_t n = tiles_per_slab * gc_div_up(nlayers, zs);
    const double waves = (double)n / (double)resident_warps;
    const double tail = waves / ceil(waves);
    const double score = tail * (double)zs / (double)(zs + 1);
    if (score > best_score + 1e-9) {
      best_score = score;
      best = zs;
    }
  }
  return best;
}

int gc_classify(gecut_ctx* ctx, gecut_result* res) {
  const GcGrid& g = res->grid;
  const int L = res->L;
  const int nx = g.n[0];
  const int wxtot = gc_words_per_row(nx);
  const int nlayers = g.k1 - g.k0;
  const int64_t rows = (g.D == 3) ? (int64_t)nlayers * g.n[1] : (int64_t)nlayers;
  GC_CHECK(gc_malloc(ctx, (void**)&res->cutcount, sizeof(uint32_t) * rows * wxtot * (res->want_n22 ? 2 : 1)));
  GC_CUDA(cudaMemsetAsync(res->cutcount, 0, sizeof(uint32_t) * rows * wxtot * (res->want_n22 ? 2 : 1), ctx->stream));
  uint32_t* cut22 = res->want_n22 ? res->cutcount + rows * wxtot : nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&res->rowcount, sizeof(uint32_t) * rows * 2));
  GC_CUDA(cudaMemsetAsync(res->rowcount, 0, sizeof(uint32_t) * rows * 2, ctx->stream));
  if (res->want_n22) GC_CHECK(gc_malloc(ctx, (void**)&res->cut22mask, sizeof(uint32_t) * rows * wxtot * 6));
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
  // More than two level sets, all in closed form: chunks of two go through the hoisted kernel (LH = 2), each chunk
  // writing its own state rows and OR-ing its cut bits into the mask -- ~3x fewer FP64 operations per node than the
  // generic evaluation of all L leaves.  NODAL / doughnut leaves keep the single generic launch.
  bool closed_form = true;
  for (int i = 0; i < L; ++i) {
    const int k = res->leaves.leaf[i].kind;
    closed_form = closed_form && (k == GECUT_LEAF_SPHERE || k == GECUT_LEAF_PLANE || k == GECUT_LEAF_CYLINDER);
  }
  const int chunk = (L > 2 && closed_form) ? 2 : L;
  // NODAL leaves: sign words first (streaming pass), the marching kernel then reads 1 bit per node instead of 8 bytes
  const int swpr = (g.nn[0] + 31) / 32;
  const int64_t srows = (int64_t)(nlayers + 1) * ((g.D == 3) ? g.nn[1] : 1);
  const int64_t swords = srows * swpr;
  uint32_t* sign_pool = nullptr;
  const uint32_t* sign_of[GC_LMAX] = {};
  {
    int nnod = 0;
    for (int i = 0; i < L; ++i) nnod += res->leaves.leaf[i].kind == GECUT_LEAF_NODAL ? 1 : 0;
    if (nnod > 0) {
      GC_CHECK(gc_malloc(ctx, (void**)&sign_pool, sizeof(uint32_t) * swords * nnod));
      const int64_t layer_nodes = (int64_t)g.nn[0] * ((g.D == 3) ? g.nn[1] : 1);
      const unsigned nbs = (unsigned)gc_div_up(swords, (int64_t)(256 / 32) * NSG_WPI * NSG_ITERS);
      int q = 0;
      for (int i = 0; i < L; ++i) {
        if (res->leaves.leaf[i].kind != GECUT_LEAF_NODAL) continue;
        uint32_t* dst = sign_pool + (int64_t)q++ * swords;
        const double* first = res->leaves.nodal[i] + ((int64_t)g.k0 * layer_nodes - res->leaves.nodal_base);
        GC_LAUNCH(ctx, "k_nodal_signs", k_nodal_signs<<<nbs, 256, 0, ctx->stream>>>(first, g.nn[0], swpr, swords, dst));
        sign_of[i] = dst;
      }
      GC_CHECK(gc_launch_check(ctx, "k_nodal_signs"));
    }
  }
  auto launch = [&](auto Dtag, auto Stag) {
    constexpr int D = decltype(Dtag)::value;
    constexpr bool S = decltype(Stag)::value;
    using C = ClsCfg<D, S>;
    const int ntx = (int)gc_div_up(wxtot, C::WXT);
    const int nty = (D == 3) ? (int)gc_div_up(g.n[1], C::TY) : 1;
    const int zs = cls_pick_zs((int64_t)ntx * nty, nlayers, (int64_t)nsm * CLS_MIN_BLOCKS * C::WARPS);
    const int64_t ntiles = (int64_t)ntx * nty * gc_div_up(nlayers, zs);
    const unsigned nb = (unsigned)gc_div_up(ntiles, C::WARPS);
    for (int ls0 = 0; ls0 < L; ls0 += chunk) {
      const int nl = std::min(chunk, L - ls0);
      GcLeaves lv{};  // the chunk's leaves as level sets 0..nl-1 of this launch
      lv.L = nl;
      lv.nodal_base = res->leaves.nodal_base;
      for (int i = 0; i < nl; ++i) {
        lv.leaf[i] = res->leaves.leaf[ls0 + i];
        lv.nodal[i] = res->leaves.nodal[ls0 + i];
        lv.signs[i] = sign_of[ls0 + i];
      }
      lv.signs_wpr = swpr;
      const int LH = nl <= 2 ? nl : 0;
      const int rowd = 1 + 2 * LH;
      const size_t per_warp = (sizeof(double) * C::RY * rowd + sizeof(uint32_t) * 2 * nl * C::PW + 15) & ~(size_t)15;
      const size_t smem = per_warp * C::WARPS;
      int8_t* states = res->lay.bg_state + (int64_t)ls0 * res->lay.bg_pitch;
      unsigned long long* counts = res->bg_counts_dev + (int64_t)ls0 * 12;
      const int acc = ls0 > 0 ? 1 : 0;
      bool any_signs = false;
      for (int i = 0; i < nl; ++i) any_signs = any_signs || lv.signs[i] != nullptr;
      // closed-form leaves in the hoisted kernels: cull first (k_classify_cull writes the one-signed tiles of CLS_CZ layers at
      // store speed and lists the others), then march only the listed tiles
      bool cull = !any_signs && LH > 0;
      for (int i = 0; i < nl; ++i) {
        const int k = lv.leaf[i].kind;
        cull = cull && (k == GECUT_LEAF_SPHERE || k == GECUT_LEAF_PLANE || k == GECUT_LEAF_CYLINDER);
      }
      uint32_t* wl = nullptr;
      int zs_l = zs;
      int64_t ntiles_l = ntiles;
      unsigned nb_l = nb;
      if (cull) {
        zs_l = CLS_CZ;
        ntiles_l = (int64_t)ntx * nty * gc_div_up(nlayers, CLS_CZ);
        nb_l = (unsigned)gc_div_up(ntiles_l, C::WARPS);
        GC_CHECK(gc_malloc(ctx, (void**)&wl, sizeof(uint32_t) * (ntiles_l + 1)));
        GC_CUDA(cudaMemsetAsync(wl + ntiles_l, 0, sizeof(uint32_t), ctx->stream));
        const unsigned nbc = (unsigned)gc_div_up(ntiles_l, 256 / 32);
#define GC_CULL(LHV)                                                                                                  \
  GC_LAUNCH(ctx, "k_classify_cull", k_classify_cull<D, S, LHV><<<nbc, 256, 0, ctx->stream>>>(                          \
      g, lv, states, res->lay.bg_pitch, counts, ntx, nty, ntiles_l, wl, wl + ntiles_l))
        if (nl == 1) GC_CULL(1); else GC_CULL(2);
#undef GC_CULL
      }
#define GC_CLS(LHV, SGV)                                                                                              \
  GC_LAUNCH(ctx, "k_classify", k_classify<D, S, LHV, SGV><<<nb_l, C::WARPS * 32, smem, ctx->stream>>>(                 \
      g, lv, states, res->lay.bg_pitch, res->cutmask, res->cutcount, cut22, res->cut22mask, res->rowcount, rows, ctx->d_simplexify_raw, counts, zs_l, ntx, nty, \
      ntiles_l, acc, wl, wl ? wl + ntiles_l : nullptr))
      // the sign-word branch is compiled out of the closed-form instances (it cost the tet kernel 24 bytes of spills)
      if (any_signs) {
        if (nl == 1) GC_CLS(1, true);
        else if (nl == 2) GC_CLS(2, true);
        else GC_CLS(0, true);
      } else {
        if (nl == 1) GC_CLS(1, false);
        else if (nl == 2) GC_CLS(2, false);
        else GC_CLS(0, false);
      }
#undef GC_CLS
      gc_free(ctx, wl);  // stream-ordered cache: not handed out again before the two launches have run
    }
    return (int)GECUT_OK;
  };
  int lst;
  if (g.D == 3) {
    if (g.simplexified)
      lst = launch(std::integral_constant<int, 3>{}, std::true_type{});
    else
      lst = launch(std::integral_constant<int, 3>{}, std::false_type{});
  } else {
    if (g.simplexified)
      lst = launch(std::integral_constant<int, 2>{}, std::true_type{});
    else
      lst = launch(std::integral_constant<int, 2>{}, std::false_type{});
  }
  gc_free(ctx, sign_pool);  // stream-ordered cache: the block is not handed out before the launches above have run
  GC_CHECK(lst);
  return gc_launch_check(ctx, "k_classify");
}

int gc_compact_cutcells(gecut_ctx* ctx, gecut_result* res) {
  const GcGrid& g = res->grid;
  const int nx = g.n[0];
  const int wxtot = gc_words_per_row(nx);
  const int64_t rows = (g.D == 3) ? (int64_t)(g.k1 - g.k0) * g.n[1] : (int64_t)(g.k1 - g.k0);
  const int64_t ngroups = rows * wxtot;
  uint32_t* counts = res->cutcount;  // cut cells per 32-cube group, written by k_classify
  uint64_t* total_dev = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&total_dev, sizeof(uint64_t) * 2));
  const int T = 256;
  const unsigned nb = (unsigned)gc_div_up(ngroups, T);
  // channel 0: cut cells per row of cubes; channel 1: TET cells split 2-2 per row -- one launch for both
  uint32_t* rowoffs = res->rowcount;
  GC_CHECK(gc_exclusive_scan_u32_multi(ctx, rowoffs, rowoffs, rows, 2, total_dev));
  uint64_t* h = (uint64_t*)ctx->h_pinned;
  h[1] = 0;
  GC_CUDA(cudaMemcpyAsync(h, total_dev, sizeof(uint64_t) * 2, cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaMemcpyAsync(h + 2, res->bg_counts_dev, sizeof(uint64_t) * res->L * 12, cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t ncut = (int64_t)h[0];
  res->n22 = (int64_t)h[1];
  for (int i = 0; i < res->L * 12; ++i) res->bg_counts[i] = (int64_t)h[2 + i];
  res->sizes.num_cutcells = ncut;
  if (ncut > 0) {
    GC_CHECK(gc_malloc(ctx, (void**)&res->lay.cutcells, sizeof(int32_t) * ncut));
    switch (g.cpc) {
      case 1: GC_LAUNCH(ctx, "k_cutmask_emit", k_cutmask_emit<1, false><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts, rowoffs, rows, nx, wxtot, res->lay.cutcells, nullptr, nullptr, nullptr)); break;
      case 2: GC_LAUNCH(ctx, "k_cutmask_emit", k_cutmask_emit<2, false><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts, rowoffs, rows, nx, wxtot, res->lay.cutcells, nullptr, nullptr, nullptr)); break;
      default:
        if (res->want_n22) {
          GC_CHECK(gc_malloc(ctx, (void**)&res->tile_n22, sizeof(uint32_t) * gc_div_up(ncut, 32)));
          GC_LAUNCH(ctx, "k_cutmask_emit", k_cutmask_emit<6, true><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts, rowoffs, rows, nx, wxtot, res->lay.cutcells, res->cut22mask, counts + ngroups, res->tile_n22));
        } else {
          GC_LAUNCH(ctx, "k_cutmask_emit", k_cutmask_emit<6, false><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts, rowoffs, rows, nx, wxtot, res->lay.cutcells, nullptr, nullptr, nullptr));
        }
        break;
    }
  }
  int st = gc_launch_check(ctx, "cutmask compaction");
  gc_free(ctx, res->cutcount);
  res->cutcount = nullptr;
  gc_free(ctx, res->cut22mask);
  res->cut22mask = nullptr;
  gc_free(ctx, res->rowcount);
  res->rowcount = nullptr;
  gc_free(ctx, total_dev);
  return st;
}

int gc_eval_leaf_nodes(gecut_ctx* ctx, const GcGrid& g, const gecut_leaf& leaf, double* out_dev, bool slab_only) {
  // whole model, or only the node planes k0..k1 of the slab (slab-local DiscreteGeometry vectors)
  const int plane0 = slab_only ? g.k0 : 0;
  const int64_t total = slab_only ? g.num_nodes : g.layer_nodes * (int64_t)g.nn[g.D - 1];
  const int T = 256;
  unsigned nb = (unsigned)gc_div_up(total, T);
  if (g.D == 3)
    GC_LAUNCH(ctx, "k_eval_leaf_nodes", k_eval_leaf_nodes<3><<<nb, T, 0, ctx->stream>>>(g, leaf, plane0, total, out_dev));
  else
    GC_LAUNCH(ctx, "k_eval_leaf_nodes", k_eval_leaf_nodes<2><<<nb, T, 0, ctx->stream>>>(g, leaf, plane0, total, out_dev));
  return gc_launch_check(ctx, "k_eval_leaf_nodes");
}
