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

// all three kernels take `nch` arrays of n elements stored back to back (blockIdx.y = channel)
__global__ void k_scan_reduce(const uint32_t* __restrict__ in, int64_t n, uint64_t* __restrict__ blocksums) {
  in += (int64_t)blockIdx.y * n;
  blocksums += (int64_t)blockIdx.y * gridDim.x;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  uint32_t tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) blocksums[blockIdx.x] = tot;
}

// single block: exclusive scan of the block sums (uint64), grand total out
__global__ void k_scan_blocksums(uint64_t* __restrict__ blocksums, int64_t nb, uint64_t* __restrict__ total) {
  blocksums += (int64_t)blockIdx.x * nb;
  total += blockIdx.x;
  __shared__ uint64_t carry_s;
  __shared__ uint64_t wsum[32];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t start = 0; start < nb; start += blockDim.x) {
    int64_t idx = start + threadIdx.x;
    uint64_t v = idx < nb ? blocksums[idx] : 0;
    uint64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      uint64_t s = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
      uint64_t si = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, si, o);
        if (lane >= o) si += t;
      }
      wsum[lane] = si - s;
    }
    __syncthreads();
    uint64_t excl = carry_s + wsum[wid] + inc - v;
    if (idx < nb) blocksums[idx] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

__global__ void k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
                             const uint64_t* __restrict__ blocksums) {
  in += (int64_t)blockIdx.y * n;
  out += (int64_t)blockIdx.y * n;
  blocksums += (int64_t)blockIdx.y * gridDim.x;
  // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS) of the tile so that the order is preserved
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  uint32_t tot;
  uint32_t excl = block_excl_scan(s, &tot);
  uint32_t run = (uint32_t)blocksums[blockIdx.x] + excl;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = run;
    run += v[i];
  }
}

}  // namespace

int gc_exclusive_scan_u32_multi(gecut_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int nch, uint64_t* totals_dev) {
  if (n <= 0) {
    GC_CUDA(cudaMemsetAsync(totals_dev, 0, sizeof(uint64_t) * nch, ctx->stream));
    return GECUT_OK;
  }
  int64_t nb = gc_div_up(n, SCAN_TILE);
  uint64_t* blocksums = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&blocksums, sizeof(uint64_t) * nb * nch));
  dim3 grid((unsigned)nb, (unsigned)nch);
  GC_LAUNCH(ctx, "k_scan_reduce", k_scan_reduce<<<grid, SCAN_THREADS, 0, ctx->stream>>>(in, n, blocksums));
  GC_LAUNCH(ctx, "k_scan_blocksums", k_scan_blocksums<<<nch, 1024, 0, ctx->stream>>>(blocksums, nb, totals_dev));
  GC_LAUNCH(ctx, "k_scan_apply", k_scan_apply<<<grid, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, blocksums));
  int st = gc_launch_check(ctx, "scan");
  gc_free(ctx, blocksums);
  return st;
}

int gc_exclusive_scan_u32(gecut_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, uint64_t* total_dev) {
  return gc_exclusive_scan_u32_multi(ctx, in, out, n, 1, total_dev);
}

// ------------------------------------------------------------------------------------------------
// fused node evaluation + classification
// ------------------------------------------------------------------------------------------------
namespace {

template <int D>
struct ClsCfg {
  static constexpr int WX = (D == 3) ? 4 : 16;  // x words (32 cells each) per block tile
  static constexpr int TY = (D == 3) ? 8 : 1;   // cell rows per tile (3-D only)
  static constexpr int RY = (D == 3) ? TY + 1 : 1;  // node rows per plane
  static constexpr int ZS = 32;                 // layers of the last direction marched by one block
  static constexpr int THREADS = (D == 3) ? 256 : 128;
};

// words per x-row of the cut mask
__host__ __device__ inline int gc_words_per_row(int nx) { return (nx + 31) / 32; }

// LH > 0: "hoisted" evaluation for models with at most LH level sets.  Every closed-form leaf splits into a part that
// depends on the in-plane coordinates only (constant while the block marches through the layers) and the marching
// coordinate: e.g. sphere ((wx*wx + wy*wy) + wz*wz) - R*R  =  (pa + wz*wz) - RR with pa kept in a register per task.
// The association of the reference formulas is preserved exactly (same rounding), only loop-invariant operations move.
template <int D, bool SIMPLEX, int LH>
__global__ void __launch_bounds__(ClsCfg<D>::THREADS)
k_classify(const GcGrid g, const GcLeaves lv, int8_t* __restrict__ states, int64_t pitch,
           uint32_t* __restrict__ cutmask, const uint8_t* __restrict__ simplexify_raw,
           unsigned long long* __restrict__ bg_counts /* [L][2][6]: IN / OUT cells per level set and cell type */) {
  using C = ClsCfg<D>;
  constexpr int WX = C::WX, TY = C::TY, RY = C::RY, ZS = C::ZS;
  constexpr int NVC = 1 << D;                         // vertices of the n-cube
  constexpr int CPC = SIMPLEX ? (D == 3 ? 6 : 2) : 1;  // bg cells per n-cube
  constexpr int NW = C::THREADS / 32;
  constexpr int NTASK = TY * WX;                      // (row, x-word) tasks of one layer: 32 cells each
  static_assert(NTASK * 8 == C::THREADS, "phase B maps one thread to 4 cells");
  extern __shared__ uint32_t planes[];  // [2][L][RY][WX+1] sign words
  __shared__ unsigned long long s_lut[SIMPLEX ? (1 << NVC) : 1];  // sign byte of an n-cube -> states of its CPC simplices
  __shared__ uint8_t s_tets[6][4];
  __shared__ unsigned int s_cnt[GC_LMAX * 12];
  const int L = lv.L;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < L * 12; i += C::THREADS) s_cnt[i] = 0u;
  const int nx = g.n[0];
  const int ny = (D == 3) ? g.n[1] : 1;
  const int wxtot = gc_words_per_row(nx);
  const int w0 = blockIdx.x * WX;   // first x word of the tile
  const int j0 = (D == 3) ? blockIdx.y * TY : 0;
  const int kz0 = g.k0 + ((D == 3) ? blockIdx.z : blockIdx.y) * ZS;  // first layer (global index)
  const int kz1 = min(kz0 + ZS, g.k1);
  const bool vec_ok = (nx & 3) == 0;
  if (SIMPLEX) {
    if (threadIdx.x < 24) s_tets[threadIdx.x / 4][threadIdx.x % 4] = simplexify_raw[(D - 2) * 24 + threadIdx.x];
    __syncthreads();
    for (int sb = threadIdx.x; sb < (1 << NVC); sb += C::THREADS) {
      unsigned long long e = 0ull;
      for (int t = 0; t < CPC; ++t) {
        bool any = false, all = true;
        for (int m = 0; m <= D; ++m) {
          const bool o = (sb >> s_tets[t][m]) & 1;
          any |= o;
          all &= o;
        }
        const unsigned long long st = all ? 0x01ull : (any ? 0x00ull : 0xFFull);  // OUT / CUT / IN
        e |= st << (8 * t);
      }
      s_lut[sb] = e;
    }
  }
  const int plane_words = L * RY * (WX + 1);
  // packed per-type tallies (one byte per cell type; <= 4 cells x 32 layers = 128 per byte, n-cubes: plain counters)
  unsigned long long acc_in = 0ull, acc_out = 0ull;
  // hoisted (plane-invariant) parts of the leaf formulas per evaluation task of this warp
  constexpr int NSLOT = (RY * WX + 1 + NW - 1) / NW;
  constexpr int LHS = LH > 0 ? LH : 1;
  double h_pa[NSLOT][LHS], h_pb[NSLOT][LHS], h_c[LHS];
  uint32_t h_valid = 0u;  // bit k: this lane evaluates a real node in task slot k
  if (LH > 0) {
#pragma unroll
    for (int ls = 0; ls < LH; ++ls) {
      h_c[ls] = 0.0;
      if (ls < L) {
        const int kind = lv.leaf[ls].kind;
        h_c[ls] = kind == GECUT_LEAF_DOUGHNUT ? __dmul_rn(lv.leaf[ls].r, lv.leaf[ls].r) : __dmul_rn(lv.leaf[ls].R, lv.leaf[ls].R);
      }
    }
#pragma unroll
    for (int k = 0; k < NSLOT; ++k) {
      const int task = wid + k * NW;
      const bool halo = task == RY * WX;
      const int r = halo ? lane : task / WX;
      const int ww = halo ? WX : task % WX;
      const int i = halo ? 32 * (w0 + WX) : 32 * (w0 + ww) + lane;
      const int j = j0 + r;
      if ((task < RY * WX + 1) && (i < g.nn[0]) && (D == 2 || j < g.nn[1]) && (!halo || lane < RY)) h_valid |= 1u << k;
      const double xc = gc_coord(g, 0, i);
      const double yc = (D == 3) ? gc_coord(g, 1, j) : 0.0;
#pragma unroll
      for (int ls = 0; ls < LH; ++ls) {
        h_pa[k][ls] = 0.0;
        h_pb[k][ls] = 0.0;
        if (ls < L) {
          const gecut_leaf& lf = lv.leaf[ls];
          const double wx = __dsub_rn(xc, lf.x0[0]);
          const double wy = (D == 3) ? __dsub_rn(yc, lf.x0[1]) : 0.0;
          // in-plane partial sums, associated exactly like gc_dot3: (a0*b0 + a1*b1) [+ a2*b2 added per plane]
          const double ww2 = (D == 3) ? __dadd_rn(__dmul_rn(wx, wx), __dmul_rn(wy, wy)) : __dmul_rn(wx, wx);
          const double wv = (D == 3) ? __dadd_rn(__dmul_rn(wx, lf.v[0]), __dmul_rn(wy, lf.v[1])) : __dmul_rn(wx, lf.v[0]);
          if (lf.kind == GECUT_LEAF_SPHERE) {
            h_pa[k][ls] = ww2;
          } else if (lf.kind == GECUT_LEAF_PLANE) {
            h_pa[k][ls] = wv;
          } else if (lf.kind == GECUT_LEAF_CYLINDER) {
            h_pa[k][ls] = wv;
            h_pb[k][ls] = ww2;
          } else if (lf.kind == GECUT_LEAF_DOUGHNUT) {
            const double t = __dsub_rn(lf.R, __dsqrt_rn(ww2));
            h_pa[k][ls] = __dmul_rn(t, t);
          }
        }
      }
    }
  }

  for (int kp = kz0; kp <= kz1; ++kp) {  // node planes kz0..kz1
    uint32_t* P = planes + ((kp - kz0) & 1) * plane_words;
    if (LH > 0) {
      // ---- hoisted evaluation: per plane only the marching coordinate enters
      const double zc = gc_coord(g, D - 1, kp);
      double wz[LH > 0 ? LH : 1], wz2[LH > 0 ? LH : 1], wzv[LH > 0 ? LH : 1];
#pragma unroll
      for (int ls = 0; ls < LH; ++ls) {
        if (ls < L) {
          wz[ls] = __dsub_rn(zc, lv.leaf[ls].x0[D - 1]);
          wz2[ls] = __dmul_rn(wz[ls], wz[ls]);
          wzv[ls] = __dmul_rn(wz[ls], lv.leaf[ls].v[D - 1]);
        }
      }
#pragma unroll
      for (int k = 0; k < NSLOT; ++k) {
        const int task = wid + k * NW;
        if (task < RY * WX + 1) {
          const bool halo = task == RY * WX;
          const int r = halo ? lane : task / WX;
          const int ww = halo ? WX : task % WX;
#pragma unroll
          for (int ls = 0; ls < LH; ++ls) {
            if (ls < L) {
              const int kind = lv.leaf[ls].kind;
              double v;
              if (kind == GECUT_LEAF_SPHERE) {
                v = __dsub_rn(__dadd_rn(h_pa[k][ls], wz2[ls]), h_c[ls]);
              } else if (kind == GECUT_LEAF_PLANE) {
                v = __dadd_rn(h_pa[k][ls], wzv[ls]);
              } else if (kind == GECUT_LEAF_CYLINDER) {
                const double A = __dadd_rn(h_pa[k][ls], wzv[ls]);
                const double H2 = __dadd_rn(h_pb[k][ls], wz2[ls]);
                v = __dsub_rn(__dsub_rn(H2, __dmul_rn(A, A)), h_c[ls]);
              } else if (kind == GECUT_LEAF_DOUGHNUT) {
                v = __dsub_rn(__dadd_rn(h_pa[k][ls], wz2[ls]), h_c[ls]);
              } else {
                const int i = halo ? 32 * (w0 + WX) : 32 * (w0 + ww) + lane;
                const int64_t node = (D == 3) ? ((int64_t)i + (int64_t)g.nn[0] * ((int64_t)(j0 + r) + (int64_t)g.nn[1] * kp))
                                              : ((int64_t)i + (int64_t)g.nn[0] * kp);
                v = ((h_valid >> k) & 1u) ? __ldg(lv.nodal[ls] + (node - lv.nodal_base)) : -1.0;
              }
              const uint32_t word = __ballot_sync(0xffffffffu, ((h_valid >> k) & 1u) && (v > 0.0));
              if (!halo) {
                if (lane == 0) P[(ls * RY + r) * (WX + 1) + ww] = word;
              } else if (lane < RY) {
                P[(ls * RY + lane) * (WX + 1) + WX] = (word >> lane) & 1u;
              }
            }
          }
        }
      }
    } else {
    // ---- evaluate the signs of node plane kp: one ballot per (row, word) task and level set.  Of the word right of the
      //      tile only bit 0 is needed (x+1 vertices of the last cell column): one transposed task evaluates that node
      //      column for all RY rows at once instead of RY mostly wasted row tasks.
      for (int task = wid; task < RY * WX + 1; task += NW) {
        const bool halo = task == RY * WX;
        const int r = halo ? lane : task / WX;
        const int ww = halo ? WX : task % WX;
        const int i = halo ? 32 * (w0 + WX) : 32 * (w0 + ww) + lane;
        const int j = j0 + r;
        const bool valid = (i < g.nn[0]) && (D == 2 || j < g.nn[1]) && (!halo || lane < RY);
        double x[3];
        x[0] = gc_coord(g, 0, i);
        if (D == 3) {
          x[1] = gc_coord(g, 1, j);
          x[2] = gc_coord(g, 2, kp);
        } else {
          x[1] = gc_coord(g, 1, kp);
          x[2] = 0.0;
        }
        const int64_t node = (D == 3) ? ((int64_t)i + (int64_t)g.nn[0] * ((int64_t)j + (int64_t)g.nn[1] * kp))
                                      : ((int64_t)i + (int64_t)g.nn[0] * kp);
        for (int ls = 0; ls < L; ++ls) {
          double v = -1.0;
          if (valid) {
            if (lv.leaf[ls].kind == GECUT_LEAF_NODAL)
              v = __ldg(lv.nodal[ls] + (node - lv.nodal_base));
            else
              v = gc_eval_leaf(lv.leaf[ls], x, D);
          }
          const uint32_t word = __ballot_sync(0xffffffffu, valid && (v > 0.0));
          if (!halo) {
            if (lane == 0) P[(ls * RY + r) * (WX + 1) + ww] = word;
          } else if (lane < RY) {
            P[(ls * RY + lane) * (WX + 1) + WX] = (word >> lane) & 1u;
          }
        }
      }
    }
    __syncthreads();
    if (kp > kz0) {
      // ---- classify layer kp-1 from planes (kp-1) [lo] and kp [hi]
      const uint32_t* P0 = planes + ((kp - 1 - kz0) & 1) * plane_words;
      const uint32_t* P1 = P;
      const int kc = kp - 1;  // global layer index
      for (int ls = 0; ls < L; ++ls) {
        // One thread per 4 consecutive cells (8 threads share a 32-cell group).  Every thread rebuilds the vertex sign
        // words of its group from the two planes (16 mostly broadcast shared loads) -- redundant across the 8 threads
        // but parallel, which removes the former single-warp "phase A" and one block barrier per layer and level set.
        const int task = threadIdx.x >> 3, quad = threadIdx.x & 7;
        const int r = task / WX, ww = task % WX;
        const int jc = j0 + r;
        const int i0 = 32 * (w0 + ww);
        const int c0 = 4 * quad;
        unsigned long long pin = 0ull, pout = 0ull;  // IN / OUT tallies, one byte per cell type
        if (i0 < nx && jc < ny) {
          uint32_t V[NVC];
#pragma unroll
          for (int v = 0; v < NVC; ++v) {
            const int xo = v & 1;
            const int ro = (D == 3) ? ((v >> 1) & 1) : 0;
            const int po = (D == 3) ? ((v >> 2) & 1) : ((v >> 1) & 1);
            const uint32_t* row = (po ? P1 : P0) + (ls * RY + r + ro) * (WX + 1);
            uint32_t a = row[ww];
            if (xo) a = (a >> 1) | (row[ww + 1] << 31);
            V[v] = a;
          }
          const int ncell = min(32, nx - i0);
          const uint32_t live = ncell == 32 ? 0xffffffffu : ((1u << ncell) - 1u);
          const int64_t rowid = (D == 3) ? ((int64_t)(kc - g.k0) * ny + jc) : (int64_t)(kc - g.k0);
          uint32_t* cm = cutmask + (rowid * wxtot + (w0 + ww)) * CPC;
          // ---- "cut by any level set" mask of the group: thread `quad` owns cell type t = quad
          if (quad < CPC) {
            uint32_t any = 0u, all = 0xffffffffu;
            if (SIMPLEX) {
#pragma unroll
              for (int m = 0; m <= D; ++m) {
                const uint32_t a = V[s_tets[quad][m]];
                any |= a;
                all &= a;
              }
            } else {
#pragma unroll
              for (int v = 0; v < NVC; ++v) {
                any |= V[v];
                all &= V[v];
              }
            }
            const uint32_t cutw = any & ~all & live;
            if (L == 1 || ls == 0)
              cm[quad] = cutw;
            else if (cutw)
              cm[quad] |= cutw;  // the same thread owns this word for every ls of the layer
          }
          // ---- state bytes of this thread's 4 cells
          if (c0 < ncell) {
            const int ncq = min(4, ncell - c0);
            const int64_t cube0 = rowid * nx + i0 + c0;
            int8_t* out = states + (int64_t)ls * pitch + cube0 * CPC;
            if (SIMPLEX) {
              uint32_t vor = 0u, vand = 0xffffffffu;
#pragma unroll
              for (int v = 0; v < NVC; ++v) {
                vor |= V[v];
                vand &= V[v];
              }
              unsigned long long e[4];
              if ((vor & live) == 0u || (vand & live) == live) {
                // uniform group (the common case away from the surfaces): constant states, no gather
                constexpr unsigned long long ONES = CPC == 6 ? 0x010101010101ull : 0x0101ull;
                const bool in = (vor & live) == 0u;
                const unsigned long long pat = in ? ONES * 0xFFull : ONES;
#pragma unroll
                for (int c = 0; c < 4; ++c) e[c] = pat;
                if (in) pin = ONES * (unsigned long long)ncq; else pout = ONES * (unsigned long long)ncq;
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  uint32_t sb = 0;
#pragma unroll
                  for (int v = 0; v < NVC; ++v) sb |= ((V[v] >> (c0 + c)) & 1u) << v;
                  e[c] = s_lut[sb];
                  if (c < ncq) {
                    const unsigned long long hi = (e[c] >> 7) & 0x010101010101ull;  // bit 7 set: IN
                    pin += hi;
                    pout += (e[c] & 0x010101010101ull) & ~hi;                       // bit 0 set, bit 7 clear: OUT
                  }
                }
              }
              if (vec_ok) {
                if (CPC == 6) {  // 4 cells x 6 bytes = 24 bytes = three 8-byte stores
                  unsigned long long* o8 = reinterpret_cast<unsigned long long*>(out);
                  o8[0] = (e[0] & 0xFFFFFFFFFFFFull) | (e[1] << 48);
                  o8[1] = ((e[1] >> 16) & 0xFFFFFFFFull) | (e[2] << 32);
                  o8[2] = ((e[2] >> 32) & 0xFFFFull) | (e[3] << 16);
                } else {  // 4 cells x 2 bytes
                  *reinterpret_cast<unsigned long long*>(out) =
                      (e[0] & 0xFFFFull) | ((e[1] & 0xFFFFull) << 16) | ((e[2] & 0xFFFFull) << 32) | ((e[3] & 0xFFFFull) << 48);
                }
              } else {
                for (int c = 0; c < ncq; ++c)
                  for (int t = 0; t < CPC; ++t) out[c * CPC + t] = (int8_t)((e[c] >> (8 * t)) & 0xFFull);
              }
            } else {
              uint32_t any = 0u, all = 0xffffffffu;
#pragma unroll
              for (int v = 0; v < NVC; ++v) {
                any |= V[v];
                all &= V[v];
              }
              const uint32_t vm = (1u << ncq) - 1u;
              const uint32_t o4 = (all >> c0) & 0xFu, n4 = (~any >> c0) & 0xFu;
              pin = __popc(n4 & vm);
              pout = __popc(o4 & vm);
              const uint32_t so = (o4 * 0x00204081u) & 0x01010101u;
              const uint32_t si = ((n4 * 0x00204081u) & 0x01010101u) * 0xFFu;
              const uint32_t w = so | si;  // OUT(1) where all, IN(0xFF) where !any, CUT(0) otherwise
              if (vec_ok) {
                *reinterpret_cast<uint32_t*>(out) = w;
              } else {
                for (int c = 0; c < ncq; ++c) out[c] = (int8_t)((w >> (8 * c)) & 0xFFu);
              }
            }
          }
        }
        if (L == 1) {  // one level set: keep the packed tallies in registers until the end of the block
          acc_in += pin;
          acc_out += pout;
        } else {
#pragma unroll
          for (int t = 0; t < CPC; ++t) {
            const unsigned int a = __reduce_add_sync(0xffffffffu, (unsigned int)(pin >> (8 * t)) & 0xFFu);
            const unsigned int b = __reduce_add_sync(0xffffffffu, (unsigned int)(pout >> (8 * t)) & 0xFFu);
            if (lane == 0) {
              if (a) atomicAdd(&s_cnt[ls * 12 + t], a);
              if (b) atomicAdd(&s_cnt[ls * 12 + 6 + t], b);
            }
          }
        }
      }
    }
    // one barrier per layer: the next plane goes to the buffer this classification step has just read
    __syncthreads();
  }
  if (L == 1) {
#pragma unroll
    for (int t = 0; t < CPC; ++t) {
      const unsigned int vi = SIMPLEX ? (unsigned int)(acc_in >> (8 * t)) & 0xFFu : (unsigned int)acc_in;
      const unsigned int vo = SIMPLEX ? (unsigned int)(acc_out >> (8 * t)) & 0xFFu : (unsigned int)acc_out;
      const unsigned int a = __reduce_add_sync(0xffffffffu, vi);
      const unsigned int b = __reduce_add_sync(0xffffffffu, vo);
      if (lane == 0) {
        if (a) atomicAdd(&s_cnt[t], a);
        if (b) atomicAdd(&s_cnt[6 + t], b);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L * 12; i += C::THREADS)
    if (s_cnt[i]) atomicAdd(&bg_counts[i], (unsigned long long)s_cnt[i]);
}

// ---- compaction of the cut mask -------------------------------------------------------------------
template <int CPC>
__global__ void k_cutmask_count(const uint32_t* __restrict__ cutmask, int64_t ngroups, uint32_t* __restrict__ counts) {
  int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= ngroups) return;
  uint32_t c = 0;
#pragma unroll
  for (int t = 0; t < CPC; ++t) c += __popc(cutmask[gidx * CPC + t]);
  counts[gidx] = c;
}

template <int CPC>
__global__ void k_cutmask_emit(const uint32_t* __restrict__ cutmask, int64_t ngroups, const uint32_t* __restrict__ offs,
                               int nx, int wxtot, int32_t* __restrict__ cutcells) {
  int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= ngroups) return;
  uint32_t m[CPC];
  uint32_t any = 0;
#pragma unroll
  for (int t = 0; t < CPC; ++t) {
    m[t] = cutmask[gidx * CPC + t];
    any |= m[t];
  }
  if (!any) return;
  int64_t rowid = gidx / wxtot;
  int w = (int)(gidx % wxtot);
  int64_t cube0 = rowid * nx + 32 * (int64_t)w;
  uint32_t o = offs[gidx];
  while (any) {
    int c = __ffs(any) - 1;
    any &= any - 1;
#pragma unroll
    for (int t = 0; t < CPC; ++t)
      if ((m[t] >> c) & 1u) cutcells[o++] = (int32_t)((cube0 + c) * CPC + t + 1);
  }
}

template <int D>
__global__ void k_eval_leaf_nodes(const GcGrid g, const gecut_leaf leaf, double* __restrict__ out) {
  int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)g.nn[0] * g.nn[1] * (D == 3 ? g.nn[2] : 1);
  if (node >= total) return;
  int i = (int)(node % g.nn[0]);
  int j = (int)((node / g.nn[0]) % g.nn[1]);
  int k = (D == 3) ? (int)(node / ((int64_t)g.nn[0] * g.nn[1])) : 0;
  double x[3] = {gc_coord(g, 0, i), gc_coord(g, 1, j), D == 3 ? gc_coord(g, 2, k) : 0.0};
  out[node] = gc_eval_leaf(leaf, x, D);
}

}  // namespace

int gc_classify(gecut_ctx* ctx, gecut_result* res) {
  const GcGrid& g = res->grid;
  const int L = res->L;
  const int nx = g.n[0];
  const int wxtot = gc_words_per_row(nx);
  const int nlayers = g.k1 - g.k0;
  auto launch = [&](auto Dtag, auto Stag) {
    constexpr int D = decltype(Dtag)::value;
    constexpr bool S = decltype(Stag)::value;
    using C = ClsCfg<D>;
    dim3 grid;
    grid.x = (unsigned)gc_div_up(wxtot, C::WX);
    if (D == 3) {
      grid.y = (unsigned)gc_div_up(g.n[1], C::TY);
      grid.z = (unsigned)gc_div_up(nlayers, C::ZS);
    } else {
      grid.y = (unsigned)gc_div_up(nlayers, C::ZS);
      grid.z = 1;
    }
    size_t smem = sizeof(uint32_t) * 2 * L * C::RY * (C::WX + 1);
    if (L == 1)
      GC_LAUNCH(ctx, "k_classify", k_classify<D, S, 1><<<grid, C::THREADS, smem, ctx->stream>>>(
          g, res->leaves, res->lay.bg_state, res->lay.bg_pitch, res->cutmask, ctx->d_simplexify_raw, res->bg_counts_dev));
    else if (L == 2)
      GC_LAUNCH(ctx, "k_classify", k_classify<D, S, 2><<<grid, C::THREADS, smem, ctx->stream>>>(
          g, res->leaves, res->lay.bg_state, res->lay.bg_pitch, res->cutmask, ctx->d_simplexify_raw, res->bg_counts_dev));
    else
      GC_LAUNCH(ctx, "k_classify", k_classify<D, S, 0><<<grid, C::THREADS, smem, ctx->stream>>>(
          g, res->leaves, res->lay.bg_state, res->lay.bg_pitch, res->cutmask, ctx->d_simplexify_raw, res->bg_counts_dev));
  };
  if (g.D == 3) {
    if (g.simplexified)
      launch(std::integral_constant<int, 3>{}, std::true_type{});
    else
      launch(std::integral_constant<int, 3>{}, std::false_type{});
  } else {
    if (g.simplexified)
      launch(std::integral_constant<int, 2>{}, std::true_type{});
    else
      launch(std::integral_constant<int, 2>{}, std::false_type{});
  }
  return gc_launch_check(ctx, "k_classify");
}

int gc_compact_cutcells(gecut_ctx* ctx, gecut_result* res) {
  const GcGrid& g = res->grid;
  const int nx = g.n[0];
  const int wxtot = gc_words_per_row(nx);
  const int64_t rows = (g.D == 3) ? (int64_t)(g.k1 - g.k0) * g.n[1] : (int64_t)(g.k1 - g.k0);
  const int64_t ngroups = rows * wxtot;
  uint32_t* counts = nullptr;
  uint64_t* total_dev = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&counts, sizeof(uint32_t) * ngroups));
  GC_CHECK(gc_malloc(ctx, (void**)&total_dev, sizeof(uint64_t)));
  const int T = 256;
  const unsigned nb = (unsigned)gc_div_up(ngroups, T);
  switch (g.cpc) {
    case 1: GC_LAUNCH(ctx, "k_cutmask_count", k_cutmask_count<1><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts)); break;
    case 2: GC_LAUNCH(ctx, "k_cutmask_count", k_cutmask_count<2><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts)); break;
    default: GC_LAUNCH(ctx, "k_cutmask_count", k_cutmask_count<6><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts)); break;
  }
  GC_CHECK(gc_exclusive_scan_u32(ctx, counts, counts, ngroups, total_dev));
  uint64_t* h = (uint64_t*)ctx->h_pinned;
  GC_CUDA(cudaMemcpyAsync(h, total_dev, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaMemcpyAsync(h + 1, res->bg_counts_dev, sizeof(uint64_t) * res->L * 12, cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t ncut = (int64_t)h[0];
  for (int i = 0; i < res->L * 12; ++i) res->bg_counts[i] = (int64_t)h[1 + i];
  res->sizes.num_cutcells = ncut;
  if (ncut > 0) {
    GC_CHECK(gc_malloc(ctx, (void**)&res->lay.cutcells, sizeof(int32_t) * ncut));
    switch (g.cpc) {
      case 1: GC_LAUNCH(ctx, "k_cutmask_emit", k_cutmask_emit<1><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts, nx, wxtot, res->lay.cutcells)); break;
      case 2: GC_LAUNCH(ctx, "k_cutmask_emit", k_cutmask_emit<2><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts, nx, wxtot, res->lay.cutcells)); break;
      default: GC_LAUNCH(ctx, "k_cutmask_emit", k_cutmask_emit<6><<<nb, T, 0, ctx->stream>>>(res->cutmask, ngroups, counts, nx, wxtot, res->lay.cutcells)); break;
    }
  }
  int st = gc_launch_check(ctx, "cutmask compaction");
  gc_free(ctx, counts);
  gc_free(ctx, total_dev);
  return st;
}

int gc_eval_leaf_nodes(gecut_ctx* ctx, const GcGrid& g, const gecut_leaf& leaf, double* out_dev) {
  int64_t total = (int64_t)g.nn[0] * g.nn[1] * (g.D == 3 ? g.nn[2] : 1);
  const int T = 256;
  unsigned nb = (unsigned)gc_div_up(total, T);
  if (g.D == 3)
    GC_LAUNCH(ctx, "k_eval_leaf_nodes", k_eval_leaf_nodes<3><<<nb, T, 0, ctx->stream>>>(g, leaf, out_dev));
  else
    GC_LAUNCH(ctx, "k_eval_leaf_nodes", k_eval_leaf_nodes<2><<<nb, T, 0, ctx->stream>>>(g, leaf, out_dev));
  return gc_launch_check(ctx, "k_eval_leaf_nodes");
}
