// gecut_select.cu -- selectors: boolean programs over the IN/OUT/CUT rows, masks and stable compaction.
//
// Reference path replaced (src/Interfaces/EmbeddedDiscretizations.jl):
//   compute_bgcell_to_inoutcut / compute_inoutcut + truth tables   :91-177
//   compute_subcell_to_inout                                       :325-339
//   Triangulation(cut, CutInOrOut / Integer / ActiveInOrOut, geo)  :297-323   (mask + findall)
//   compute_inoutboundary + orientation rules, EmbeddedBoundary    :179-245, 408-463
//   SubCellData(st,ids) / SubFacetData(st,ids,orientation)         SubCellTriangulations.jl:9-17, SubFacetTriangulations.jl:10-21
//
// The reference evaluates the CSG tree by broadcasting over whole state vectors (one temporary vector per tree node);
// here the tree is a postfix program evaluated per element in registers, fused with the mask, and `findall` is a
// flag -> exclusive scan -> scatter (stable, so ids come out ascending like the reference).
#include <algorithm>

#include "gecut_internal.cuh"

namespace {

__global__ void k_eval_rows(const int8_t* __restrict__ rows, int64_t pitch, int64_t n, const GcProgram prog,
                            int8_t* __restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  out[e] = (int8_t)gc_eval_inoutcut(prog, rows, pitch, e);
}

// truth-table programs, 16 elements per thread: one 128-bit load per used state row, one 128-bit store (rows and `out`
// are padded to 16 bytes)
__global__ void __launch_bounds__(256)
k_eval_rows_lut16(const int8_t* __restrict__ rows, int64_t pitch, int64_t n, const GcProgram prog, int8_t* __restrict__ out) {
  const int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (e0 >= n) return;
  uint32_t idx[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) idx[c] = 0u;
  for (int k = prog.nused - 1; k >= 0; --k) {
    const uint4 r4 = *reinterpret_cast<const uint4*>(rows + (int64_t)prog.used[k] * pitch + e0);
    const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
    for (int c = 0; c < 16; ++c) idx[c] = idx[c] * 3u + (uint32_t)((int)(int8_t)((w4[c >> 2] >> (8 * (c & 3))) & 0xFFu) + 1);
  }
  uint32_t o4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const uint32_t v = e0 + c < n ? (uint32_t)(uint8_t)__ldg(prog.lut + idx[c]) : 0u;
    o4[c >> 2] |= v << (8 * (c & 3));
  }
  *reinterpret_cast<uint4*>(out + e0) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
}

// Triangulation(cut,CutInOrOut(side),geo): bg state == CUT && subcell state == side, as a stream compaction over the
// subcells: 16 consecutive subcells per thread (one 128-bit load of the state row for single-leaf programs), 4096 per
// block.  check_bg == 0 (one level set: every cut cell is CUT for every program over it) skips the bg-cell gather.
// EMIT == false: number of selected subcells per block; EMIT == true: their 1-based ids at the scanned block offsets.
constexpr int SELSC_T = 256, SELSC_PER = 16;

template <bool EMIT>
__global__ void __launch_bounds__(SELSC_T)
k_select_subcells(const GcLayout lay, int64_t nsc, const GcProgram prog, int side, int check_bg,
                  uint32_t* __restrict__ counts_or_offs, int32_t* __restrict__ ids) {
  __shared__ uint32_t s_w[SELSC_T / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t e0 = ((int64_t)blockIdx.x * SELSC_T + threadIdx.x) * SELSC_PER;
  uint32_t bits = 0u;  // bit i: subcell e0+i is selected
  if (e0 < nsc) {
    const bool leaf1 = prog.len == 1 || (prog.len == 2 && prog.tok[1] == GECUT_OP_COMPLEMENT);
    if (leaf1 || prog.sc_row) {
      const int want = (!prog.sc_row && prog.len == 2) ? -side : side;  // complement swaps IN and OUT
      // rows are padded to a multiple of 256 bytes: the vector load may run past nsc (memoised rows: allocation granularity)
      const int8_t* row = prog.sc_row ? prog.sc_row : lay.sc_state + (int64_t)prog.tok[0] * lay.sc_pitch;
      const uint4 q = *reinterpret_cast<const uint4*>(row + e0);
      const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
      const uint32_t pat = (uint32_t)(uint8_t)(int8_t)want * 0x01010101u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t x = w4[j] ^ pat;  // zero byte <=> state == want
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (((x >> (8 * i)) & 0xFFu) == 0u) bits |= 1u << (4 * j + i);
      }
    } else if (prog.lut) {
      // truth table: one 128-bit load per used state row (rows are padded), base-3 index per subcell, one gather; runs of
      // 16 subcells with one state per row need a single gather
      bool uni = true;
      uint32_t idx0 = 0u;
      for (int k = prog.nused - 1; k >= 0; --k) {
        const uint4 q = *reinterpret_cast<const uint4*>(lay.sc_state + (int64_t)prog.used[k] * lay.sc_pitch + e0);
        uni = uni && q.x == q.y && q.y == q.z && q.z == q.w && q.x == (q.x & 0xFFu) * 0x01010101u;
        idx0 = idx0 * 3u + (uint32_t)((int)(int8_t)(q.x & 0xFFu) + 1);
      }
      if (uni) {
        bits = ((int)__ldg(prog.lut + idx0) == side) ? 0xFFFFu : 0u;
      } else {
        uint32_t idx[SELSC_PER];
#pragma unroll
        for (int c = 0; c < SELSC_PER; ++c) idx[c] = 0u;
        for (int k = prog.nused - 1; k >= 0; --k) {
          const uint4 q = *reinterpret_cast<const uint4*>(lay.sc_state + (int64_t)prog.used[k] * lay.sc_pitch + e0);
          const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int c = 0; c < SELSC_PER; ++c)
            idx[c] = idx[c] * 3u + (uint32_t)((int)(int8_t)((w4[c >> 2] >> (8 * (c & 3))) & 0xFFu) + 1);
        }
        const int nval = (int)(nsc - e0 < SELSC_PER ? nsc - e0 : SELSC_PER);  // bytes past nsc are padding: no gather
#pragma unroll
        for (int c = 0; c < SELSC_PER; ++c)
          if (c < nval && (int)__ldg(prog.lut + idx[c]) == side) bits |= 1u << c;
      }
    } else {
#pragma unroll 4
      for (int i = 0; i < SELSC_PER; ++i)
        if (e0 + i < nsc && gc_eval_inoutcut(prog, lay.sc_state, lay.sc_pitch, e0 + i) == side) bits |= 1u << i;
    }
    const int64_t left = nsc - e0;
    if (left < SELSC_PER) bits &= (1u << (int)left) - 1u;
    if (check_bg) {
      // consecutive subcells mostly share their bg cell: the state of the last one looked up is kept
      int32_t last_bg = 0;
      bool last_ok = false;
      for (uint32_t m = bits; m; m &= m - 1) {
        const int i = __ffs(m) - 1;
        const int32_t bgc = lay.sc_c2bg[e0 + i];
        if (bgc != last_bg) {
          last_bg = bgc;
          last_ok = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, (int64_t)bgc - 1) == GC_CUT;
        }
        if (!last_ok) bits &= ~(1u << i);
      }
    }
  }
  const uint32_t cnt = __popc(bits);
  // block-level exclusive prefix of the per-thread counts (thread order = ascending subcell ids)
  uint32_t inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  uint32_t wbase = 0u, total = 0u;
#pragma unroll
  for (int k = 0; k < SELSC_T / 32; ++k) {
    const uint32_t v = s_w[k];
    if (k < wid) wbase += v;
    total += v;
  }
  if (!EMIT) {
    if (threadIdx.x == 0) counts_or_offs[blockIdx.x] = total;
  } else {
    uint32_t o = counts_or_offs[blockIdx.x] + wbase + inc - cnt;
    for (uint32_t m = bits; m; m &= m - 1) ids[o++] = (int32_t)(e0 + (__ffs(m) - 1) + 1);
  }
}

// bg cells: count per cell type of the cells with state == side (mode 0) or CUT||side (mode 1); optionally flags
__global__ void k_count_bgcells(const GcLayout lay, int64_t nc, const GcProgram prog, int side, int mode, int cpc,
                                unsigned long long* __restrict__ counts, uint32_t* __restrict__ flags) {
  __shared__ unsigned int s_cnt[8];
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nc; e += (int64_t)gridDim.x * blockDim.x) {
    int bg = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, e);
    bool keep = mode == 0 ? (bg == side) : (bg == GC_CUT || bg == side);
    if (flags) flags[e] = keep ? 1u : 0u;
    if (keep) atomicAdd(&s_cnt[(int)(e % cpc)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// count-only variant: 16 cells per thread with 128-bit loads of every state row (HBM-bound: L bytes per cell)
__global__ void __launch_bounds__(256)
k_count_bgcells_vec(const GcLayout lay, int64_t nc, int L, const GcProgram prog, int side, int mode, int cpc,
                    unsigned long long* __restrict__ counts) {
  __shared__ unsigned int s_cnt[8];
  if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  unsigned int loc[6] = {0, 0, 0, 0, 0, 0};
  // per group of 16 cells the per-type tallies are packed one byte per cell type (<= 16 each) and unpacked once
  auto unpack = [&](unsigned long long pk) {
#pragma unroll
    for (int t = 0; t < 6; ++t) loc[t] += (unsigned int)(pk >> (8 * t)) & 0xFFu;
  };
  const int64_t ngroups = (nc + 15) / 16;
  for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = gi * 16;
    const int nvalid = (int)(nc - e0 < 16 ? nc - e0 : 16);
    int ty = (int)(e0 % cpc);
    if (prog.len == 1 || prog.bg_row) {
      const int8_t* row = prog.bg_row ? prog.bg_row : lay.bg_state + (int64_t)prog.tok[0] * lay.bg_pitch;
      const uint4 r4 = *reinterpret_cast<const uint4*>(row + e0);
      const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
      unsigned long long pk = 0ull;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int bg = (int)(int8_t)((w4[c >> 2] >> (8 * (c & 3))) & 0xFFu);
        const bool keep = (c < nvalid) && (mode == 0 ? (bg == side) : (bg == GC_CUT || bg == side));
        pk += (unsigned long long)(keep ? 1u : 0u) << (8 * ty);
        ty = (ty + 1 == cpc) ? 0 : ty + 1;
      }
      unpack(pk);
    } else if (prog.lut) {
      // groups whose 16 cells have the same state in every used row (all but the ones a surface crosses): one gather
      {
        bool uni = nvalid == 16;
        uint32_t idx0 = 0u;
        for (int k = prog.nused - 1; k >= 0; --k) {
          const uint4 r4 = *reinterpret_cast<const uint4*>(lay.bg_state + (int64_t)prog.used[k] * lay.bg_pitch + e0);
          uni = uni && r4.x == r4.y && r4.y == r4.z && r4.z == r4.w && r4.x == (r4.x & 0xFFu) * 0x01010101u;
          idx0 = idx0 * 3u + (uint32_t)((int)(int8_t)(r4.x & 0xFFu) + 1);
        }
        if (uni) {
          const int bg = (int)__ldg(prog.lut + idx0);
          if (mode == 0 ? (bg == side) : (bg == GC_CUT || bg == side)) {
            if (cpc == 1) {
              loc[0] += 16u;
            } else {
              const int base = 16 / cpc, rem = 16 % cpc;
#pragma unroll
              for (int t = 0; t < 6; ++t) {
                int dlt = t - ty;
                if (dlt < 0) dlt += cpc;
                if (t < cpc) loc[t] += (unsigned int)(base + (dlt < rem ? 1 : 0));
              }
            }
          }
          continue;
        }
      }
      // truth table: one 128-bit load per used level-set row, base-3 index per cell, one gather
      uint32_t idx[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) idx[c] = 0u;
      for (int k = prog.nused - 1; k >= 0; --k) {
        const uint4 r4 = *reinterpret_cast<const uint4*>(lay.bg_state + (int64_t)prog.used[k] * lay.bg_pitch + e0);
        const uint32_t w4[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int c = 0; c < 16; ++c)
          idx[c] = idx[c] * 3u + (uint32_t)((int)(int8_t)((w4[c >> 2] >> (8 * (c & 3))) & 0xFFu) + 1);
      }
      unsigned long long pk = 0ull;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int bg = c < nvalid ? (int)__ldg(prog.lut + idx[c]) : GC_OUT;
        const bool keep = (c < nvalid) && (mode == 0 ? (bg == side) : (bg == GC_CUT || bg == side));
        pk += (unsigned long long)(keep ? 1u : 0u) << (8 * ty);
        ty = (ty + 1 == cpc) ? 0 : ty + 1;
      }
      unpack(pk);
    } else {
      for (int c = 0; c < nvalid; ++c) {
        int8_t vals[GC_LMAX];
        for (int ls = 0; ls < L; ++ls) vals[ls] = lay.bg_state[(int64_t)ls * lay.bg_pitch + e0 + c];
        const int bg = gc_eval_inoutcut_vals(prog, vals);
        const bool keep = mode == 0 ? (bg == side) : (bg == GC_CUT || bg == side);
#pragma unroll
        for (int t = 0; t < 6; ++t) loc[t] += (keep && t == ty) ? 1u : 0u;
        ty = (ty + 1 == cpc) ? 0 : ty + 1;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    unsigned int v = loc[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[t], v);
  }
  __syncthreads();
  if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void k_flag_boundary(const GcLayout lay, int64_t nsf, const GcProgram prog1, const GcProgram prog2,
                                int has2, uint32_t* __restrict__ flags, int8_t* __restrict__ orient_all) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nsf) return;
  int s1, o1;
  if (prog1.blut) {  // truth table: one gather instead of the stack machine (whose stacks live in local memory)
    int idx = 0;
    for (int k = prog1.nused - 1; k >= 0; --k) idx = idx * 3 + (lay.sf_state[(int64_t)prog1.used[k] * lay.sf_pitch + e] + 1);
    const int v = __ldg(prog1.blut + idx);
    s1 = (v & 3) - 1;
    o1 = (v >> 2) - 1;
  } else {
    gc_eval_inoutboundary(prog1, lay.sf_state, lay.sf_pitch, e, &s1, &o1);
  }
  bool keep = s1 == GECUT_INTERFACE;
  if (keep && has2) keep = gc_eval_inoutcut(prog2, lay.sf_state, lay.sf_pitch, e) == GECUT_INTERFACE;
  flags[e] = keep ? 1u : 0u;
  orient_all[e] = (int8_t)o1;
}

// scatter of the flagged elements: ids[offs[e]] = e+1 (1-based), optional orientation
__global__ void k_scatter_ids(const uint32_t* __restrict__ flags_scanned, const uint32_t* __restrict__ flags_raw_next,
                              int64_t n, uint64_t total, int32_t* __restrict__ ids, const int8_t* __restrict__ orient_all,
                              int8_t* __restrict__ orient) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  uint32_t o = flags_scanned[e];
  uint32_t nxt = (e + 1 < n) ? flags_scanned[e + 1] : (uint32_t)total;
  if (nxt != o) {
    ids[o] = (int32_t)(e + 1);
    if (orient) orient[o] = orient_all[e];
  }
  (void)flags_raw_next;
}

template <int STRIDE>
__global__ void k_gather_rows(const int32_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ c2p,
                              const int32_t* __restrict__ c2bg, const double* __restrict__ normals,
                              const int8_t* __restrict__ orient, int D, int32_t* __restrict__ out_c2p,
                              int32_t* __restrict__ out_c2bg, double* __restrict__ out_normals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t e = (int64_t)ids[i] - 1;
  if (out_c2p)
#pragma unroll
    for (int k = 0; k < STRIDE; ++k) out_c2p[i * STRIDE + k] = c2p[e * STRIDE + k];
  if (out_c2bg) out_c2bg[i] = c2bg[e];
  if (out_normals) {
    double o = (double)orient[i];
    for (int d = 0; d < D; ++d) out_normals[i * D + d] = __dmul_rn(normals[e * D + d], o);
  }
}

}  // namespace

int gc_eval_rows(gecut_ctx* ctx, const int8_t* rows, int64_t pitch, int64_t n, const GcProgram& prog, int8_t* out_dev,
                 bool out_padded) {
  if (n <= 0) return GECUT_OK;
  const int T = 256;
  if (prog.lut && out_padded && (reinterpret_cast<uintptr_t>(out_dev) & 15) == 0)
    GC_LAUNCH(ctx, "k_eval_rows", k_eval_rows_lut16<<<(unsigned)gc_div_up(gc_div_up(n, 16), T), T, 0, ctx->stream>>>(rows, pitch, n, prog, out_dev));
  else
    GC_LAUNCH(ctx, "k_eval_rows", k_eval_rows<<<(unsigned)gc_div_up(n, T), T, 0, ctx->stream>>>(rows, pitch, n, prog, out_dev));
  return gc_launch_check(ctx, "k_eval_rows");
}

// flags (uint32, n) -> ids (1-based, ascending); *count out.  flags buffer is consumed (scanned in place).
int gc_compact_flags(gecut_ctx* ctx, uint32_t* flags, int64_t n, int32_t** ids_out, int64_t* count,
                     const int8_t* orient_all, int8_t** orient_out) {
  *ids_out = nullptr;
  if (orient_out) *orient_out = nullptr;
  *count = 0;
  if (n <= 0) return GECUT_OK;
  GcScratch scratch(ctx);  // released on every exit path
  uint64_t* total_dev = nullptr;
  GC_CHECK(scratch.alloc(&total_dev, sizeof(uint64_t)));
  GC_CHECK(gc_exclusive_scan_u32(ctx, flags, flags, n, total_dev));
  uint64_t* h = (uint64_t*)ctx->h_pinned;
  GC_CUDA(cudaMemcpyAsync(h, total_dev, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  uint64_t total = h[0];
  *count = (int64_t)total;
  if (total == 0) return GECUT_OK;
  GC_CHECK(gc_malloc(ctx, (void**)ids_out, sizeof(int32_t) * total));
  if (orient_out) GC_CHECK(gc_malloc(ctx, (void**)orient_out, sizeof(int8_t) * total));
  const int T = 256;
  GC_LAUNCH(ctx, "k_scatter_ids", k_scatter_ids<<<(unsigned)gc_div_up(n, T), T, 0, ctx->stream>>>(flags, nullptr, n, total, *ids_out, orient_all,
                                                                  orient_out ? *orient_out : nullptr));
  return gc_launch_check(ctx, "k_scatter_ids");
}

// count -> scan -> emit of the selected subcells (ascending ids = findall)
static int select_subcells_now(gecut_ctx* ctx, gecut_selection* sel) {
  const gecut_result* res = sel->res;
  const GcLayout& lay = res->lay;
  const int64_t nsc = res->sizes.num_subcells;
  {
    const int64_t nblk = gc_div_up(nsc, (int64_t)SELSC_T * SELSC_PER);
    const int check_bg = res->L > 1 ? 1 : 0;  // one level set: every cut cell is CUT for every program over it
    GcScratch scratch(ctx);
    uint32_t* counts = nullptr;
    uint64_t* total_dev = nullptr;
    GC_CHECK(scratch.alloc(&counts, sizeof(uint32_t) * nblk));
    GC_CHECK(scratch.alloc(&total_dev, sizeof(uint64_t)));
    const unsigned nbk = (unsigned)nblk;
    GC_LAUNCH(ctx, "k_select_subcells", k_select_subcells<false><<<nbk, SELSC_T, 0, ctx->stream>>>(
        lay, nsc, sel->prog, sel->side, check_bg, counts, nullptr));
    GC_CHECK(gc_exclusive_scan_u32(ctx, counts, counts, nblk, total_dev));
    uint64_t* h = (uint64_t*)ctx->h_pinned;
    GC_CUDA(cudaMemcpyAsync(h, total_dev, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    sel->n_sub = (int64_t)h[0];
    if (sel->n_sub > 0) {
      GC_CHECK(gc_malloc(ctx, (void**)&sel->sub_ids, sizeof(int32_t) * sel->n_sub));
      GC_LAUNCH(ctx, "k_select_subcells", k_select_subcells<true><<<nbk, SELSC_T, 0, ctx->stream>>>(
          lay, nsc, sel->prog, sel->side, check_bg, counts, sel->sub_ids));
    }
  }
  return gc_launch_check(ctx, "select_subcells");
}

int gc_select_cells(gecut_ctx* ctx, gecut_selection* sel) {
  const gecut_result* res = sel->res;
  const GcLayout& lay = res->lay;
  const int64_t nsc = res->sizes.num_subcells, nc = res->sizes.num_bgcells;
  const int T = 256;
  const bool want_sub = sel->kind == GECUT_SEL_PHYSICAL || sel->kind == GECUT_SEL_CUTONLY;
  const bool want_bg = sel->kind == GECUT_SEL_PHYSICAL || sel->kind == GECUT_SEL_ACTIVE || sel->kind == GECUT_SEL_BGONLY;
  if (want_sub && nsc > 0) {
    const GcProgram& p = sel->prog;
    const bool leaf1 = (p.len == 1 && p.tok[0] >= 0) || (p.len == 2 && p.tok[0] >= 0 && p.tok[1] == GECUT_OP_COMPLEMENT);
    if (res->has_l1_stats && leaf1) {
      // one level set: count (and measure) were summed by the fill kernel; ids on demand (gc_materialize_sub_ids)
      sel->sub_lazy = true;
      sel->lazy_side = p.len == 1 ? sel->side : -sel->side;
      const int64_t n_in = (int64_t)res->l1_stats[3];
      sel->n_sub = sel->lazy_side == GECUT_IN ? n_in : nsc - n_in;
    } else {
      GC_CHECK(select_subcells_now(ctx, sel));
    }
  }
  // single-leaf programs ([ls] or [ls, !]): the classification kernel already tallied IN / OUT cells per level set and
  // cell type, no pass over the Nc state bytes is needed
  bool from_tally = false;
  if (want_bg && nc > 0) {
    const GcProgram& p = sel->prog;
    const bool leaf = p.len == 1 && p.tok[0] >= 0;
    const bool notleaf = p.len == 2 && p.tok[0] >= 0 && p.tok[1] == GECUT_OP_COMPLEMENT;
    if (leaf || notleaf) {
      const int ls = p.tok[0];
      const int side_eff = notleaf ? -sel->side : sel->side;
      const int64_t* cin = res->bg_counts + ls * 12;
      const int64_t* cout = cin + 6;
      sel->n_bg = 0;
      for (int t = 0; t < res->grid.cpc; ++t) {
        int64_t c;
        if (sel->kind == GECUT_SEL_ACTIVE)  // CUT or side  =  all cells of the type minus the other side
          c = res->grid.num_cubes - (side_eff == GECUT_IN ? cout[t] : cin[t]);
        else
          c = side_eff == GECUT_IN ? cin[t] : cout[t];
        sel->bg_type_count[t] = c;
        sel->n_bg += c;
      }
      from_tally = true;
    }
  }
  if (want_bg && nc > 0 && !from_tally) {
    GcScratch scratch(ctx);
    unsigned long long* counts = nullptr;
    GC_CHECK(scratch.alloc(&counts, sizeof(unsigned long long) * 8));
    GC_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 8, ctx->stream));
    const int mode = sel->kind == GECUT_SEL_ACTIVE ? 1 : 0;
    unsigned nb = (unsigned)std::min<int64_t>(gc_div_up(gc_div_up(nc, 16), T), 148 * 64);
    GC_LAUNCH(ctx, "k_count_bgcells", k_count_bgcells_vec<<<nb, T, 0, ctx->stream>>>(lay, nc, res->L, sel->prog, sel->side, mode,
                                                                                   res->grid.cpc, counts));
    uint64_t* h = (uint64_t*)ctx->h_pinned;
    GC_CUDA(cudaMemcpyAsync(h, counts, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    sel->n_bg = 0;
    for (int t = 0; t < 8; ++t) {
      sel->bg_type_count[t] = (int64_t)h[t];
      sel->n_bg += (int64_t)h[t];
    }
  }
  return gc_launch_check(ctx, "select_cells");
}

int gc_materialize_bg_ids(gecut_ctx* ctx, gecut_selection* sel) {
  if (sel->bg_ids || sel->n_bg == 0) return GECUT_OK;
  const gecut_result* res = sel->res;
  const int64_t nc = res->sizes.num_bgcells;
  const int T = 256;
  GcScratch scratch(ctx);
  uint32_t* flags = nullptr;
  unsigned long long* counts = nullptr;
  GC_CHECK(scratch.alloc(&flags, sizeof(uint32_t) * nc));
  GC_CHECK(scratch.alloc(&counts, sizeof(unsigned long long) * 8));
  GC_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 8, ctx->stream));
  const int mode = sel->kind == GECUT_SEL_ACTIVE ? 1 : 0;
  unsigned nb = (unsigned)std::min<int64_t>(gc_div_up(nc, T), 148 * 16);
  GC_LAUNCH(ctx, "k_count_bgcells", k_count_bgcells<<<nb, T, 0, ctx->stream>>>(res->lay, nc, sel->prog, sel->side, mode, res->grid.cpc, counts, flags));
  int64_t n = 0;
  return gc_compact_flags(ctx, flags, nc, &sel->bg_ids, &n, nullptr, nullptr);
}

__global__ void k_iota_boundary(int64_t n, int32_t* __restrict__ ids, int8_t* __restrict__ orient) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  ids[e] = (int32_t)(e + 1);
  orient[e] = 1;
}

int gc_select_boundary(gecut_ctx* ctx, gecut_selection* sel) {
  const gecut_result* res = sel->res;
  const int64_t nsf = res->sizes.num_subfacets;
  if (nsf <= 0) return GECUT_OK;
  const int T = 256;
  // one level set, program = that leaf (and no second geometry, or the same leaf): every subfacet is INTERFACE with
  // orientation +1 (compute_inoutboundary of a Leaf, EmbeddedDiscretizations.jl:179-182) -> the selection is 1..Nsf
  if (res->L == 1 && sel->prog.len == 1 && (!sel->has_prog2 || sel->prog2.len == 1)) {
    sel->n_sub = nsf;
    sel->sub_identity = true;  // ids / orientation are written when somebody asks for them (gc_materialize_sub_ids)
    return GECUT_OK;
  }
  GcScratch scratch(ctx);
  uint32_t* flags = nullptr;
  int8_t* orient_all = nullptr;
  GC_CHECK(scratch.alloc(&flags, sizeof(uint32_t) * nsf));
  GC_CHECK(scratch.alloc(&orient_all, sizeof(int8_t) * nsf));
  GC_LAUNCH(ctx, "k_flag_boundary", k_flag_boundary<<<(unsigned)gc_div_up(nsf, T), T, 0, ctx->stream>>>(res->lay, nsf, sel->prog, sel->prog2,
                                                                      sel->has_prog2 ? 1 : 0, flags, orient_all));
  return gc_compact_flags(ctx, flags, nsf, &sel->sub_ids, &sel->n_sub, orient_all, &sel->orientation);
}

int gc_materialize_sub_ids(gecut_ctx* ctx, gecut_selection* sel) {
  if (sel->sub_lazy && !sel->sub_ids && sel->n_sub > 0) {
    const int64_t expect = sel->n_sub;
    GC_CHECK(select_subcells_now(ctx, sel));
    if (sel->n_sub != expect) {
      ctx->err = "selection: id list and tile statistics disagree";
      return GECUT_ERR_CUDA;
    }
    return GECUT_OK;
  }
  if (!sel->sub_identity || sel->sub_ids || sel->n_sub <= 0) return GECUT_OK;
  const int64_t n = sel->n_sub;
  const int T = 256;
  GC_CHECK(gc_malloc(ctx, (void**)&sel->sub_ids, sizeof(int32_t) * n));
  GC_CHECK(gc_malloc(ctx, (void**)&sel->orientation, sizeof(int8_t) * n));
  GC_LAUNCH(ctx, "k_iota_boundary", k_iota_boundary<<<(unsigned)gc_div_up(n, T), T, 0, ctx->stream>>>(n, sel->sub_ids, sel->orientation));
  return gc_launch_check(ctx, "k_iota_boundary");
}

int gc_selection_gather(gecut_ctx* ctx, const gecut_selection* sel, int32_t* to_points_dev, int32_t* to_bgcell_dev,
                        double* normals_dev) {
  const gecut_result* res = sel->res;
  const int64_t n = sel->n_sub;
  if (n <= 0) return GECUT_OK;
  GC_CHECK(gc_materialize_sub_ids(ctx, const_cast<gecut_selection*>(sel)));
  const int D = res->grid.D;
  const int T = 256;
  const unsigned nb = (unsigned)gc_div_up(n, T);
  const GcLayout& lay = res->lay;
  if (sel->kind == GECUT_SEL_BOUNDARY) {
    if (D == 2)
      GC_LAUNCH(ctx, "k_gather_rows", k_gather_rows<2><<<nb, T, 0, ctx->stream>>>(sel->sub_ids, n, lay.sf_c2p, lay.sf_c2bg, lay.sf_N, sel->orientation, D,
                                                  to_points_dev, to_bgcell_dev, normals_dev));
    else
      GC_LAUNCH(ctx, "k_gather_rows", k_gather_rows<3><<<nb, T, 0, ctx->stream>>>(sel->sub_ids, n, lay.sf_c2p, lay.sf_c2bg, lay.sf_N, sel->orientation, D,
                                                  to_points_dev, to_bgcell_dev, normals_dev));
  } else {
    if (D == 2)
      GC_LAUNCH(ctx, "k_gather_rows", k_gather_rows<3><<<nb, T, 0, ctx->stream>>>(sel->sub_ids, n, lay.sc_c2p, lay.sc_c2bg, nullptr, nullptr, D,
                                                  to_points_dev, to_bgcell_dev, nullptr));
    else
      GC_LAUNCH(ctx, "k_gather_rows", k_gather_rows<4><<<nb, T, 0, ctx->stream>>>(sel->sub_ids, n, lay.sc_c2p, lay.sc_c2bg, nullptr, nullptr, D,
                                                  to_points_dev, to_bgcell_dev, nullptr));
  }
  return gc_launch_check(ctx, "k_gather_rows");
}
