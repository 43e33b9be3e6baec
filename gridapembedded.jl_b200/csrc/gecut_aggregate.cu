// gecut_aggregate.cu -- AgFEM cell aggregation and the GhostSkeleton facet mask on the device.
//
// Reference path replaced (src/AgFEM/CellAggregation.jl):
//   aggregate(::AggregateCutCellsByThreshold, cut, [cut_facets,] geo, in_or_out)   :72-102
//   _aggregate_by_threshold                                                        :104-131
//   _aggregate_by_threshold_barrier / _find_best_neighbor / _touch_aggregated_cells! :133-241
//
// B200 design.  The reference sweeps all cells up to 20 times with lazy-array caches for the topology.  Inside one sweep
// a cell only reads the `touched` flags and the roots of TOUCHED neighbours, and those are frozen until the end of the
// sweep (touched cells are never revisited), so a sweep is a Jacobi step: one thread per cut cell, in place.  The
// topology of the Cartesian model is closed form (gecut_facets.cuh): no cell_to_faces / face_to_cells tables exist.
// Only the O(n^2) cut cells are visited after the O(n^3) initialisation pass; the unit cut measure of a cut cell is the
// sum of its selected subcells in subcell order (move_contributions), divided by the cell measure.
#include <cstdio>
#include <cstring>

#include "gecut_facets.cuh"

namespace {

// cell_to_cellin / touched for every cell: roots are the cells with unit cut measure >= threshold (:141-146); uncut
// cells have unit measure 1 (state == loc) or 0
__global__ void k_agg_init(const GcLayout lay, int64_t nc, const GcProgram prog, int loc, double threshold,
                           int32_t* __restrict__ cellin, uint8_t* __restrict__ touched) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int st = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, c);
  const double unit = st == loc ? 1.0 : 0.0;  // CUT cells are overwritten by k_agg_init_cut
  const bool root = st != GC_CUT && unit >= threshold;
  cellin[c] = root ? (int32_t)(c + 1) : 0;
  touched[c] = root ? 1 : 0;
}

__global__ void k_agg_init_cut(const GcLayout lay, const uint32_t* __restrict__ sc_ptr, int64_t ncut, const GcProgram prog,
                               int loc, double threshold, double cell_meas, int32_t* __restrict__ cellin,
                               uint8_t* __restrict__ touched) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncut) return;
  const int64_t c = (int64_t)lay.cutcells[k] - 1;
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, c) != GC_CUT) return;
  double V = 0.0;
  const uint32_t s0 = sc_ptr[2 * k], s1 = sc_ptr[2 * k + 2];
  for (uint32_t e = s0; e < s1; ++e)
    if (gc_eval_inoutcut(prog, lay.sc_state, lay.sc_pitch, (int64_t)e) == loc) V += lay.sc_meas[e];
  const bool root = V / cell_meas >= threshold;
  cellin[c] = root ? (int32_t)(c + 1) : 0;
  touched[c] = root ? 1 : 0;
}

// max over the vertex pairs of two cells of |p - q| (_find_best_neighbor :213-218), vertices in Gridap order
template <int D>
__device__ __forceinline__ double max_vertex_distance(const GcGrid& g, const int* a, const int* b) {
  double d = 0.0;
  for (int p = 0; p < (1 << D); ++p)
    for (int q = 0; q < (1 << D); ++q) {
      double s2 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double w = __dsub_rn(gc_coord(g, c, a[c] + ((p >> c) & 1)), gc_coord(g, c, b[c] + ((q >> c) & 1)));
        s2 = __dadd_rn(s2, __dmul_rn(w, w));
      }
      d = fmax(d, __dsqrt_rn(s2));
    }
  return d;
}

// one sweep (:155-176): untouched CUT cells look for the touched face neighbour whose root is closest
template <int D>
__global__ void k_agg_sweep(const GcGrid g, const GcLayout lay, int64_t ncut, const GcProgram prog, int loc,
                            const int8_t* __restrict__ facet_state, int32_t* __restrict__ cellin,
                            const uint8_t* __restrict__ touched, uint32_t* __restrict__ not_aggregated) {
  const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= ncut) return;
  const int64_t cell = (int64_t)lay.cutcells[kk] - 1;
  if (touched[cell]) return;
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell) != GC_CUT) return;
  int c[3];
  c[0] = (int)(cell % g.n[0]);
  c[1] = D == 3 ? (int)((cell / g.n[0]) % g.n[1]) : (int)(cell / g.n[0]);
  c[2] = D == 3 ? (int)(cell / ((int64_t)g.n[0] * g.n[1])) : 0;
  double dmin = INFINITY;
  int64_t best = -1;
  // faces in Gridap's local order: highest direction first, low then high
#pragma unroll
  for (int d = D - 1; d >= 0; --d) {
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int64_t f = facet_id_of<D>(g, c[0], c[1], c[2], d, side);
      const int io = facet_state[f];
      if (io != GC_CUT && io != loc) continue;
      int nb[3] = {c[0], c[1], c[2]};
      nb[d] += side == 0 ? -1 : 1;
      if (nb[d] < 0 || nb[d] >= g.n[d]) continue;  // boundary facet: the cell itself is the only one around
      const int64_t neigh = (int64_t)nb[0] + (int64_t)g.n[0] * ((int64_t)nb[1] + (int64_t)g.n[1] * nb[2]);
      if (!touched[neigh]) continue;
      const int64_t root = (int64_t)cellin[neigh] - 1;
      int r[3];
      r[0] = (int)(root % g.n[0]);
      r[1] = D == 3 ? (int)((root / g.n[0]) % g.n[1]) : (int)(root / g.n[0]);
      r[2] = D == 3 ? (int)(root / ((int64_t)g.n[0] * g.n[1])) : 0;
      const double dist = max_vertex_distance<D>(g, c, r);
      if (__dmul_rn(1.0 + 1.0e-9, dist) < dmin) {
        dmin = dist;
        best = neigh;
      }
    }
  }
  if (best >= 0)
    cellin[cell] = cellin[best];  // `best` is touched: its root is final, no race with other threads of the sweep
  else
    atomicAdd(not_aggregated, 1u);
}

// _touch_aggregated_cells! (:235-241) restricted to the cells that can change (cut cells)
__global__ void k_agg_touch(const GcLayout lay, int64_t ncut, const int32_t* __restrict__ cellin,
                            uint8_t* __restrict__ touched) {
  const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= ncut) return;
  const int64_t cell = (int64_t)lay.cutcells[kk] - 1;
  if (cellin[cell] > 0) touched[cell] = 1;
}

// GhostSkeleton mask (_fill_ghost_skeleton_mask!, src/Interfaces/EmbeddedDiscretizations.jl:520-543): interior facets
// with at least one CUT cell around and both cells active (CUT or `side`).  Every such facet touches a CUT cell, so
// the O(n^2) cut cells drive the pass; a facet between two CUT cells is written by both with the same value.
template <int D>
__global__ void k_ghost_mask(const GcGrid g, const GcLayout lay, int64_t ncut, const GcProgram prog, int side,
                             int8_t* __restrict__ mask, unsigned long long* __restrict__ count) {
  const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= ncut) return;
  const int64_t cell = (int64_t)lay.cutcells[kk] - 1;
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell) != GC_CUT) return;
  int c[3];
  c[0] = (int)(cell % g.n[0]);
  c[1] = D == 3 ? (int)((cell / g.n[0]) % g.n[1]) : (int)(cell / g.n[0]);
  c[2] = D == 3 ? (int)(cell / ((int64_t)g.n[0] * g.n[1])) : 0;
#pragma unroll
  for (int d = D - 1; d >= 0; --d) {
#pragma unroll
    for (int sd = 0; sd < 2; ++sd) {
      int nb[3] = {c[0], c[1], c[2]};
      nb[d] += sd == 0 ? -1 : 1;
      if (nb[d] < 0 || nb[d] >= g.n[d]) continue;
      const int64_t neigh = (int64_t)nb[0] + (int64_t)g.n[0] * ((int64_t)nb[1] + (int64_t)g.n[1] * nb[2]);
      const int st = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, neigh);
      if (st != GC_CUT && st != side) continue;
      mask[facet_id_of<D>(g, c[0], c[1], c[2], d, sd)] = 1;
      // counted once: by the CUT cell on the low side, or by this cell when the neighbour is not CUT
      if (st != GC_CUT || sd == 1) atomicAdd(count, 1ull);
    }
  }
}

__global__ void k_mask_to_flags(const int8_t* __restrict__ mask, int64_t n, uint32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = mask[i] ? 1u : 0u;
}

inline int64_t pad_pitch(int64_t n) { return ((n + 255) / 256) * 256; }

}  // namespace

extern "C" int gecut_ghost_skeleton(const gecut_result* cut, const int32_t* program, int len, int side, int64_t* n_ghost,
                                    int8_t* facet_to_mask, int32_t* facet_ids) {
  if (!cut) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = cut->ctx;
  const GcGrid& g = cut->grid;
  if (side != GECUT_IN && side != GECUT_OUT && side != GECUT_CUT) {
    ctx->err = "ghost skeleton: in_or_out must be IN, OUT or CUT";
    return GECUT_ERR_INVALID;
  }
  if (g.simplexified || g.k0 != 0 || g.k1 != g.n[g.D - 1]) {
    ctx->err = "GhostSkeleton is implemented for whole n-cube Cartesian models";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  DeviceGuard guard(ctx->device);
  GcProgram p;
  GC_CHECK(gc_make_program(ctx, program, len, cut->L, &p));
  const int D = g.D;
  const int64_t nf = D == 2 ? num_facets<2>(g) : num_facets<3>(g);
  const int64_t ncut = cut->sizes.num_cutcells;
  int8_t* mask = nullptr;
  unsigned long long* cnt = nullptr;
  GC_CHECK(gc_malloc(ctx, (void**)&mask, (size_t)pad_pitch(nf)));
  int st = gc_malloc(ctx, (void**)&cnt, sizeof(unsigned long long));
  auto done = [&](int code) {
    gc_free(ctx, mask);
    gc_free(ctx, cnt);
    return code;
  };
  if (st != GECUT_OK) return done(st);
  if (cudaMemsetAsync(mask, 0, (size_t)nf, ctx->stream) != cudaSuccess ||
      cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), ctx->stream) != cudaSuccess) {
    ctx->err = "ghost skeleton: memset failed";
    return done(GECUT_ERR_CUDA);
  }
  const int T = 256;
  if (ncut > 0) {
    if (D == 2)
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask<2><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(g, cut->lay, ncut, p, side, mask, cnt));
    else
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask<3><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(g, cut->lay, ncut, p, side, mask, cnt));
    st = gc_launch_check(ctx, "k_ghost_mask");
    if (st != GECUT_OK) return done(st);
  }
  uint64_t* h = (uint64_t*)ctx->h_pinned;
  if (cudaMemcpyAsync(h, cnt, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "ghost skeleton failed";
    return done(GECUT_ERR_CUDA);
  }
  const int64_t total = (int64_t)h[0];
  if (n_ghost) *n_ghost = total;
  if (facet_to_mask && cudaMemcpyAsync(facet_to_mask, mask, (size_t)nf, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
    ctx->err = "ghost skeleton: D2H copy failed";
    return done(GECUT_ERR_CUDA);
  }
  if (facet_ids && total > 0) {
    uint32_t* flags = nullptr;
    int32_t* ids = nullptr;
    int64_t n = 0;
    st = gc_malloc(ctx, (void**)&flags, sizeof(uint32_t) * nf);
    if (st != GECUT_OK) return done(st);
    GC_LAUNCH(ctx, "k_mask_to_flags", k_mask_to_flags<<<(unsigned)gc_div_up(nf, T), T, 0, ctx->stream>>>(mask, nf, flags));
    st = gc_launch_check(ctx, "k_mask_to_flags");
    if (st == GECUT_OK) st = gc_compact_flags(ctx, flags, nf, &ids, &n, nullptr, nullptr);
    gc_free(ctx, flags);
    if (st == GECUT_OK && n != total) {
      ctx->err = "ghost skeleton: id list and count disagree";
      st = GECUT_ERR_CUDA;
    }
    if (st == GECUT_OK && cudaMemcpyAsync(facet_ids, ids, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      st = GECUT_ERR_CUDA;
    cudaStreamSynchronize(ctx->stream);
    gc_free(ctx, ids);
    if (st != GECUT_OK) return done(st);
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "ghost skeleton failed";
    return done(GECUT_ERR_CUDA);
  }
  return done(GECUT_OK);
}

extern "C" int gecut_aggregate(const gecut_result* cut, const gecut_facet_result* facets, const int32_t* program,
                               int len, const int32_t* facet_program, int facet_len, int side, double threshold,
                               int32_t* cell_to_cellin, int on_device) {
  if (!cut || !facets || !cell_to_cellin) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = cut->ctx;
  if (facets->ctx != ctx) {
    ctx->err = "aggregate: the cell and facet discretizations belong to different contexts";
    return GECUT_ERR_INVALID;
  }
  if (!(threshold >= 0.0 && threshold <= 1.0)) {
    ctx->err = "Agg. threshold must be in [0,1]";  // AggregateCutCellsByThreshold (:64-68)
    return GECUT_ERR_INVALID;
  }
  if (side != GECUT_IN && side != GECUT_OUT) {
    ctx->err = "aggregate: in_or_out must be IN or OUT";
    return GECUT_ERR_INVALID;
  }
  const GcGrid& g = cut->grid;
  const GcGrid& gf = facets->grid;
  if (g.simplexified || g.k0 != 0 || g.k1 != g.n[g.D - 1]) {
    ctx->err = "aggregate is implemented for whole n-cube Cartesian models";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  if (g.D != gf.D || g.n[0] != gf.n[0] || g.n[1] != gf.n[1] || g.n[2] != gf.n[2] || cut->L != facets->L) {
    ctx->err = "aggregate: cut and cut_facets come from different models / geometries";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  GcProgram p;
  GC_CHECK(gc_make_program(ctx, program, len, cut->L, &p));
  // compute_bgfacet_to_inoutcut(cut_facets, geo) evaluates the geometry over cut_facets' OWN oid_to_ls (CellAggregation.jl:90):
  // its level-set rows may be numbered differently from the cell discretization's
  GcProgram pf = p;
  if (facet_program) GC_CHECK(gc_make_program(ctx, facet_program, facet_len, facets->L, &pf));
  const int64_t nc = cut->sizes.num_bgcells, ncut = cut->sizes.num_cutcells, nf = facets->sizes.num_bgfacets;
  const int D = g.D;
  int32_t* cellin = nullptr;
  uint8_t* touched = nullptr;
  int8_t* frow = nullptr;
  uint32_t* flag = nullptr;
  int st = GECUT_OK;
  auto done = [&](int code) {
    if (!on_device) gc_free(ctx, cellin);
    gc_free(ctx, touched);
    gc_free(ctx, frow);
    gc_free(ctx, flag);
    return code;
  };
  if (on_device) {
    cellin = cell_to_cellin;
  } else {
    st = gc_malloc(ctx, (void**)&cellin, sizeof(int32_t) * nc);
    if (st != GECUT_OK) return done(st);
  }
  st = gc_malloc(ctx, (void**)&touched, (size_t)nc);
  if (st == GECUT_OK) st = gc_malloc(ctx, (void**)&flag, sizeof(uint32_t));
  if (st != GECUT_OK) return done(st);
  // compute_bgfacet_to_inoutcut(cut_facets, geo): the level set's own row for a single leaf, else the program's row
  const int8_t* facet_state = nullptr;
  if (pf.len == 1) {
    facet_state = facets->lay.bg_state + (int64_t)pf.tok[0] * facets->lay.bg_pitch;
  } else {
    st = gc_malloc(ctx, (void**)&frow, (size_t)pad_pitch(nf));
    if (st == GECUT_OK) st = gc_eval_rows(ctx, facets->lay.bg_state, facets->lay.bg_pitch, nf, pf, frow, true);
    if (st != GECUT_OK) return done(st);
    facet_state = frow;
  }
  const int T = 256;
  GC_LAUNCH(ctx, "k_agg_init", k_agg_init<<<(unsigned)gc_div_up(nc, T), T, 0, ctx->stream>>>(cut->lay, nc, p, side, threshold, cellin, touched));
  st = gc_launch_check(ctx, "k_agg_init");
  if (st != GECUT_OK) return done(st);
  if (ncut > 0 && !cut->cut_sc_ptr) {
    ctx->err = "aggregate needs the subcells of cut(), not a classification-only result";
    return done(GECUT_ERR_INVALID);
  }
  if (ncut > 0) {
    // get_cell_measure(bgtrian): tensor Gauss rule of degree 2 on the n-cube, sum_q w_q |det J|
    const double detJ = g.dx[0] * g.dx[1] * (D == 3 ? g.dx[2] : 1.0);
    const int nq = 1 << D;
    const double w = D == 2 ? 0.25 : 0.125;
    double cell_meas = 0.0;
    for (int q = 0; q < nq; ++q) cell_meas += w * fabs(detJ);
    st = gc_ensure_measures(ctx, cut);
    if (st != GECUT_OK) return done(st);
    GC_LAUNCH(ctx, "k_agg_init_cut", k_agg_init_cut<<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(
        cut->lay, cut->cut_sc_ptr, ncut, p, side, threshold, cell_meas, cellin, touched));
    st = gc_launch_check(ctx, "k_agg_init_cut");
    if (st != GECUT_OK) return done(st);
    bool all_aggregated = false;
    for (int iter = 0; iter < 20 && !all_aggregated; ++iter) {  // max_iters (:153)
      if (cudaMemsetAsync(flag, 0, sizeof(uint32_t), ctx->stream) != cudaSuccess) return done(GECUT_ERR_CUDA);
      if (D == 2)
        GC_LAUNCH(ctx, "k_agg_sweep", k_agg_sweep<2><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(g, cut->lay, ncut, p, side, facet_state, cellin, touched, flag));
      else
        GC_LAUNCH(ctx, "k_agg_sweep", k_agg_sweep<3><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(g, cut->lay, ncut, p, side, facet_state, cellin, touched, flag));
      st = gc_launch_check(ctx, "k_agg_sweep");
      if (st != GECUT_OK) return done(st);
      uint32_t* h = (uint32_t*)ctx->h_pinned;
      if (cudaMemcpyAsync(h, flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        ctx->err = "aggregate: sweep failed";
        return done(GECUT_ERR_CUDA);
      }
      all_aggregated = h[0] == 0u;
      if (!all_aggregated) {
        GC_LAUNCH(ctx, "k_agg_touch", k_agg_touch<<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(cut->lay, ncut, cellin, touched));
        st = gc_launch_check(ctx, "k_agg_touch");
        if (st != GECUT_OK) return done(st);
      }
    }
    if (!all_aggregated) {
      ctx->err = "aggregate: not all cut cells could be aggregated in 20 sweeps (reference: @assert all_aggregated)";
      return done(GECUT_ERR_INVALID);
    }
  }
  if (!on_device &&
      cudaMemcpyAsync(cell_to_cellin, cellin, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
    ctx->err = "aggregate: D2H copy failed";
    return done(GECUT_ERR_CUDA);
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "aggregate failed";
    return done(GECUT_ERR_CUDA);
  }
  return done(GECUT_OK);
}
