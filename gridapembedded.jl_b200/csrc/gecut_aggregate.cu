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
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "gecut_facets.cuh"

namespace {

// cell_to_cellin / touched for every cell: roots are the cells with unit cut measure >= threshold (:141-146); uncut
// cells have unit measure 1 (state == loc) or 0
// z-slabs: the arrays hold the slab's own layers behind an optional ghost layer (`shift` = cells of the lower ghost layer);
// roots are GLOBAL cell ids (`goff` = cells below the slab), as the distributed reference stores them
// (src/Distributed/DistributedAgFEM.jl:101-105).  Whole model: shift = goff = 0.
__global__ void k_agg_init(const GcLayout lay, int64_t nc, const GcProgram prog, int loc, double threshold,
                           int64_t shift, int64_t goff, int32_t* __restrict__ cellin, uint8_t* __restrict__ touched) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int st = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, c);
  const double unit = st == loc ? 1.0 : 0.0;  // CUT cells are overwritten by k_agg_init_cut
  const bool root = st != GC_CUT && unit >= threshold;
  cellin[c + shift] = root ? (int32_t)(c + goff + 1) : 0;
  touched[c + shift] = root ? 1 : 0;
}

__global__ void k_agg_init_cut(const GcLayout lay, const uint32_t* __restrict__ sc_ptr, int64_t ncut, const GcProgram prog,
                               int loc, double threshold, double cell_meas, int64_t shift, int64_t goff,
                               int32_t* __restrict__ cellin, uint8_t* __restrict__ touched) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncut) return;
  const int64_t c = (int64_t)lay.cutcells[k] - 1;
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, c) != GC_CUT) return;
  double V = 0.0;
  const uint32_t s0 = sc_ptr[2 * k], s1 = sc_ptr[2 * k + 2];
  for (uint32_t e = s0; e < s1; ++e)
    if (gc_eval_inoutcut(prog, lay.sc_state, lay.sc_pitch, (int64_t)e) == loc) V += lay.sc_meas[e];
  const bool root = V / cell_meas >= threshold;
  cellin[c + shift] = root ? (int32_t)(c + goff + 1) : 0;
  touched[c + shift] = root ? 1 : 0;
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

// one sweep (:155-176): untouched CUT cells look for the touched face neighbour whose root is closest.  `g` is the cut's grid
// (global sizes, slab [k0, k1)), `gf` the facet grid of the rank's local model (local numbering); neighbours across the
// slab's lower / upper plane live in the ghost layers (`glo`, `ghi`: the layer exists), roots are global cell ids.
template <int D>
__global__ void k_agg_sweep(const GcGrid g, const GcGrid gf, const GcLayout lay, int64_t ncut, const GcProgram prog, int loc,
                            const int8_t* __restrict__ facet_state, int64_t shift, int glo, int ghi,
                            int32_t* __restrict__ cellin, const uint8_t* __restrict__ touched,
                            uint32_t* __restrict__ not_aggregated) {
  const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= ncut) return;
  const int64_t cell = (int64_t)lay.cutcells[kk] - 1;  // slab-local
  if (touched[cell + shift]) return;
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell) != GC_CUT) return;
  const int nl = g.k1 - g.k0;  // own layers
  int c[3];                    // local: the last index counts from the slab's first layer
  c[0] = (int)(cell % g.n[0]);
  c[1] = D == 3 ? (int)((cell / g.n[0]) % g.n[1]) : (int)(cell / g.n[0]);
  c[2] = D == 3 ? (int)(cell / ((int64_t)g.n[0] * g.n[1])) : 0;
  int cg[3] = {c[0], c[1], c[2]};  // global
  cg[D - 1] += g.k0;
  double dmin = INFINITY;
  int64_t best = -1;
  // faces in Gridap's local order: highest direction first, low then high
#pragma unroll
  for (int d = D - 1; d >= 0; --d) {
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const int64_t f = facet_id_of<D>(gf, c[0], c[1], c[2], d, side);
      const int io = facet_state[f];
      if (io != GC_CUT && io != loc) continue;
      int nb[3] = {c[0], c[1], c[2]};
      nb[d] += side == 0 ? -1 : 1;
      if (d == D - 1) {
        // boundary facet of the MODEL: the cell itself is the only one around; of the slab: the neighbour is a ghost
        if ((nb[d] < 0 && !glo) || (nb[d] >= nl && !ghi)) continue;
      } else if (nb[d] < 0 || nb[d] >= g.n[d]) {
        continue;
      }
      const int64_t neigh = (D == 3 ? (int64_t)nb[0] + (int64_t)g.n[0] * ((int64_t)nb[1] + (int64_t)g.n[1] * nb[2])
                                    : (int64_t)nb[0] + (int64_t)g.n[0] * nb[1]) + shift;
      if (!touched[neigh]) continue;
      const int64_t root = (int64_t)cellin[neigh] - 1;  // global
      int r[3];
      r[0] = (int)(root % g.n[0]);
      r[1] = D == 3 ? (int)((root / g.n[0]) % g.n[1]) : (int)(root / g.n[0]);
      r[2] = D == 3 ? (int)(root / ((int64_t)g.n[0] * g.n[1])) : 0;
      const double dist = max_vertex_distance<D>(g, cg, r);
      if (__dmul_rn(1.0 + 1.0e-9, dist) < dmin) {
        dmin = dist;
        best = neigh;
      }
    }
  }
  if (best >= 0)
    cellin[cell + shift] = cellin[best];  // `best` is touched: its root is final, no race with other threads of the sweep
  else
    atomicAdd(not_aggregated, 1u);
}

// _touch_aggregated_cells! (:235-241) restricted to the cells that can change (cut cells)
__global__ void k_agg_touch(const GcLayout lay, int64_t ncut, int64_t shift, const int32_t* __restrict__ cellin,
                            uint8_t* __restrict__ touched) {
  const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= ncut) return;
  const int64_t cell = (int64_t)lay.cutcells[kk] - 1 + shift;
  if (cellin[cell] > 0) touched[cell] = 1;
}

// GhostSkeleton mask (_fill_ghost_skeleton_mask!, src/Interfaces/EmbeddedDiscretizations.jl:520-543): interior facets
// with at least one CUT cell around and both cells active (CUT or `side`).  Every such facet touches a CUT cell, so
// the O(n^2) cut cells drive the pass; a facet between two CUT cells is written by both with the same value.
template <int D>
__global__ void k_ghost_mask(const GcGrid g, const GcGrid gf, const GcLayout lay, int64_t ncut, const GcProgram prog, int side,
                             const int8_t* __restrict__ lo_rows, int64_t lo_pitch, int64_t lo_base,
                             const int8_t* __restrict__ hi_rows, int64_t hi_pitch, int64_t hi_base,
                             int8_t* __restrict__ mask, unsigned long long* __restrict__ count) {
  const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= ncut) return;
  const int64_t cell = (int64_t)lay.cutcells[kk] - 1;  // slab-local
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell) != GC_CUT) return;
  const int nl = g.k1 - g.k0;
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
      int st;
      if (d == D - 1 && (nb[d] < 0 || nb[d] >= nl)) {
        // across the slab's lower / upper plane: the neighbour's states come from the adjacent slab (its last / first layer);
        // no adjacent slab = boundary of the model
        const int8_t* rows = nb[d] < 0 ? lo_rows : hi_rows;
        if (!rows) continue;
        const int64_t in_layer = D == 3 ? (int64_t)nb[0] + (int64_t)g.n[0] * nb[1] : (int64_t)nb[0];
        st = gc_eval_inoutcut(prog, rows, nb[d] < 0 ? lo_pitch : hi_pitch, (nb[d] < 0 ? lo_base : hi_base) + in_layer);
      } else {
        if (nb[d] < 0 || nb[d] >= g.n[d]) continue;
        const int64_t neigh = D == 3 ? (int64_t)nb[0] + (int64_t)g.n[0] * ((int64_t)nb[1] + (int64_t)g.n[1] * nb[2])
                                     : (int64_t)nb[0] + (int64_t)g.n[0] * nb[1];
        st = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, neigh);
      }
      if (st != GC_CUT && st != side) continue;
      mask[facet_id_of<D>(gf, c[0], c[1], c[2], d, sd)] = 1;
      // whole model: counted once -- by the CUT cell on the low side, or by this cell when the neighbour is not CUT
      if (count && (st != GC_CUT || sd == 1)) atomicAdd(count, 1ull);
    }
  }
}

// facets on the plane a slab shares with its lower (which = 0) / upper (which = 1) neighbour: one thread per cell of the slab's
// boundary layer; the facet is a ghost facet when one of the two cells around it is CUT and both are active -- also when the
// CUT cell is the neighbour's (the cut-cell driven pass above only sees this slab's own cut cells)
template <int D>
__global__ void k_ghost_mask_plane(const GcGrid g, const GcGrid gf, const GcLayout lay, const GcProgram prog, int side, int which,
                                   const int8_t* __restrict__ nb_rows, int64_t nb_pitch, int64_t nb_base,
                                   int8_t* __restrict__ mask) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.layer_cubes) return;
  const int nl = g.k1 - g.k0;
  const int64_t cell = which == 0 ? e : (int64_t)(nl - 1) * g.layer_cubes + e;
  const int so = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell);
  const int sn = gc_eval_inoutcut(prog, nb_rows, nb_pitch, nb_base + e);
  if (so != GC_CUT && sn != GC_CUT) return;
  if ((so != GC_CUT && so != side) || (sn != GC_CUT && sn != side)) return;
  const int i = (int)(e % g.n[0]), j = D == 3 ? (int)(e / g.n[0]) : 0;
  const int c[3] = {i, D == 3 ? j : (which == 0 ? 0 : nl - 1), D == 3 ? (which == 0 ? 0 : nl - 1) : 0};
  mask[facet_id_of<D>(gf, c[0], c[1], c[2], D - 1, which)] = 1;
}

// ------------------------------------------------------------------------------------------------
// simplexified models: cell = nt * cube + simplex; the neighbour across local facet lf is the slot table's partner --
// another simplex of the same cube, or a simplex of the cube across the face (gecut_facets.cuh)
// ------------------------------------------------------------------------------------------------
struct SxNeighbour {
  int cube[3];  // slab-local cube of the neighbour (the last index may be -1 / nl: ghost layer)
  int t;        // its simplex
  bool exists;  // false: boundary facet of the model in a direction without slabs
};

template <int D>
__device__ __forceinline__ SxNeighbour sx_neighbour(const GcGrid& g, const GcSxFacets& sx, const int* c, int slot) {
  SxNeighbour r;
  r.cube[0] = c[0];
  r.cube[1] = c[1];
  r.cube[2] = c[2];
  r.t = sx.partner[slot] / sx.nlf;
  r.exists = true;
  const int fc = sx.face[slot];
  if (fc >= 0) {
    const int d = fc >> 1;
    r.cube[d] += (fc & 1) ? 1 : -1;
    if (d != D - 1 && (r.cube[d] < 0 || r.cube[d] >= g.n[d])) r.exists = false;
  }
  return r;
}

template <int D>
__device__ __forceinline__ double max_vertex_distance_sx(const GcGrid& g, const uint8_t* __restrict__ simp_raw, const int* a,
                                                         int ta, const int* b, int tb) {
  double d = 0.0;
  for (int p = 0; p <= D; ++p) {
    const int lp = simp_raw[(D - 2) * 24 + ta * 4 + p];
    for (int q = 0; q <= D; ++q) {
      const int lq = simp_raw[(D - 2) * 24 + tb * 4 + q];
      double s2 = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double w = __dsub_rn(gc_coord(g, c, a[c] + ((lp >> c) & 1)), gc_coord(g, c, b[c] + ((lq >> c) & 1)));
        s2 = __dadd_rn(s2, __dmul_rn(w, w));
      }
      d = fmax(d, __dsqrt_rn(s2));
    }
  }
  return d;
}

template <int D>
__global__ void k_agg_sweep_sx(const GcGrid g, const GcGrid gf, const GcSxFacets sx, const uint8_t* __restrict__ simp_raw,
                               const GcLayout lay, int64_t ncut, const GcProgram prog, int loc,
                               const int8_t* __restrict__ facet_state, int64_t shift, int glo, int ghi,
                               int32_t* __restrict__ cellin, const uint8_t* __restrict__ touched,
                               uint32_t* __restrict__ not_aggregated) {
  const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= ncut) return;
  const int64_t cell = (int64_t)lay.cutcells[kk] - 1;  // slab-local
  if (touched[cell + shift]) return;
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell) != GC_CUT) return;
  const int nl = g.k1 - g.k0, nt = sx.nt;
  const int64_t cube = cell / nt;
  const int t = (int)(cell % nt);
  int c[3];
  c[0] = (int)(cube % g.n[0]);
  c[1] = D == 3 ? (int)((cube / g.n[0]) % g.n[1]) : (int)(cube / g.n[0]);
  c[2] = D == 3 ? (int)(cube / ((int64_t)g.n[0] * g.n[1])) : 0;
  int cg[3] = {c[0], c[1], c[2]};
  cg[D - 1] += g.k0;
  double dmin = INFINITY;
  int64_t best = -1;
  for (int lf = 0; lf <= D; ++lf) {  // local facets in Gridap's order
    const int slot = t * sx.nlf + lf;
    const int64_t f = sx_facet_id_of<D>(gf, sx, c[0], c[1], c[2], slot);
    const int io = facet_state[f];
    if (io != GC_CUT && io != loc) continue;
    const SxNeighbour nb = sx_neighbour<D>(g, sx, c, slot);
    if (!nb.exists) continue;
    if ((nb.cube[D - 1] < 0 && !glo) || (nb.cube[D - 1] >= nl && !ghi)) continue;
    const int64_t ncube = D == 3 ? (int64_t)nb.cube[0] + (int64_t)g.n[0] * ((int64_t)nb.cube[1] + (int64_t)g.n[1] * nb.cube[2])
                                 : (int64_t)nb.cube[0] + (int64_t)g.n[0] * nb.cube[1];
    const int64_t neigh = ncube * nt + nb.t + shift;
    if (!touched[neigh]) continue;
    const int64_t root = (int64_t)cellin[neigh] - 1;  // global
    const int64_t rcube = root / nt;
    const int rt = (int)(root % nt);
    int r[3];
    r[0] = (int)(rcube % g.n[0]);
    r[1] = D == 3 ? (int)((rcube / g.n[0]) % g.n[1]) : (int)(rcube / g.n[0]);
    r[2] = D == 3 ? (int)(rcube / ((int64_t)g.n[0] * g.n[1])) : 0;
    const double dist = max_vertex_distance_sx<D>(g, simp_raw, cg, t, r, rt);
    if (__dmul_rn(1.0 + 1.0e-9, dist) < dmin) {
      dmin = dist;
      best = neigh;
    }
  }
  if (best >= 0)
    cellin[cell + shift] = cellin[best];
  else
    atomicAdd(not_aggregated, 1u);
}

// unit cut measure of the cut simplices.  get_cell_measure of a simplexified model integrates over every cell's OWN node
// coordinates (an UnstructuredGrid: |det J| from pmin + i * dx differences, not from dx), so the measure of a cell varies
// in the last bits from cube to cube -- and a cut cell whose `loc` part is the whole cell (surface through its vertices)
// sits exactly on the threshold.  Evaluated per cell with the operation order of the oracle / Gridap.
template <int D>
__global__ void k_agg_init_cut_sx(const GcGrid g, const uint8_t* __restrict__ simp_raw, const GcLayout lay,
                                  const uint32_t* __restrict__ sc_ptr, int64_t ncut, const GcProgram prog, int loc,
                                  double threshold, int nt, int64_t shift, int64_t goff, int32_t* __restrict__ cellin,
                                  uint8_t* __restrict__ touched) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ncut) return;
  const int64_t c = (int64_t)lay.cutcells[k] - 1;
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, c) != GC_CUT) return;
  double V = 0.0;
  const uint32_t s0 = sc_ptr[2 * k], s1 = sc_ptr[2 * k + 2];
  for (uint32_t e = s0; e < s1; ++e)
    if (gc_eval_inoutcut(prog, lay.sc_state, lay.sc_pitch, (int64_t)e) == loc) V += lay.sc_meas[e];
  const int64_t cube = c / nt;
  const int t = (int)(c % nt);
  int ijk[3];
  ijk[0] = (int)(cube % g.n[0]);
  ijk[1] = D == 3 ? (int)((cube / g.n[0]) % g.n[1]) : (int)(cube / g.n[0]);
  ijk[2] = D == 3 ? (int)(cube / ((int64_t)g.n[0] * g.n[1])) : 0;
  ijk[D - 1] += g.k0;
  double x[D + 1][3];
#pragma unroll
  for (int m = 0; m <= D; ++m) {
    const int lv = simp_raw[(D - 2) * 24 + t * 4 + m];
#pragma unroll
    for (int d = 0; d < 3; ++d) x[m][d] = d < D ? gc_coord(g, d, ijk[d] + ((lv >> d) & 1)) : 0.0;
  }
  const double cell_meas = integrate_one<D>(simplex_meas<D, D>(x));
  const bool root = V / cell_meas >= threshold;
  cellin[c + shift] = root ? (int32_t)(c + goff + 1) : 0;
  touched[c + shift] = root ? 1 : 0;
}

template <int D>
__global__ void k_ghost_mask_sx(const GcGrid g, const GcGrid gf, const GcSxFacets sx, const GcLayout lay, int64_t ncut,
                                const GcProgram prog, int side, const int8_t* __restrict__ lo_rows, int64_t lo_pitch,
                                int64_t lo_base, const int8_t* __restrict__ hi_rows, int64_t hi_pitch, int64_t hi_base,
                                int8_t* __restrict__ mask, unsigned long long* __restrict__ count) {
  const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= ncut) return;
  const int64_t cell = (int64_t)lay.cutcells[kk] - 1;  // slab-local
  if (gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cell) != GC_CUT) return;
  const int nl = g.k1 - g.k0, nt = sx.nt;
  const int64_t cube = cell / nt;
  const int t = (int)(cell % nt);
  int c[3];
  c[0] = (int)(cube % g.n[0]);
  c[1] = D == 3 ? (int)((cube / g.n[0]) % g.n[1]) : (int)(cube / g.n[0]);
  c[2] = D == 3 ? (int)(cube / ((int64_t)g.n[0] * g.n[1])) : 0;
  for (int lf = 0; lf <= D; ++lf) {
    const int slot = t * sx.nlf + lf;
    const SxNeighbour nb = sx_neighbour<D>(g, sx, c, slot);
    if (!nb.exists) continue;
    int st;
    int64_t neigh = -1;
    if (nb.cube[D - 1] < 0 || nb.cube[D - 1] >= nl) {
      const int8_t* rows = nb.cube[D - 1] < 0 ? lo_rows : hi_rows;
      if (!rows) continue;  // boundary of the model
      const int64_t in_layer = D == 3 ? (int64_t)nb.cube[0] + (int64_t)g.n[0] * nb.cube[1] : (int64_t)nb.cube[0];
      st = gc_eval_inoutcut(prog, rows, nb.cube[D - 1] < 0 ? lo_pitch : hi_pitch,
                            (nb.cube[D - 1] < 0 ? lo_base : hi_base) + in_layer * nt + nb.t);
    } else {
      const int64_t ncube = D == 3 ? (int64_t)nb.cube[0] + (int64_t)g.n[0] * ((int64_t)nb.cube[1] + (int64_t)g.n[1] * nb.cube[2])
                                   : (int64_t)nb.cube[0] + (int64_t)g.n[0] * nb.cube[1];
      neigh = ncube * nt + nb.t;
      st = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, neigh);
    }
    if (st != GC_CUT && st != side) continue;
    mask[sx_facet_id_of<D>(gf, sx, c[0], c[1], c[2], slot)] = 1;
    // whole model: a facet between two CUT cells is counted by the one with the lower id
    if (count && (st != GC_CUT || neigh > cell)) atomicAdd(count, 1ull);
  }
}

// facets on the plane a slab shares with its lower (which = 0) / upper (which = 1) neighbour, a thread per cube of the
// slab's boundary layer (see k_ghost_mask_plane)
template <int D>
__global__ void k_ghost_mask_plane_sx(const GcGrid g, const GcGrid gf, const GcSxFacets sx, const GcLayout lay,
                                      const GcProgram prog, int side, int which, const int8_t* __restrict__ nb_rows,
                                      int64_t nb_pitch, int64_t nb_base, int8_t* __restrict__ mask) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.layer_cubes) return;
  const int nl = g.k1 - g.k0, nt = sx.nt;
  const int64_t cube = which == 0 ? e : (int64_t)(nl - 1) * g.layer_cubes + e;
  const int i = (int)(e % g.n[0]), j = D == 3 ? (int)(e / g.n[0]) : 0;
  const int c[3] = {i, D == 3 ? j : (which == 0 ? 0 : nl - 1), D == 3 ? (which == 0 ? 0 : nl - 1) : 0};
  for (int slot = 0; slot < sx.nslots; ++slot) {
    if (sx.face[slot] != (((D - 1) << 1) | which)) continue;
    const int so = gc_eval_inoutcut(prog, lay.bg_state, lay.bg_pitch, cube * nt + slot / sx.nlf);
    const int sn = gc_eval_inoutcut(prog, nb_rows, nb_pitch, nb_base + e * nt + sx.partner[slot] / sx.nlf);
    if (so != GC_CUT && sn != GC_CUT) continue;
    if ((so != GC_CUT && so != side) || (sn != GC_CUT && sn != side)) continue;
    mask[sx_facet_id_of<D>(gf, sx, c[0], c[1], c[2], slot)] = 1;
  }
}

// number of marked facets (slab version: a facet on a plane two slabs share is marked -- and counted -- in both local models)
__global__ void __launch_bounds__(256) k_count_mask(const int8_t* __restrict__ mask, int64_t n, unsigned long long* __restrict__ count) {
  unsigned int mine = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) mine += mask[i] ? 1u : 0u;
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, (unsigned long long)mine);
}

__global__ void k_mask_to_flags(const int8_t* __restrict__ mask, int64_t n, uint32_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = mask[i] ? 1u : 0u;
}

inline int64_t pad_pitch(int64_t n) { return ((n + 255) / 256) * 256; }
// bg cells of one layer of the last direction (a layer of cubes holds nt simplices per cube when simplexified)
inline int64_t layer_cells(const GcGrid& g) { return g.layer_cubes * (g.simplexified ? g.nt : 1); }

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
  if (g.k0 != 0 || g.k1 != g.n[g.D - 1]) {
    ctx->err = "GhostSkeleton of a z-slab: use gecut_ghost_skeleton_slabs / gecut_comm_ghost_skeleton";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  DeviceGuard guard(ctx->device);
  GcProgram p;
  GC_CHECK(gc_make_program(ctx, program, len, cut->L, &p));
  const int D = g.D;
  GcSxFacets sx{};
  if (g.simplexified && !gc_build_sx_facets(D, &sx)) {
    ctx->err = "ghost skeleton: the simplexify table of the n-cube does not tile conformingly";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  const int64_t nf = g.simplexified ? (D == 2 ? sx_num_facets<2>(g, sx) : sx_num_facets<3>(g, sx))
                                    : (D == 2 ? num_facets<2>(g) : num_facets<3>(g));
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
    if (g.simplexified && D == 2)
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask_sx<2><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(g, g, sx, cut->lay, ncut, p, side, nullptr, 0, 0, nullptr, 0, 0, mask, cnt));
    else if (g.simplexified)
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask_sx<3><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(g, g, sx, cut->lay, ncut, p, side, nullptr, 0, 0, nullptr, 0, 0, mask, cnt));
    else if (D == 2)
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask<2><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(g, g, cut->lay, ncut, p, side, nullptr, 0, 0, nullptr, 0, 0, mask, cnt));
    else
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask<3><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(g, g, cut->lay, ncut, p, side, nullptr, 0, 0, nullptr, 0, 0, mask, cnt));
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

// ------------------------------------------------------------------------------------------------
// Aggregation over one or several z-slabs.  A sweep of the reference is a Jacobi step (a cell reads only `touched` flags and
// roots frozen during the sweep), so the distributed algorithm (src/Distributed/DistributedAgFEM.jl:86-144: own cells
// swept, then `consistent!` on touched / cellin, `reduction!(&, all_aggregated)`) gives exactly the roots of the serial one
// when the ghost layers are refreshed after every sweep.  One AggSlab per slab; the transport of the ghost layers is the
// only difference between the whole model (none), several slabs on one GPU (device copies) and one slab per rank (NCCL).
// ------------------------------------------------------------------------------------------------
struct AggSlab {
  const gecut_result* cut = nullptr;
  const gecut_facet_result* fac = nullptr;
  GcProgram p{}, pf{};
  int32_t* cellin = nullptr;  // [ghost lo][own layers][ghost hi]
  uint8_t* touched = nullptr;
  int8_t* frow = nullptr;
  const int8_t* facet_state = nullptr;
  uint32_t* flag = nullptr;
  int64_t layer = 0, nc = 0, ncut = 0, shift = 0, goff = 0, ntot = 0;
  int glo = 0, ghi = 0;
  bool owns_cellin = true;
  GcSxFacets sx{};  // simplexified models
};

static void agg_release(gecut_ctx* ctx, AggSlab& a) {
  if (a.owns_cellin) gc_free(ctx, a.cellin);
  gc_free(ctx, a.touched);
  gc_free(ctx, a.frow);
  gc_free(ctx, a.flag);
  a.cellin = nullptr;
  a.touched = nullptr;
  a.frow = nullptr;
  a.flag = nullptr;
}

// validation + allocation + the two initialisation passes of one slab
static int agg_setup(gecut_ctx* ctx, AggSlab& a, const gecut_result* cut, const gecut_facet_result* facets,
                     const int32_t* program, int len, const int32_t* facet_program, int facet_len, int side,
                     double threshold, int32_t* dev_out /* whole model only: write straight into the caller's array */) {
  a.cut = cut;
  a.fac = facets;
  if (facets->ctx != ctx || cut->ctx != ctx) {
    ctx->err = "aggregate: the cell and facet discretizations belong to different contexts";
    return GECUT_ERR_INVALID;
  }
  const GcGrid& g = cut->grid;
  const GcGrid& gf = facets->grid;
  const int D = g.D;
  if (g.simplexified && !gc_build_sx_facets(D, &a.sx)) {
    ctx->err = "aggregate: the simplexify table of the n-cube does not tile conformingly";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  bool same = g.D == gf.D && cut->L == facets->L && gf.k0 == g.k0 && gf.n[D - 1] == g.k1 - g.k0 &&
              g.simplexified == gf.simplexified;
  for (int d = 0; d + 1 < D; ++d) same = same && g.n[d] == gf.n[d];
  if (!same) {
    ctx->err = "aggregate: cut and cut_facets come from different models / slabs / geometries";
    return GECUT_ERR_INVALID;
  }
  GC_CHECK(gc_make_program(ctx, program, len, cut->L, &a.p));
  // compute_bgfacet_to_inoutcut(cut_facets, geo) evaluates the geometry over cut_facets' OWN oid_to_ls (CellAggregation.jl:90):
  // its level-set rows may be numbered differently from the cell discretization's
  a.pf = a.p;
  if (facet_program) GC_CHECK(gc_make_program(ctx, facet_program, facet_len, facets->L, &a.pf));
  a.nc = cut->sizes.num_bgcells;
  a.ncut = cut->sizes.num_cutcells;
  a.layer = layer_cells(g);
  a.glo = g.k0 > 0 ? 1 : 0;
  a.ghi = g.k1 < g.n[D - 1] ? 1 : 0;
  a.shift = a.glo ? a.layer : 0;
  a.goff = (int64_t)g.k0 * a.layer;
  a.ntot = a.nc + (a.glo + a.ghi) * a.layer;
  if (a.ncut > 0 && !cut->cut_sc_ptr) {
    ctx->err = "aggregate needs the subcells of cut(), not a classification-only result";
    return GECUT_ERR_INVALID;
  }
  if (dev_out && a.ntot == a.nc) {
    a.cellin = dev_out;
    a.owns_cellin = false;
  } else {
    GC_CHECK(gc_malloc(ctx, (void**)&a.cellin, sizeof(int32_t) * a.ntot));
  }
  GC_CHECK(gc_malloc(ctx, (void**)&a.touched, (size_t)a.ntot));
  GC_CHECK(gc_malloc(ctx, (void**)&a.flag, sizeof(uint32_t)));
  if (a.ntot != a.nc) {  // ghost layers start untouched; the first exchange fills them
    GC_CUDA(cudaMemsetAsync(a.cellin, 0, sizeof(int32_t) * a.ntot, ctx->stream));
    GC_CUDA(cudaMemsetAsync(a.touched, 0, (size_t)a.ntot, ctx->stream));
  }
  const int64_t nf = facets->sizes.num_bgfacets;
  if (a.pf.len == 1) {  // the level set's own row for a single leaf, else the program's row
    a.facet_state = facets->lay.bg_state + (int64_t)a.pf.tok[0] * facets->lay.bg_pitch;
  } else {
    GC_CHECK(gc_malloc(ctx, (void**)&a.frow, (size_t)pad_pitch(nf)));
    GC_CHECK(gc_eval_rows(ctx, facets->lay.bg_state, facets->lay.bg_pitch, nf, a.pf, a.frow, true));
    a.facet_state = a.frow;
  }
  const int T = 256;
  GC_LAUNCH(ctx, "k_agg_init", k_agg_init<<<(unsigned)gc_div_up(a.nc, T), T, 0, ctx->stream>>>(
      cut->lay, a.nc, a.p, side, threshold, a.shift, a.goff, a.cellin, a.touched));
  GC_CHECK(gc_launch_check(ctx, "k_agg_init"));
  if (a.ncut > 0) {
    // get_cell_measure(bgtrian): tensor Gauss rule of degree 2 on the n-cube, sum_q w_q |det J|
    const double detJ = g.dx[0] * g.dx[1] * (D == 3 ? g.dx[2] : 1.0);
    const int nq = 1 << D;
    const double w = D == 2 ? 0.25 : 0.125;
    double cell_meas = 0.0;
    for (int q = 0; q < nq; ++q) cell_meas += w * fabs(detJ);
    GC_CHECK(gc_ensure_measures(ctx, cut));
    if (g.simplexified) {
      if (D == 2)
        GC_LAUNCH(ctx, "k_agg_init_cut", k_agg_init_cut_sx<2><<<(unsigned)gc_div_up(a.ncut, T), T, 0, ctx->stream>>>(
            g, ctx->d_simplexify_raw, cut->lay, cut->cut_sc_ptr, a.ncut, a.p, side, threshold, g.nt, a.shift, a.goff, a.cellin,
            a.touched));
      else
        GC_LAUNCH(ctx, "k_agg_init_cut", k_agg_init_cut_sx<3><<<(unsigned)gc_div_up(a.ncut, T), T, 0, ctx->stream>>>(
            g, ctx->d_simplexify_raw, cut->lay, cut->cut_sc_ptr, a.ncut, a.p, side, threshold, g.nt, a.shift, a.goff, a.cellin,
            a.touched));
    } else
    GC_LAUNCH(ctx, "k_agg_init_cut", k_agg_init_cut<<<(unsigned)gc_div_up(a.ncut, T), T, 0, ctx->stream>>>(
        cut->lay, cut->cut_sc_ptr, a.ncut, a.p, side, threshold, cell_meas, a.shift, a.goff, a.cellin, a.touched));
    GC_CHECK(gc_launch_check(ctx, "k_agg_init_cut"));
  }
  return GECUT_OK;
}

static int agg_sweep(gecut_ctx* ctx, AggSlab& a, int side) {
  GC_CUDA(cudaMemsetAsync(a.flag, 0, sizeof(uint32_t), ctx->stream));
  if (a.ncut == 0) return GECUT_OK;
  const int T = 256;
  const GcGrid& g = a.cut->grid;
  if (g.simplexified) {
    if (g.D == 2)
      GC_LAUNCH(ctx, "k_agg_sweep", k_agg_sweep_sx<2><<<(unsigned)gc_div_up(a.ncut, T), T, 0, ctx->stream>>>(
          g, a.fac->grid, a.sx, ctx->d_simplexify_raw, a.cut->lay, a.ncut, a.p, side, a.facet_state, a.shift, a.glo, a.ghi,
          a.cellin, a.touched, a.flag));
    else
      GC_LAUNCH(ctx, "k_agg_sweep", k_agg_sweep_sx<3><<<(unsigned)gc_div_up(a.ncut, T), T, 0, ctx->stream>>>(
          g, a.fac->grid, a.sx, ctx->d_simplexify_raw, a.cut->lay, a.ncut, a.p, side, a.facet_state, a.shift, a.glo, a.ghi,
          a.cellin, a.touched, a.flag));
    return gc_launch_check(ctx, "k_agg_sweep");
  }
  if (g.D == 2)
    GC_LAUNCH(ctx, "k_agg_sweep", k_agg_sweep<2><<<(unsigned)gc_div_up(a.ncut, T), T, 0, ctx->stream>>>(
        g, a.fac->grid, a.cut->lay, a.ncut, a.p, side, a.facet_state, a.shift, a.glo, a.ghi, a.cellin, a.touched, a.flag));
  else
    GC_LAUNCH(ctx, "k_agg_sweep", k_agg_sweep<3><<<(unsigned)gc_div_up(a.ncut, T), T, 0, ctx->stream>>>(
        g, a.fac->grid, a.cut->lay, a.ncut, a.p, side, a.facet_state, a.shift, a.glo, a.ghi, a.cellin, a.touched, a.flag));
  return gc_launch_check(ctx, "k_agg_sweep");
}

static int agg_touch(gecut_ctx* ctx, AggSlab& a) {
  if (a.ncut == 0) return GECUT_OK;
  const int T = 256;
  GC_LAUNCH(ctx, "k_agg_touch", k_agg_touch<<<(unsigned)gc_div_up(a.ncut, T), T, 0, ctx->stream>>>(a.cut->lay, a.ncut, a.shift, a.cellin, a.touched));
  return gc_launch_check(ctx, "k_agg_touch");
}

static int agg_copy_out(gecut_ctx* ctx, AggSlab& a, int32_t* out, int on_device) {
  if (!a.owns_cellin) return GECUT_OK;  // written in place
  GC_CUDA(cudaMemcpyAsync(out, a.cellin + a.shift, sizeof(int32_t) * a.nc, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                          ctx->stream));
  return GECUT_OK;
}

static int agg_check_args(gecut_ctx* ctx, int side, double threshold) {
  if (!(threshold >= 0.0 && threshold <= 1.0)) {
    ctx->err = "Agg. threshold must be in [0,1]";  // AggregateCutCellsByThreshold (:64-68)
    return GECUT_ERR_INVALID;
  }
  if (side != GECUT_IN && side != GECUT_OUT) {
    ctx->err = "aggregate: in_or_out must be IN or OUT";
    return GECUT_ERR_INVALID;
  }
  return GECUT_OK;
}

// slabs of ONE context (one GPU), in ascending layer order, together covering the model; ghost layers by device copies
static int agg_run_local(gecut_ctx* ctx, std::vector<AggSlab>& S, int side, int32_t* const* outs, int on_device) {
  const int n = (int)S.size();
  auto exchange = [&]() -> int {
    for (int r = 0; r + 1 < n; ++r) {
      AggSlab &lo = S[r], &hi = S[r + 1];
      const int64_t L = lo.layer;
      // last own layer of `lo` -> lower ghost of `hi`; first own layer of `hi` -> upper ghost of `lo`
      GC_CUDA(cudaMemcpyAsync(hi.cellin, lo.cellin + lo.shift + lo.nc - L, sizeof(int32_t) * L, cudaMemcpyDeviceToDevice, ctx->stream));
      GC_CUDA(cudaMemcpyAsync(hi.touched, lo.touched + lo.shift + lo.nc - L, (size_t)L, cudaMemcpyDeviceToDevice, ctx->stream));
      GC_CUDA(cudaMemcpyAsync(lo.cellin + lo.shift + lo.nc, hi.cellin + hi.shift, sizeof(int32_t) * L, cudaMemcpyDeviceToDevice, ctx->stream));
      GC_CUDA(cudaMemcpyAsync(lo.touched + lo.shift + lo.nc, hi.touched + hi.shift, (size_t)L, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return GECUT_OK;
  };
  GC_CHECK(exchange());
  int64_t ncut_all = 0;
  for (auto& a : S) ncut_all += a.ncut;
  bool all_aggregated = ncut_all == 0;
  for (int iter = 0; iter < 20 && !all_aggregated; ++iter) {  // max_iters (:153)
    for (auto& a : S) GC_CHECK(agg_sweep(ctx, a, side));
    uint32_t* h = (uint32_t*)ctx->h_pinned;
    for (int r = 0; r < n; ++r) GC_CUDA(cudaMemcpyAsync(h + r, S[r].flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    all_aggregated = true;
    for (int r = 0; r < n; ++r) all_aggregated = all_aggregated && h[r] == 0u;
    if (!all_aggregated) {
      for (auto& a : S) GC_CHECK(agg_touch(ctx, a));
      GC_CHECK(exchange());
    }
  }
  if (!all_aggregated) {
    ctx->err = "aggregate: not all cut cells could be aggregated in 20 sweeps (reference: @assert all_aggregated)";
    return GECUT_ERR_INVALID;
  }
  for (int r = 0; r < n; ++r) GC_CHECK(agg_copy_out(ctx, S[r], outs[r], on_device));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  return GECUT_OK;
}

extern "C" int gecut_aggregate(const gecut_result* cut, const gecut_facet_result* facets, const int32_t* program,
                               int len, const int32_t* facet_program, int facet_len, int side, double threshold,
                               int32_t* cell_to_cellin, int on_device) {
  if (!cut || !facets || !cell_to_cellin) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = cut->ctx;
  GC_CHECK(agg_check_args(ctx, side, threshold));
  const GcGrid& g = cut->grid;
  if (g.k0 != 0 || g.k1 != g.n[g.D - 1]) {
    ctx->err = "aggregate: this is a z-slab of the model (gecut_aggregate_slabs / gecut_comm_aggregate)";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  std::vector<AggSlab> S(1);
  int st = agg_setup(ctx, S[0], cut, facets, program, len, facet_program, facet_len, side, threshold,
                     on_device ? cell_to_cellin : nullptr);
  int32_t* outs[1] = {cell_to_cellin};
  if (st == GECUT_OK) st = agg_run_local(ctx, S, side, outs, on_device);
  agg_release(ctx, S[0]);
  return st;
}

extern "C" int gecut_aggregate_slabs(const gecut_result* const* cuts, const gecut_facet_result* const* facets, int nslabs,
                                     const int32_t* program, int len, const int32_t* facet_program, int facet_len, int side,
                                     double threshold, int32_t* const* cell_to_cellin, int on_device) {
  if (!cuts || !facets || !cell_to_cellin || nslabs < 1 || nslabs > 256) return GECUT_ERR_INVALID;
  for (int r = 0; r < nslabs; ++r)
    if (!cuts[r] || !facets[r] || !cell_to_cellin[r]) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = cuts[0]->ctx;
  GC_CHECK(agg_check_args(ctx, side, threshold));
  // the slabs must tile the layer axis of one model in ascending order
  const GcGrid& g0 = cuts[0]->grid;
  bool ok = g0.k0 == 0 && cuts[nslabs - 1]->grid.k1 == g0.n[g0.D - 1];
  for (int r = 0; r + 1 < nslabs && ok; ++r) {
    const GcGrid &a = cuts[r]->grid, &b = cuts[r + 1]->grid;
    ok = a.k1 == b.k0 && a.D == b.D && a.n[0] == b.n[0] && a.n[1] == b.n[1] && a.n[2] == b.n[2];
  }
  if (!ok) {
    ctx->err = "aggregate_slabs: the slabs must cover the layers of one model in ascending order";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  std::vector<AggSlab> S((size_t)nslabs);
  int st = GECUT_OK;
  for (int r = 0; r < nslabs && st == GECUT_OK; ++r)
    st = agg_setup(ctx, S[r], cuts[r], facets[r], program, len, facet_program, facet_len, side, threshold, nullptr);
  if (st == GECUT_OK) st = agg_run_local(ctx, S, side, cell_to_cellin, on_device);
  for (auto& a : S) agg_release(ctx, a);
  return st;
}

// one slab per rank: ghost layers by ncclSend / ncclRecv with the ranks below and above, termination by all-reduce
int gc_comm_exchange_layers(gecut_ctx* ctx, const void* send_lo, void* recv_lo, const void* send_hi, void* recv_hi, size_t bytes);
int gc_comm_allreduce_sum_u32(gecut_ctx* ctx, uint32_t* value_dev);

extern "C" int gecut_comm_aggregate(const gecut_result* cut, const gecut_facet_result* facets, const int32_t* program,
                                    int len, const int32_t* facet_program, int facet_len, int side, double threshold,
                                    int32_t* cell_to_cellin, int on_device) {
  if (!cut || !facets || !cell_to_cellin) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = cut->ctx;
  GC_CHECK(agg_check_args(ctx, side, threshold));
  if (!ctx->comm) {
    ctx->err = "no communicator on this context (gecut_comm_init)";
    return GECUT_ERR_INVALID;
  }
  const GcGrid& g = cut->grid;
  // rank r owns a slab above rank r-1's: the first rank starts at layer 0, the last one ends at the top
  if ((ctx->comm_rank == 0) != (g.k0 == 0) || (ctx->comm_rank == ctx->comm_world - 1) != (g.k1 == g.n[g.D - 1])) {
    ctx->err = "comm_aggregate: the ranks must own the z-slabs in ascending order";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  AggSlab a;
  int st = agg_setup(ctx, a, cut, facets, program, len, facet_program, facet_len, side, threshold, nullptr);
  auto exchange = [&]() -> int {
    const int64_t L = a.layer;
    GC_CHECK(gc_comm_exchange_layers(ctx, a.cellin + a.shift, a.cellin, a.cellin + a.shift + a.nc - L, a.cellin + a.shift + a.nc,
                                     sizeof(int32_t) * L));
    return gc_comm_exchange_layers(ctx, a.touched + a.shift, a.touched, a.touched + a.shift + a.nc - L,
                                   a.touched + a.shift + a.nc, (size_t)L);
  };
  // a failure before the collectives would leave the other ranks waiting: every rank goes through the same sequence and
  // the local status joins the all-reduced flag
  if (st == GECUT_OK) st = exchange();
  bool all_aggregated = false;
  for (int iter = 0; iter < 20 && !all_aggregated && st == GECUT_OK; ++iter) {
    st = agg_sweep(ctx, a, side);
    if (st == GECUT_OK) st = gc_comm_allreduce_sum_u32(ctx, a.flag);
    if (st != GECUT_OK) break;
    uint32_t* h = (uint32_t*)ctx->h_pinned;
    if (cudaMemcpyAsync(h, a.flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      ctx->err = "comm_aggregate: sweep failed";
      st = GECUT_ERR_CUDA;
      break;
    }
    all_aggregated = h[0] == 0u;
    if (!all_aggregated) {
      st = agg_touch(ctx, a);
      if (st == GECUT_OK) st = exchange();
    }
  }
  if (st == GECUT_OK && !all_aggregated) {
    ctx->err = "aggregate: not all cut cells could be aggregated in 20 sweeps (reference: @assert all_aggregated)";
    st = GECUT_ERR_INVALID;
  }
  if (st == GECUT_OK) st = agg_copy_out(ctx, a, cell_to_cellin, on_device);
  if (st == GECUT_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    ctx->err = "comm_aggregate failed";
    st = GECUT_ERR_CUDA;
  }
  agg_release(ctx, a);
  return st;
}

// ------------------------------------------------------------------------------------------------
// GhostSkeleton over z-slabs: the mask of the rank's LOCAL model (the facet numbering of gecut_cut_facets with the same
// slab).  Facets on the plane two slabs share are interior facets of the model: the state of the cell on the other side
// comes from the adjacent slab's boundary cell layer -- read in place when the slabs live in one context, received by
// ncclSend / ncclRecv otherwise.
// ------------------------------------------------------------------------------------------------
static int ghost_slab(gecut_ctx* ctx, const gecut_result* cut, const GcProgram& p, int side, const int8_t* lo_rows,
                      int64_t lo_pitch, int64_t lo_base, const int8_t* hi_rows, int64_t hi_pitch, int64_t hi_base,
                      int64_t* n_ghost, int8_t* mask_host) {
  const GcGrid& g = cut->grid;
  const int D = g.D;
  GcGrid gf = g;  // the local model of the slab
  gf.n[D - 1] = g.k1 - g.k0;
  gf.nn[D - 1] = gf.n[D - 1] + 1;
  GcSxFacets sx{};
  if (g.simplexified && !gc_build_sx_facets(D, &sx)) {
    ctx->err = "ghost skeleton: the simplexify table of the n-cube does not tile conformingly";
    return GECUT_ERR_NOTIMPLEMENTED;
  }
  const int64_t nf = g.simplexified ? (D == 2 ? sx_num_facets<2>(gf, sx) : sx_num_facets<3>(gf, sx))
                                    : (D == 2 ? num_facets<2>(gf) : num_facets<3>(gf));
  const int64_t ncut = cut->sizes.num_cutcells;
  GcScratch scratch(ctx);
  int8_t* mask = nullptr;
  unsigned long long* cnt = nullptr;
  GC_CHECK(scratch.alloc(&mask, (size_t)pad_pitch(nf)));
  GC_CHECK(scratch.alloc(&cnt, sizeof(unsigned long long)));
  GC_CUDA(cudaMemsetAsync(mask, 0, (size_t)nf, ctx->stream));
  GC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), ctx->stream));
  const int T = 256;
  if (ncut > 0) {
    if (g.simplexified && D == 2)
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask_sx<2><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(
          g, gf, sx, cut->lay, ncut, p, side, lo_rows, lo_pitch, lo_base, hi_rows, hi_pitch, hi_base, mask, nullptr));
    else if (g.simplexified)
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask_sx<3><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(
          g, gf, sx, cut->lay, ncut, p, side, lo_rows, lo_pitch, lo_base, hi_rows, hi_pitch, hi_base, mask, nullptr));
    else if (D == 2)
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask<2><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(
          g, gf, cut->lay, ncut, p, side, lo_rows, lo_pitch, lo_base, hi_rows, hi_pitch, hi_base, mask, nullptr));
    else
      GC_LAUNCH(ctx, "k_ghost_mask", k_ghost_mask<3><<<(unsigned)gc_div_up(ncut, T), T, 0, ctx->stream>>>(
          g, gf, cut->lay, ncut, p, side, lo_rows, lo_pitch, lo_base, hi_rows, hi_pitch, hi_base, mask, nullptr));
    GC_CHECK(gc_launch_check(ctx, "k_ghost_mask"));
  }
  for (int which = 0; which < 2; ++which) {
    const int8_t* rows = which == 0 ? lo_rows : hi_rows;
    if (!rows) continue;
    const unsigned nbp = (unsigned)gc_div_up(g.layer_cubes, T);
    if (g.simplexified && D == 2)
      GC_LAUNCH(ctx, "k_ghost_mask_plane", k_ghost_mask_plane_sx<2><<<nbp, T, 0, ctx->stream>>>(
          g, gf, sx, cut->lay, p, side, which, rows, which == 0 ? lo_pitch : hi_pitch, which == 0 ? lo_base : hi_base, mask));
    else if (g.simplexified)
      GC_LAUNCH(ctx, "k_ghost_mask_plane", k_ghost_mask_plane_sx<3><<<nbp, T, 0, ctx->stream>>>(
          g, gf, sx, cut->lay, p, side, which, rows, which == 0 ? lo_pitch : hi_pitch, which == 0 ? lo_base : hi_base, mask));
    else if (D == 2)
      GC_LAUNCH(ctx, "k_ghost_mask_plane", k_ghost_mask_plane<2><<<nbp, T, 0, ctx->stream>>>(
          g, gf, cut->lay, p, side, which, rows, which == 0 ? lo_pitch : hi_pitch, which == 0 ? lo_base : hi_base, mask));
    else
      GC_LAUNCH(ctx, "k_ghost_mask_plane", k_ghost_mask_plane<3><<<nbp, T, 0, ctx->stream>>>(
          g, gf, cut->lay, p, side, which, rows, which == 0 ? lo_pitch : hi_pitch, which == 0 ? lo_base : hi_base, mask));
    GC_CHECK(gc_launch_check(ctx, "k_ghost_mask_plane"));
  }
  {
    GC_LAUNCH(ctx, "k_count_mask", k_count_mask<<<(unsigned)std::min<int64_t>(1184, gc_div_up(nf, T)), T, 0, ctx->stream>>>(mask, nf, cnt));
    GC_CHECK(gc_launch_check(ctx, "k_count_mask"));
  }
  uint64_t* h = (uint64_t*)ctx->h_pinned;
  GC_CUDA(cudaMemcpyAsync(h, cnt, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (mask_host) GC_CUDA(cudaMemcpyAsync(mask_host, mask, (size_t)nf, cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  if (n_ghost) *n_ghost = (int64_t)h[0];
  return GECUT_OK;
}

static int ghost_check_args(gecut_ctx* ctx, const gecut_result* cut, int side) {
  if (side != GECUT_IN && side != GECUT_OUT && side != GECUT_CUT) {
    ctx->err = "ghost skeleton: in_or_out must be IN, OUT or CUT";
    return GECUT_ERR_INVALID;
  }
  (void)cut;
  return GECUT_OK;
}

extern "C" int gecut_ghost_skeleton_slabs(const gecut_result* const* cuts, int nslabs, const int32_t* program, int len, int side,
                                          int64_t* n_ghost, int8_t* const* facet_to_mask) {
  if (!cuts || nslabs < 1 || nslabs > 256) return GECUT_ERR_INVALID;
  for (int r = 0; r < nslabs; ++r)
    if (!cuts[r]) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = cuts[0]->ctx;
  const GcGrid& g0 = cuts[0]->grid;
  bool ok = g0.k0 == 0 && cuts[nslabs - 1]->grid.k1 == g0.n[g0.D - 1];
  for (int r = 0; r < nslabs && ok; ++r) ok = cuts[r]->ctx == ctx && cuts[r]->L == cuts[0]->L;
  for (int r = 0; r + 1 < nslabs && ok; ++r) ok = cuts[r]->grid.k1 == cuts[r + 1]->grid.k0;
  if (!ok) {
    ctx->err = "ghost_skeleton_slabs: the slabs must cover the layers of one model in ascending order, in one context";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  for (int r = 0; r < nslabs; ++r) {
    GC_CHECK(ghost_check_args(ctx, cuts[r], side));
    GcProgram p;
    GC_CHECK(gc_make_program(ctx, program, len, cuts[r]->L, &p));
    const gecut_result* lo = r > 0 ? cuts[r - 1] : nullptr;
    const gecut_result* hi = r + 1 < nslabs ? cuts[r + 1] : nullptr;
    const int64_t layer = layer_cells(g0);
    GC_CHECK(ghost_slab(ctx, cuts[r], p, side, lo ? lo->lay.bg_state : nullptr, lo ? lo->lay.bg_pitch : 0,
                        lo ? lo->sizes.num_bgcells - layer : 0, hi ? hi->lay.bg_state : nullptr, hi ? hi->lay.bg_pitch : 0, 0,
                        n_ghost ? n_ghost + r : nullptr, facet_to_mask ? facet_to_mask[r] : nullptr));
  }
  return GECUT_OK;
}

extern "C" int gecut_comm_ghost_skeleton(const gecut_result* cut, const int32_t* program, int len, int side, int64_t* n_ghost,
                                         int8_t* facet_to_mask) {
  if (!cut) return GECUT_ERR_INVALID;
  gecut_ctx* ctx = cut->ctx;
  GC_CHECK(ghost_check_args(ctx, cut, side));
  if (!ctx->comm) {
    ctx->err = "no communicator on this context (gecut_comm_init)";
    return GECUT_ERR_INVALID;
  }
  const GcGrid& g = cut->grid;
  if ((ctx->comm_rank == 0) != (g.k0 == 0) || (ctx->comm_rank == ctx->comm_world - 1) != (g.k1 == g.n[g.D - 1])) {
    ctx->err = "comm_ghost_skeleton: the ranks must own the z-slabs in ascending order";
    return GECUT_ERR_INVALID;
  }
  DeviceGuard guard(ctx->device);
  GcProgram p;
  GC_CHECK(gc_make_program(ctx, program, len, cut->L, &p));
  // the L state rows of the slab's first / last cell layer, packed [L][layer], to the ranks below / above
  const int L = cut->L;
  const int64_t layer = layer_cells(g), nc = cut->sizes.num_bgcells;
  GcScratch scratch(ctx);
  int8_t* buf = nullptr;  // [send lo][send hi][recv lo][recv hi], L * layer bytes each
  const size_t blk = (size_t)L * layer;
  GC_CHECK(scratch.alloc(&buf, 4 * blk));
  for (int ls = 0; ls < L; ++ls) {
    const int8_t* row = cut->lay.bg_state + (int64_t)ls * cut->lay.bg_pitch;
    GC_CUDA(cudaMemcpyAsync(buf + (size_t)ls * layer, row, (size_t)layer, cudaMemcpyDeviceToDevice, ctx->stream));
    GC_CUDA(cudaMemcpyAsync(buf + blk + (size_t)ls * layer, row + nc - layer, (size_t)layer, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  GC_CHECK(gc_comm_exchange_layers(ctx, buf, buf + 2 * blk, buf + blk, buf + 3 * blk, blk));
  const bool has_lo = ctx->comm_rank > 0, has_hi = ctx->comm_rank + 1 < ctx->comm_world;
  return ghost_slab(ctx, cut, p, side, has_lo ? buf + 2 * blk : nullptr, layer, 0, has_hi ? buf + 3 * blk : nullptr, layer, 0,
                    n_ghost, facet_to_mask);
}
