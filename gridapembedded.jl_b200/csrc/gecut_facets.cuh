// gecut_facets.cuh -- the facet grid of a Cartesian model in closed form + the objects behind the facet handles
// (shared by gecut_facets.cu and gecut_aggregate.cu; not part of the ABI).
#pragma once
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
  uint8_t* isb;       // Nsf: 1 when the bg facet of the sub-facet lies on the boundary of the model
};

// slot table of the facet grid of a simplexified model (explained at the end of this file)
struct GcSxFacets {
  int D, nt, nlf, nslots;
  int C0, cd;
  uint8_t vtx[24][3];    // cube-local vertex ids (bit d = offset along d) of the slot's facet nodes, facet node order
  int8_t face[24];       // -1: inside the cube; else (d << 1) | side
  uint8_t partner[24];   // the other slot with the same facet: same cube (inside), or the cube across the face
  uint8_t order[8][24];  // class -> owned slots, ascending
  uint8_t rank[8][24];   // class, slot -> position among the owned slots (0xff: met first elsewhere)
};

struct gecut_facet_result {
  gecut_ctx* ctx = nullptr;
  mutable int live_selections = 0;  // see gecut_result: a result outlives its selections
  bool free_requested = false;
  GcGrid grid{};
  int L = 0;
  GcLeaves leaves{};
  gecut_facet_sizes sizes{};
  GcFacetLayout lay{};
  int64_t tallies[GC_LMAX * 12] = {0};  // whole facets per level set: ((s*3 + d)*2 + isb), s = 0 IN / 1 OUT (k_facet_rows)
  bool owns_nodal[GC_LMAX] = {false};
  double* nodal_dev[GC_LMAX] = {nullptr};
  GcSxFacets sx{};  // valid when grid.simplexified
};

struct gecut_facet_selection {
  const gecut_facet_result* res = nullptr;
  int which = 0, kind = 0, side = 0;
  int64_t n_sub = 0, n_bg = 0;
  int32_t* sub_ids = nullptr;
  int32_t* bg_ids = nullptr;
  int64_t bg_dir_count[3] = {0, 0, 0};  // selected whole facets per normal direction (n-cube models)
  int64_t bg_slot_count[24] = {0};      // ... per slot of the cube (simplexified models)
  bool sub_lazy = false;                // n_sub / sub_measure known, the id list is written only on demand
  double sub_measure = 0.0;             // sum of the measures of the selected sub-facets
  GcProgram prog{};                     // for the lazy materialisation of bg_ids
};

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
__host__ __device__ inline FacetIdx facet_decode(const GcGrid& g, uint32_t f) {
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
__host__ __device__ inline void facet_node(const FacetIdx& fi, int m, int* nijk) {
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
__host__ __device__ inline bool facet_is_boundary(const GcGrid& g, const FacetIdx& fi) {
  return fi.side == 0 || fi.ijk[fi.d] == g.n[fi.d] - 1;
}

// first facet id (0-based) of the row of cells (j,k)
template <int D>
__host__ __device__ inline int64_t facet_row_base(const GcGrid& g, int j, int k) {
  int64_t b = 0;
  if (D == 3 && k > 0) b += facets_per_layer(g, 0) + (int64_t)(k - 1) * facets_per_layer(g, 1);
  if (j > 0) b += facets_per_row<D>(g, 0, k) + (int64_t)(j - 1) * facets_per_row<D>(g, 1, k);
  return b;
}


// id (0-based) of the facet of cell (i,j,k) normal to d on side `side` (0 low, 1 high): a low facet that is not on the
// model boundary is the high facet of the previous cell along d
template <int D>
__host__ __device__ inline int64_t facet_id_of(const GcGrid& g, int i, int j, int k, int d, int side) {
  int c[3] = {i, j, k};
  if (side == 0 && c[d] > 0) {
    c[d] -= 1;
    side = 1;
  }
  const int per = D + (c[1] == 0 ? 1 : 0) + ((D == 3 && c[2] == 0) ? 1 : 0);
  int64_t f = facet_row_base<D>(g, c[1], c[2]) + (int64_t)c[0] * per + (c[0] > 0 ? 1 : 0);
  for (int dd = D - 1; dd >= 0; --dd) {
    if (c[dd] == 0) {
      if (dd == d && side == 0) return f;
      f++;
    }
    if (dd == d) return f;  // side == 1
    f++;
  }
  return f;
}

// ------------------------------------------------------------------------------------------------
// Facet grid of a SIMPLEXIFIED Cartesian model (simplexify(CartesianDiscreteModel): cell = nt * cube + simplex, nodes of a
// simplex = cube nodes picked by simplexify(QUAD / HEX)).  Same first-encounter rule over (cell, local facet), local
// facets of TRI [1,2],[1,3],[2,3] and TET [1,2,3],[1,2,4],[1,3,4],[2,3,4]; a facet takes the node order of the local
// facet that meets it first.  A cube has nt * (D + 1) "slots" (simplex, local facet).  A slot's facet lies inside the
// cube (shared with a later or earlier simplex of the cube), on a high face (met here first: the neighbour across has
// a larger cube id) or on a low face (met first by the neighbour across, or here when the cube sits on the model
// boundary).  So the slots a cube OWNS depend only on which of its indices are 0 (a class of 2^D bits), the owned
// count is C0 + cd * (number of zero indices), and ids are closed-form again: rows of cubes are contiguous id ranges.
// The table is built on the host from the simplexify table (gc_build_sx_facets) and passed to kernels by value.
// ------------------------------------------------------------------------------------------------

// host: fills the table for D = 2, 3; false when the simplexify table does not tile conformingly
bool gc_build_sx_facets(int D, GcSxFacets* t);

struct SxFacetIdx {
  int ijk[3];
  int slot;
};

template <int D>
__host__ __device__ inline int sx_class(int i, int j, int k) {
  return (i == 0 ? 1 : 0) | (j == 0 ? 2 : 0) | ((D == 3 && k == 0) ? 4 : 0);
}
template <int D>
__host__ __device__ inline int64_t sx_facets_per_row(const GcGrid& g, const GcSxFacets& t, int j, int k) {
  return (int64_t)g.n[0] * (t.C0 + (j == 0 ? t.cd : 0) + ((D == 3 && k == 0) ? t.cd : 0)) + t.cd;
}
__host__ __device__ inline int64_t sx_facets_per_layer(const GcGrid& g, const GcSxFacets& t, int k) {
  return sx_facets_per_row<3>(g, t, 0, k) + (int64_t)(g.n[1] - 1) * sx_facets_per_row<3>(g, t, 1, k);
}
template <int D>
__host__ __device__ inline int64_t sx_num_facets(const GcGrid& g, const GcSxFacets& t) {
  if (D == 2) return sx_facets_per_row<2>(g, t, 0, 0) + (int64_t)(g.n[1] - 1) * sx_facets_per_row<2>(g, t, 1, 0);
  return sx_facets_per_layer(g, t, 0) + (int64_t)(g.n[2] - 1) * sx_facets_per_layer(g, t, 1);
}
template <int D>
__host__ __device__ inline int64_t sx_facet_row_base(const GcGrid& g, const GcSxFacets& t, int j, int k) {
  int64_t b = 0;
  if (D == 3 && k > 0) b += sx_facets_per_layer(g, t, 0) + (int64_t)(k - 1) * sx_facets_per_layer(g, t, 1);
  if (j > 0) b += sx_facets_per_row<D>(g, t, 0, k) + (int64_t)(j - 1) * sx_facets_per_row<D>(g, t, 1, k);
  return b;
}

template <int D>
__host__ __device__ inline SxFacetIdx sx_facet_decode(const GcGrid& g, const GcSxFacets& t, uint32_t f) {
  SxFacetIdx r;
  uint32_t rem = f;
  int k = 0;
  if (D == 3) {
    const uint32_t L0 = (uint32_t)sx_facets_per_layer(g, t, 0);
    if (rem >= L0) {
      const uint32_t L1 = (uint32_t)sx_facets_per_layer(g, t, 1);
      rem -= L0;
      k = 1 + (int)(rem / L1);
      rem = rem % L1;
    }
  }
  int j = 0;
  const uint32_t R0 = (uint32_t)sx_facets_per_row<D>(g, t, 0, k);
  if (rem >= R0) {
    const uint32_t R1 = (uint32_t)sx_facets_per_row<D>(g, t, 1, k);
    rem -= R0;
    j = 1 + (int)(rem / R1);
    rem = rem % R1;
  }
  const uint32_t per = (uint32_t)(t.C0 + (j == 0 ? t.cd : 0) + ((D == 3 && k == 0) ? t.cd : 0));
  int i = 0;
  if (rem >= per + (uint32_t)t.cd) {
    rem -= per + (uint32_t)t.cd;
    i = 1 + (int)(rem / per);
    rem = rem % per;
  }
  r.ijk[0] = i;
  r.ijk[1] = j;
  r.ijk[2] = k;
  r.slot = t.order[sx_class<D>(i, j, k)][rem];
  return r;
}

// id (0-based) of the facet of slot `slot` of cube (i,j,k), whoever met it first
template <int D>
__host__ __device__ inline int64_t sx_facet_id_of(const GcGrid& g, const GcSxFacets& t, int i, int j, int k, int slot) {
  int c[3] = {i, j, k};
  int cls = sx_class<D>(c[0], c[1], c[2]);
  if (t.rank[cls][slot] == 0xff) {
    const int fc = t.face[slot];
    if (fc < 0) {
      slot = t.partner[slot];  // an earlier simplex of the same cube
    } else {
      c[fc >> 1] -= 1;  // low face with a neighbour across: its high-face slot
      slot = t.partner[slot];
      cls = sx_class<D>(c[0], c[1], c[2]);
    }
  }
  const int per = t.C0 + (c[1] == 0 ? t.cd : 0) + ((D == 3 && c[2] == 0) ? t.cd : 0);
  return sx_facet_row_base<D>(g, t, c[1], c[2]) + (int64_t)c[0] * per + (c[0] > 0 ? t.cd : 0) + t.rank[cls][slot];
}

template <int D>
__host__ __device__ inline void sx_facet_node(const SxFacetIdx& fi, const GcSxFacets& t, int m, int* nijk) {
  const int lv = t.vtx[fi.slot][m];
  nijk[0] = fi.ijk[0] + (lv & 1);
  nijk[1] = fi.ijk[1] + ((lv >> 1) & 1);
  nijk[2] = D == 3 ? fi.ijk[2] + ((lv >> 2) & 1) : 0;
}

template <int D>
__host__ __device__ inline bool sx_facet_is_boundary(const GcGrid& g, const SxFacetIdx& fi, const GcSxFacets& t) {
  const int fc = t.face[fi.slot];
  if (fc < 0) return false;
  const int d = fc >> 1;
  return (fc & 1) == 0 ? fi.ijk[d] == 0 : fi.ijk[d] == g.n[d] - 1;
}
