// gecut_cutcells.cu -- K3: marching-simplex subtriangulation of the cut background cells.
//
// Reference path replaced (src/LevelSetCutters/CutTriangulations.jl):
//   _simplexify_and_isolate_cells_in_cutgrid / _simplexify   :505-617   stage-0 mesh of the cut cells
//   cut_sub_triangulation_with_boundary_several_levelsets    :309-343   for ls = L..1: count -> allocate -> fill
//   _cut_cell! / set_point_data!                             :179-222, 123-135  connectivity, vertex copies, edge lerps
//   _cut_cell_boundary! / _setup_normal                      :268-292, LookupTables.jl:304-347  facets + normals
//   cut_sub_triangulation_several_levelsets (facet re-cuts)  :294-307, 405-441
//   merge_sub_face_data                                      src/Interfaces/SubFacetTriangulations.jl:161-208
//
// B200 design: the reference re-emits the WHOLE sub-mesh once per level set (O(L * Nsub) traffic and 2L passes).
// Every stage keeps the children of a cell contiguous and in order, so the final arrays are exactly the depth-first
// traversal of each stage-0 simplex' refinement tree; the facets of level set ls are the depth-first traversal of the
// facet re-cut trees hanging off depth L-ls.  Counting pass (reference: count_sub_triangulation_with_boundary :238-251),
// exclusive scan (stable: preserves the reference ordering), filling pass that writes the final layout only.  Node
// values are recomputed from the leaf records (bit-identical to the classification pass) or gathered from NODAL vectors.
//
//   L == 1 : k_cut1_count(_sx) -> scan over tiles of 32 simplices -> k_cut1_fill   (no tree: case -> table row)
//   L  > 1 : k_cut_fast<count> (+ k_cut_walk<count> for the recorded simplices) -> scan over simplices ->
//            k_cut_fast<fill> (+ k_cut_walk<fill>).  k_cut_fast handles the simplices cut by at most one level set
//            without a tree; k_cut_walk is the generic depth-first walk, restricted to the level sets with mixed signs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gecut_internal.cuh"

namespace {

template <int D, int LM>
struct Pt {
  double x[D];
  double r[D];
  double w[LM];
};

template <int D>
struct WalkCfg {
  static constexpr int NPC = (D == 3) ? 8 : 5;  // max points of a cut D-simplex (vertices + cut edges)
  static constexpr int NPF = (D == 3) ? 5 : 3;  // max points of a cut (D-1)-simplex
};

template <int D>
__device__ __forceinline__ void unit_normal(const double* p1, const double* p2, const double* p3, double orient, double* n) {
  // _setup_normal / _normal_vector / _orthogonal_vector (LookupTables.jl:309-360), then orientation*normal (:287-288)
  if (D == 2) {
    double v1 = __dsub_rn(p2[0], p1[0]), v2 = __dsub_rn(p2[1], p1[1]);
    double w1 = v2, w2 = -v1;
    double m = __dsqrt_rn(__dadd_rn(__dmul_rn(w1, w1), __dmul_rn(w2, w2)));
    if (m < 2.220446049250313e-16) {
      n[0] = __dmul_rn(orient, 0.0);
      n[1] = __dmul_rn(orient, 0.0);
    } else {
      n[0] = __dmul_rn(orient, __ddiv_rn(w1, m));
      n[1] = __dmul_rn(orient, __ddiv_rn(w2, m));
    }
  } else {
    double a0 = __dsub_rn(p2[0], p1[0]), a1 = __dsub_rn(p2[1], p1[1]), a2 = __dsub_rn(p2[2], p1[2]);
    double b0 = __dsub_rn(p3[0], p1[0]), b1 = __dsub_rn(p3[1], p1[1]), b2 = __dsub_rn(p3[2], p1[2]);
    double w1 = __dsub_rn(__dmul_rn(a1, b2), __dmul_rn(a2, b1));
    double w2 = __dsub_rn(__dmul_rn(a2, b0), __dmul_rn(a0, b2));
    double w3 = __dsub_rn(__dmul_rn(a0, b1), __dmul_rn(a1, b0));
    double m = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(w1, w1), __dmul_rn(w2, w2)), __dmul_rn(w3, w3)));
    if (m < 2.220446049250313e-16) {
      n[0] = __dmul_rn(orient, 0.0);
      n[1] = __dmul_rn(orient, 0.0);
      n[2] = __dmul_rn(orient, 0.0);
    } else {
      n[0] = __dmul_rn(orient, __ddiv_rn(w1, m));
      n[1] = __dmul_rn(orient, __ddiv_rn(w2, m));
      n[2] = __dmul_rn(orient, __ddiv_rn(w3, m));
    }
  }
}

// new point on the edge (a,b) cut by the level set in value slot `sl` (_cut_cell!, CutTriangulations.jl:206-218).
// Points carry the values of the nw level sets that can still be looked at further down the tree, in compact slots.
// WITH_XR = false (counting walk): only the level-set values are interpolated -- the tree shape depends on nothing else
template <int D, int NW, bool WITH_XR = true>
__device__ __forceinline__ void cut_point(const Pt<D, NW>& a, const Pt<D, NW>& b, int sl, int nw, Pt<D, NW>& o) {
  double w1 = fabs(a.w[sl]), w2 = fabs(b.w[sl]);
  double coeff = __ddiv_rn(w1, __dadd_rn(w1, w2));
  if (WITH_XR) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      o.x[d] = gc_lerp(a.x[d], b.x[d], coeff);
      o.r[d] = gc_lerp(a.r[d], b.r[d], coeff);
    }
  }
  for (int j = 0; j < nw; ++j) o.w[j] = gc_lerp(a.w[j], b.w[j], coeff);
}

// Stage-0 simplex `t` of the cut bg cell `cell0` (0-based, slab-local): points with coordinates and reference coordinates
// of the bg polytope; first two vertices swapped when det J < 0 in physical space (_simplexify :562-617,
// _ensure_positive_jacobians! :532-533 -> LookupTables.jl:199-216).  Every level set is evaluated at the D+1 vertices:
// *mixed / *outm get one bit per level set (mixed signs / all OUT); the values of the level sets in
// wmask = mixed | {0, 1} are kept, in ascending level order, in the compact slots w[0..nw).
template <int D, int NW>
__device__ __forceinline__ void setup_stage0(const GcGrid& g, const GcLeaves& lv, int64_t cell0, int t,
                                             const uint8_t* simp_raw, const uint8_t* simp_pos,
                                             const int8_t* __restrict__ bg_state, int64_t bg_pitch, Pt<D, NW>* S,
                                             uint32_t* mixed_out, uint32_t* outm_out) {
  const int L = lv.L;
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
  int64_t node[D + 1];
#pragma unroll
  for (int m = 0; m <= D; ++m) {
    int lvid;
    if (g.simplexified) {
      lvid = simp_raw[(D - 2) * 24 + typ * 4 + m];
    } else {
      lvid = simp_pos[(D - 2) * 24 + t * 4 + m];
    }
    int nijk[3] = {ijk[0] + (lvid & 1), ijk[1] + ((lvid >> 1) & 1), ijk[2] + ((lvid >> 2) & 1)};
    double x[3] = {gc_coord(g, 0, nijk[0]), gc_coord(g, 1, nijk[1]), D == 3 ? gc_coord(g, 2, nijk[2]) : 0.0};
#pragma unroll
    for (int d = 0; d < D; ++d) {
      S[m].x[d] = x[d];
      if (g.simplexified)
        S[m].r[d] = (m == d + 1) ? 1.0 : 0.0;  // simplex vertices: origin then unit axes
      else
        S[m].r[d] = (double)((lvid >> d) & 1);  // n-cube vertices, x fastest
    }
    node[m] = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                       : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
  }
  uint32_t mixed = 0u, outm = 0u;
  int slot = 0;
  for (int ls = 0; ls < L; ++ls) {
    double val[D + 1];
    bool any = false, all = true;
    // a level set that does not cut the bg cell has one sign on all its vertices and only that sign is ever looked at
    // (its stage never produces a cut point): no evaluation
    const int bgst = bg_state[(int64_t)ls * bg_pitch + cell0];
#pragma unroll
    for (int m = 0; m <= D; ++m) {
      if (bgst != GC_CUT) {
        val[m] = bgst == GC_OUT ? 1.0 : -1.0;
      } else if (lv.leaf[ls].kind == GECUT_LEAF_NODAL) {
        val[m] = __ldg(lv.nodal[ls] + (node[m] - lv.nodal_base));
      } else {
        const double x[3] = {S[m].x[0], S[m].x[1], D == 3 ? S[m].x[2] : 0.0};
        val[m] = gc_eval_leaf(lv.leaf[ls], x, D);
      }
      const bool o = val[m] > 0.0;
      any |= o;
      all &= o;
    }
    if (any && !all) mixed |= 1u << ls;
    if (all) outm |= 1u << ls;
    if (ls < 2 || (any && !all)) {
      if (slot < NW) {
#pragma unroll
        for (int m = 0; m <= D; ++m) S[m].w[slot] = val[m];
      }
      slot++;
    }
  }
  *mixed_out = mixed;
  *outm_out = outm;
  // J[a][b] = x_{a+1}[b] - x_0[b]; flip iff det(J) < 0
  double J[3][3];
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = 0; b < D; ++b) J[a][b] = __dsub_rn(S[a + 1].x[b], S[0].x[b]);
  double dV;
  if (D == 2) {
    dV = __dsub_rn(__dmul_rn(J[0][0], J[1][1]), __dmul_rn(J[0][1], J[1][0]));
  } else {
    double p1 = __dmul_rn(__dmul_rn(J[0][0], J[1][1]), J[2][2]);
    double p2 = __dmul_rn(__dmul_rn(J[0][1], J[1][2]), J[2][0]);
    double p3 = __dmul_rn(__dmul_rn(J[0][2], J[1][0]), J[2][1]);
    double q1 = __dmul_rn(__dmul_rn(J[0][0], J[1][2]), J[2][1]);
    double q2 = __dmul_rn(__dmul_rn(J[0][1], J[1][0]), J[2][2]);
    double q3 = __dmul_rn(__dmul_rn(J[0][2], J[1][1]), J[2][0]);
    dV = __dsub_rn(__dadd_rn(__dadd_rn(p1, p2), p3), __dadd_rn(__dadd_rn(q1, q2), q3));
  }
  if (dV < 0.0) {
    Pt<D, NW> tmp = S[0];
    S[0] = S[1];
    S[1] = tmp;
  }
}

// one row of cell_to_points: a single 16-byte store for tetrahedra
template <int D>
__device__ __forceinline__ void store_c2p_row(int32_t* __restrict__ c2p, uint32_t cid, const int32_t* row) {
  if (D == 3) {
    reinterpret_cast<int4*>(c2p)[cid] = make_int4(row[0], row[1], row[2], row[3]);
  } else {
#pragma unroll
    for (int i = 0; i <= D; ++i) c2p[(int64_t)cid * (D + 1) + i] = row[i];
  }
}

// Generic walk (any L): index-based depth-first traversal.  All points a thread ever needs live in one per-thread pool
// (stage-0 vertices, then the cut points of the levels on the current path, then the cut points of the facet being
// re-cut); frames only hold small indices into the pool, so descending a level copies D+1 bytes, not D+1 points, and a
// level that does not cut the simplex costs one case evaluation and an index permutation.  The pool shrinks back when
// the walk returns from a level.
constexpr int WALK_T = 64;  // threads per block of the generic walk
// tree-walk launch buckets: level sets tracked per recorded simplex (mixed signs + the emission stages 0 and 1)
constexpr int WALK_NB = 5;
__host__ __device__ constexpr int walk_nw(int b) { return b == 0 ? 4 : b == 1 ? 6 : b == 2 ? 8 : b == 3 ? 12 : GC_LMAX; }

// Work split: a recorded simplex is walked by WALK_ITEMS = 8 lanes.  Every lane enters the top visited stage (case, cut
// points: cheap and redundant), then lane j < 6 walks only the subtree of top-stage child j and lanes 6, 7 only the
// re-cut trees of the top stage's facets 0, 1.  The subtrees write disjoint, consecutive ranges: children in order for
// the cell / point arrays and the facet arrays of the lower level sets, the top-stage facets alone own the facet arrays
// of the top level set.  walk_item<FILL=false> returns the item's counters; the count kernel adds them up per simplex,
// the fill kernel turns them into the item's start offsets with an 8-lane exclusive scan and walks again, writing.
constexpr int WALK_ITEMS = 8;

template <int D, int NW, bool FILL>
__device__ __forceinline__ void walk_item(const GcGrid& g, const GcLeaves& lv, const GcSimplexTable& TC,
                                          const GcSimplexTable& TF, const uint8_t* __restrict__ simp_raw,
                                          const uint8_t* __restrict__ simp_pos, const GcLayout& lay, int64_t s,
                                          int32_t bgcell1, int item, uint32_t& cell_off, uint32_t& pt_off,
                                          uint32_t* fac_off, uint32_t* fpt_off) {
  constexpr int NPC = WalkCfg<D>::NPC, NPF = WalkCfg<D>::NPF;
  constexpr int CE = NPC - (D + 1), FE = NPF - D;  // max new points per cell / facet cut
  // NW bounds the level sets a recorded simplex tracks (mixed signs + the two emission stages 0 and 1), hence the depth
  // of both walks and the size of the per-thread point pool; level-set IDS still run up to GC_LMAX.
  constexpr int LM = GC_LMAX;
  constexpr int NPOOL = (D + 1) + CE * NW + FE * NW;
  constexpr int LF = NW;
  const int L = lv.L;
  const int t = (int)(s % g.nt0);

  Pt<D, NW> pool[NPOOL];
  uint8_t v[NW][D + 1], e[NW][NPC];      // simplex vertices / expanded point set of every visited stage on the path
  uint8_t fv[LF][D], fe[LF][NPF];         // the same for the facet being re-cut
  int8_t child[NW], cs[NW], npE[NW], st[LM], ptop_at[NW];
  int8_t fchild[LF], fcs[LF], fptop_at[LF], fst[LM];

  uint32_t mixed = 0u, outm = 0u;
  setup_stage0<D, NW>(g, lv, bgcell1 - 1, t, simp_raw, simp_pos, lay.bg_state, lay.bg_pitch, pool, &mixed, &outm);
  int ptop = D + 1;
#pragma unroll
  for (int i = 0; i <= D; ++i) v[0][i] = (uint8_t)i;

  // Level compression.  A level set whose values have one sign on the D+1 vertices of the stage-0 simplex has that sign
  // on every point of the refinement tree (cut points are convex combinations), so its stage never cuts anything: every
  // simplex goes through it as the single child of the all-IN / all-OUT table row, i.e. with its vertices permuted by
  // that row and a constant state.  The walk therefore visits only the level sets with mixed signs (plus the last one
  // of each sequence, where the output is emitted) and applies the permutations of the skipped stages in between.
  int8_t lsd[NW], fk[LF];
  for (int k = 0; k < L; ++k) {
    st[k] = ((outm >> k) & 1u) ? (int8_t)GC_OUT : (int8_t)GC_IN;  // overwritten on the way down for the visited level sets
    fst[k] = st[k];
  }
  const uint32_t wmask = (mixed | 3u) & ((1u << L) - 1u);  // tracked level sets: mixed signs + the emission stages 0 and 1
  const int nw = __popc(wmask);
  if (nw > NW) return;  // cannot happen: the launch bucket is chosen from the same count (k_cut_fast)
  auto ws = [&](int k) -> int { return __popc(wmask & ((1u << k) - 1u)); };  // value slot of level set k
  const int c_allout = (1 << (D + 1)) - 1, fc_allout = (1 << D) - 1;
  // skipped cell stages hi-1 ... lo+1 applied to a vertex list
  auto skip_cell_stages = [&](uint8_t* vl, int hi, int lo) {
    for (int k = hi - 1; k > lo; --k) {
      const uint8_t* pm = TC.sub_pts[((outm >> k) & 1u) ? c_allout : 0][0];
      uint8_t tmp[D + 1];
#pragma unroll
      for (int i = 0; i <= D; ++i) tmp[i] = vl[pm[i]];
#pragma unroll
      for (int i = 0; i <= D; ++i) vl[i] = tmp[i];
    }
  };
  // the same for the facet re-cut stages (level set `skip` is not part of that sequence)
  auto skip_facet_stages = [&](uint8_t* vl, int hi, int lo, int skip) {
    for (int k = hi - 1; k > lo; --k) {
      if (k == skip) continue;
      const uint8_t* pm = TF.sub_pts[((outm >> k) & 1u) ? fc_allout : 0][0];
      uint8_t tmp[D];
#pragma unroll
      for (int i = 0; i < D; ++i) tmp[i] = vl[pm[i]];
#pragma unroll
      for (int i = 0; i < D; ++i) vl[i] = tmp[i];
    }
  };
  const uint32_t cmask = (mixed | 1u) & ((1u << L) - 1u);  // visited cell stages
  lsd[0] = (int8_t)(31 - __clz(cmask));
  skip_cell_stages(v[0], L, lsd[0]);

  int depth = 0;
  child[0] = -1;
  while (depth >= 0) {
    const int ls = lsd[depth];
    const int wls = ws(ls);
    if (child[depth] < 0) {
      // ---- enter: case, cut points (CutTriangulations.jl:179-222)
      int c = 0;
#pragma unroll
      for (int i = 0; i <= D; ++i) {
        e[depth][i] = v[depth][i];
        if (pool[v[depth][i]].w[wls] > 0.0) c |= 1 << i;
      }
      cs[depth] = (int8_t)c;
      ptop_at[depth] = (int8_t)ptop;
      int np = D + 1;
      if (TC.inoutcut[c] == GC_CUT) {
        for (int ed = 0; ed < TC.nedges; ++ed) {
          const int a = e[depth][TC.edges[ed][0]], b = e[depth][TC.edges[ed][1]];
          if ((pool[a].w[wls] > 0.0) != (pool[b].w[wls] > 0.0)) {
            cut_point<D, NW, FILL>(pool[a], pool[b], wls, nw, pool[ptop]);
            e[depth][np++] = (uint8_t)ptop++;
          }
        }
      }
      npE[depth] = (int8_t)np;
      const int nfac = TC.nfac[c];
      // ---- boundary facets of this stage (CutTriangulations.jl:268-292); at the top stage only by their own lanes
      {
        for (int f = 0; f < nfac; ++f) {
          if (depth == 0 && item != 6 + f) continue;
          const uint8_t* fp = TC.fac_pts[c][f];
          double nrm[3] = {0.0, 0.0, 0.0};
          if (FILL)
            unit_normal<D>(pool[e[depth][fp[0]]].x, pool[e[depth][fp[1]]].x, pool[e[depth][fp[D - 1]]].x,
                           (double)TC.fac_orient[c][f], nrm);
#pragma unroll
          for (int i = 0; i < D; ++i) fv[0][i] = e[depth][fp[i]];
          // ---- re-cut the facet by every other level set, last to first (CutTriangulations.jl:294-307,324-327);
          //      visited: the other level sets with mixed signs + the last one of the sequence (emission)
          const int ptop_cell = ptop;
          const int klast = ls == 0 ? 1 : 0;
          const uint32_t fmask = ((mixed & ~(1u << ls)) | (1u << klast)) & ((1u << L) - 1u);
          fk[0] = (int8_t)(31 - __clz(fmask));
          skip_facet_stages(fv[0], L, fk[0], ls);
          int fdepth = 0;
          fchild[0] = -1;
          while (fdepth >= 0) {
            const int k = fk[fdepth];
            const int wk = ws(k);
            if (fchild[fdepth] < 0) {
              int fc = 0;
#pragma unroll
              for (int i = 0; i < D; ++i) {
                fe[fdepth][i] = fv[fdepth][i];
                if (pool[fv[fdepth][i]].w[wk] > 0.0) fc |= 1 << i;
              }
              fcs[fdepth] = (int8_t)fc;
              fptop_at[fdepth] = (int8_t)ptop;
              int fn = D;
              if (TF.inoutcut[fc] == GC_CUT) {
                for (int ed = 0; ed < TF.nedges; ++ed) {
                  const int a = fe[fdepth][TF.edges[ed][0]], b = fe[fdepth][TF.edges[ed][1]];
                  if ((pool[a].w[wk] > 0.0) != (pool[b].w[wk] > 0.0)) {
                    cut_point<D, NW, FILL>(pool[a], pool[b], wk, nw, pool[ptop]);
                    fe[fdepth][fn++] = (uint8_t)ptop++;
                  }
                }
              }
              if (k == klast) {
                // final facets of level set ls: point block + children
                const int nsf = TF.nsub[fc];
                if (FILL) {
                  const uint32_t pb = fpt_off[ls];
                  for (int p = 0; p < fn; ++p)
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                      lay.sf_X[(int64_t)(pb + p) * D + d] = pool[fe[fdepth][p]].x[d];
                      lay.sf_R[(int64_t)(pb + p) * D + d] = pool[fe[fdepth][p]].r[d];
                    }
                  for (int j = 0; j < nsf; ++j) {
                    const uint32_t fid = fac_off[ls] + j;
                    double pf[D][3];
#pragma unroll
                    for (int i = 0; i < D; ++i) {
                      lay.sf_c2p[(int64_t)fid * D + i] = (int32_t)(pb + TF.sub_pts[fc][j][i] + 1);
                      lay.sf_N[(int64_t)fid * D + i] = nrm[i];
#pragma unroll
                      for (int d = 0; d < D; ++d) pf[i][d] = pool[fe[fdepth][TF.sub_pts[fc][j][i]]].x[d];
                    }
                    lay.sf_c2bg[fid] = bgcell1;
                    lay.sf_meas[fid] = integrate_one<D - 1>(simplex_meas<D, D - 1>(pf));
                    for (int kk = 0; kk < L; ++kk) {
                      int8_t val = (kk == ls) ? (int8_t)GECUT_INTERFACE : ((kk == k) ? TF.sub_inout[fc][j] : fst[kk]);
                      lay.sf_state[(int64_t)kk * lay.sf_pitch + fid] = val;
                    }
                  }
                }
                fac_off[ls] += nsf;
                fpt_off[ls] += fn;
                ptop = fptop_at[fdepth];
                fdepth--;
                continue;
              }
              fchild[fdepth] = 0;
            }
            const int fc = fcs[fdepth];
            if (fchild[fdepth] < TF.nsub[fc]) {
              const int j = fchild[fdepth]++;
              fst[k] = TF.sub_inout[fc][j];
#pragma unroll
              for (int i = 0; i < D; ++i) fv[fdepth + 1][i] = fe[fdepth][TF.sub_pts[fc][j][i]];
              fk[fdepth + 1] = (int8_t)(31 - __clz(fmask & ((1u << k) - 1u)));
              skip_facet_stages(fv[fdepth + 1], k, fk[fdepth + 1], ls);
              fchild[fdepth + 1] = -1;
              fdepth++;
            } else {
              ptop = fptop_at[fdepth];
              fdepth--;
            }
          }
          ptop = ptop_cell;
        }
      }
      if (ls == 0) {
        // ---- final subcells of this branch: point block + children (CutTriangulations.jl:186-203)
        const int nsub = TC.nsub[c];
        if (FILL) {
          for (int p = 0; p < np; ++p)
#pragma unroll
            for (int d = 0; d < D; ++d) {
              lay.sc_X[(int64_t)(pt_off + p) * D + d] = pool[e[depth][p]].x[d];
              lay.sc_R[(int64_t)(pt_off + p) * D + d] = pool[e[depth][p]].r[d];
            }
          for (int j = 0; j < nsub; ++j) {
            const uint32_t cid = cell_off + j;
            double pc[D + 1][3];
#pragma unroll
            for (int i = 0; i <= D; ++i) {
              lay.sc_c2p[(int64_t)cid * (D + 1) + i] = (int32_t)(pt_off + TC.sub_pts[c][j][i] + 1);
#pragma unroll
              for (int d = 0; d < D; ++d) pc[i][d] = pool[e[depth][TC.sub_pts[c][j][i]]].x[d];
            }
            lay.sc_c2bg[cid] = bgcell1;
            lay.sc_meas[cid] = integrate_one<D>(simplex_meas<D, D>(pc));
            lay.sc_state[cid] = TC.sub_inout[c][j];
            for (int kk = 1; kk < L; ++kk) lay.sc_state[(int64_t)kk * lay.sc_pitch + cid] = st[kk];
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
    if (depth == 0) {
      // top stage: this lane owns child `item` only
      if (child[0] != 0 || item >= TC.nsub[c]) break;
      child[0] = (int8_t)item;
    }
    if (child[depth] < TC.nsub[c]) {
      const int j = child[depth]++;
      st[ls] = TC.sub_inout[c][j];
#pragma unroll
      for (int i = 0; i <= D; ++i) v[depth + 1][i] = e[depth][TC.sub_pts[c][j][i]];
      lsd[depth + 1] = (int8_t)(31 - __clz(cmask & ((1u << ls) - 1u)));
      skip_cell_stages(v[depth + 1], ls, lsd[depth + 1]);
      child[depth + 1] = -1;
      depth++;
    } else {
      ptop = ptop_at[depth];
      depth--;
    }
  }

}

// ---- bookkeeping of the multi-level-set cut (L > 1), tile design of the single-level-set path -------------------------
// A warp owns a TILE of 32 consecutive stage-0 simplices.  Channels: 0 subcells, 1 subcell points, 2+k subfacets of level
// set k, 2+L+k subfacet points of level set k.  The count pass leaves one TOTAL per tile and channel (not per simplex),
// an exclusive scan over the tiles turns them into tile bases, and the fill pass rebuilds the offsets of the 32 simplices
// of a tile with warp scans.  Simplices cut by two or more level sets ("recorded", walked by k_cut_walk) get a slot:
// their counters live in slot-indexed arrays, are added to the tile totals with atomics (integer sums: order-free) and are
// replaced by their start offsets by the fill kernel of the tile before the walk writes.
constexpr uint32_t CM_SLOW = 0x80000000u;  // sinfo: recorded simplex, low bits = slot
constexpr int CM_NBK = WALK_NB + 1;          // launch buckets of the recorded simplices
struct GcMulti {
  int nch;                // 2 + 2L
  int64_t ntiles;
  uint32_t* tile;         // [nch][ntiles] totals (count pass) -> exclusive bases (scan)
  uint32_t* sinfo;        // [nsimp] fast: outm | a << 16 | case << 20 | mixed << 24; recorded: CM_SLOW | slot
  uint32_t* slow_simp;    // [slot] -> stage-0 simplex
  uint8_t* slow_bucket;   // [slot] -> launch bucket: 0 = exactly two mixed level sets (k_cut_two), 1 + b = generic walk
  uint32_t* slow_mask;    // [slot] -> mixed-sign level sets | all-OUT level sets << 16
  uint32_t* slow_val;     // [nch][slow_n] counters (walk count) -> start offsets (fill of the tile)
  int64_t slow_n;
  uint32_t* slow_count;   // [CM_NBK] recorded simplices per bucket, [CM_NBK] their total (slot allocator)
  unsigned long long* two_items;  // [bucket-0 slots x WALK_ITEMS] per-item counters of k_cut_two (count -> fill)
  uint32_t* walk_items;   // [nch][walk_items_n] per-item counters of k_cut_walk (count -> fill), items of all walk buckets
  int64_t walk_items_n;
  uint32_t fac_base[GC_LMAX];  // global offset of the facets of level set ls (merge_sub_face_data order)
  uint32_t fpt_base[GC_LMAX];
};

template <int D, int NW, bool FILL>
__global__ void __launch_bounds__(WALK_T)
k_cut_walk(const GcGrid g, const GcLeaves lv, const GcSimplexTable* __restrict__ tables,
           const uint8_t* __restrict__ simp_raw, const uint8_t* __restrict__ simp_pos,
           const int32_t* __restrict__ cutcells, GcMulti mu, GcLayout lay,
           const uint32_t* __restrict__ slots, int64_t nslots, int64_t items_base) {
  // WALK_ITEMS lanes per RECORDED simplex (those cut by two or more level sets, listed by k_cutm_count): the tree walk is
  // an order of magnitude longer than the no-tree path, so it runs in its own dense launches.
  constexpr int LM = GC_LMAX;
  __shared__ GcSimplexTable s_tab[2];  // [0]: D-simplex (cells), [1]: (D-1)-simplex (facet re-cuts)
  {
    const uint32_t* src0 = reinterpret_cast<const uint32_t*>(tables + (D - 1));
    const uint32_t* src1 = reinterpret_cast<const uint32_t*>(tables + (D - 2));
    uint32_t* dst0 = reinterpret_cast<uint32_t*>(&s_tab[0]);
    uint32_t* dst1 = reinterpret_cast<uint32_t*>(&s_tab[1]);
    for (int i = threadIdx.x; i < (int)(sizeof(GcSimplexTable) / 4); i += blockDim.x) {
      dst0[i] = src0[i];
      dst1[i] = src1[i];
    }
  }
  __syncthreads();
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t idx = gtid / WALK_ITEMS;
  const int item = (int)(gtid % WALK_ITEMS);
  const bool valid = idx < nslots;  // idle lanes still take part in the shuffles
  const uint32_t slot = valid ? slots[idx] : 0u;
  const int64_t s = valid ? (int64_t)mu.slow_simp[slot] : 0;
  const int L = lv.L;
  const int32_t bgcell1 = cutcells[s / g.nt0];
  uint32_t* val = mu.slow_val + slot;  // channel ch at val[ch * slow_n]
  // counting walk of this item
  uint32_t n_cells = 0, n_pts = 0, n_fac[LM], n_fpt[LM];
#pragma unroll
  for (int k = 0; k < LM; ++k) n_fac[k] = n_fpt[k] = 0;
  // the count launch walks and leaves the item's counters in mu.walk_items; the fill launch (same slot list, same gtid)
  // reads them back instead of walking twice
  uint32_t* wi = mu.walk_items + items_base + gtid;  // channel ch at wi[ch * walk_items_n]
  if (valid && !FILL) {
    walk_item<D, NW, false>(g, lv, s_tab[0], s_tab[1], simp_raw, simp_pos, lay, s, bgcell1, item, n_cells, n_pts, n_fac, n_fpt);
    wi[0] = n_cells;
    wi[mu.walk_items_n] = n_pts;
    for (int k = 0; k < L; ++k) {
      wi[(int64_t)(2 + k) * mu.walk_items_n] = n_fac[k];
      wi[(int64_t)(2 + L + k) * mu.walk_items_n] = n_fpt[k];
    }
  } else if (valid) {
    n_cells = wi[0];
    n_pts = wi[mu.walk_items_n];
    for (int k = 0; k < L; ++k) {
      n_fac[k] = wi[(int64_t)(2 + k) * mu.walk_items_n];
      n_fpt[k] = wi[(int64_t)(2 + L + k) * mu.walk_items_n];
    }
  }
  // inclusive scan over the WALK_ITEMS lanes of the simplex
  const int lane = threadIdx.x & 31;
  auto seg_scan = [&](uint32_t v) -> uint32_t {
#pragma unroll
    for (int o = 1; o < WALK_ITEMS; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, v, o);
      if ((lane & (WALK_ITEMS - 1)) >= o) v += up;
    }
    return v;
  };
  if (!FILL) {
    // per-slot counters + the simplex' share of its tile's totals
    const int64_t tile = s / 32;
    const bool last = valid && item == WALK_ITEMS - 1;
    const uint32_t tc = seg_scan(n_cells), tp = seg_scan(n_pts);
    if (last) {
      val[0] = tc;
      val[mu.slow_n] = tp;
      atomicAdd(mu.tile + tile, tc);
      atomicAdd(mu.tile + mu.ntiles + tile, tp);
    }
    for (int k = 0; k < L; ++k) {
      const uint32_t tf = seg_scan(n_fac[k]), tq = seg_scan(n_fpt[k]);
      if (last) {
        val[(int64_t)(2 + k) * mu.slow_n] = tf;
        val[(int64_t)(2 + L + k) * mu.slow_n] = tq;
        if (tf) atomicAdd(mu.tile + (int64_t)(2 + k) * mu.ntiles + tile, tf);
        if (tq) atomicAdd(mu.tile + (int64_t)(2 + L + k) * mu.ntiles + tile, tq);
      }
    }
  } else {
    uint32_t cell_off = seg_scan(n_cells) - n_cells, pt_off = seg_scan(n_pts) - n_pts;
    uint32_t fac_off[LM], fpt_off[LM];
#pragma unroll
    for (int k = 0; k < LM; ++k) fac_off[k] = fpt_off[k] = 0;
    for (int k = 0; k < L; ++k) {
      fac_off[k] = seg_scan(n_fac[k]) - n_fac[k];
      fpt_off[k] = seg_scan(n_fpt[k]) - n_fpt[k];
    }
    if (valid) {
      // start offsets of the simplex, left by the fill kernel of its tile (k_cutm_fill)
      cell_off += val[0];
      pt_off += val[mu.slow_n];
      for (int k = 0; k < L; ++k) {
        fac_off[k] += mu.fac_base[k] + val[(int64_t)(2 + k) * mu.slow_n];
        fpt_off[k] += mu.fpt_base[k] + val[(int64_t)(2 + L + k) * mu.slow_n];
      }
      walk_item<D, NW, true>(g, lv, s_tab[0], s_tab[1], simp_raw, simp_pos, lay, s, bgcell1, item, cell_off, pt_off, fac_off, fpt_off);
    }
  }
}

// recorded simplices -> one dense slot list per launch bucket (any order inside a bucket); one atomic per warp and bucket
__global__ void k_cutm_bucketize(const uint8_t* __restrict__ slow_bucket, int64_t n, const uint32_t* __restrict__ base,
                                 uint32_t* __restrict__ cursor, uint32_t* __restrict__ lists) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int b = slot < n ? (int)slow_bucket[slot] : -1;
  for (int bk = 0; bk < CM_NBK; ++bk) {
    const uint32_t bm = __ballot_sync(0xffffffffu, b == bk);
    if (!bm) continue;
    uint32_t first = 0u;
    if (lane == __ffs(bm) - 1) first = atomicAdd(cursor + bk, (uint32_t)__popc(bm));
    first = __shfl_sync(0xffffffffu, first, __ffs(bm) - 1);
    if (b == bk) lists[base[bk] + first + (uint32_t)__popc(bm & ((1u << lane) - 1u))] = (uint32_t)slot;
  }
}

// ================================================================================================
// L > 1, simplices cut by AT MOST ONE level set (the bulk of every multi-level-set model): no tree, no stack, no local
// memory.  Such a simplex reaches the cutting stage `a` with its vertices permuted by the skipped stages, is cut there
// exactly like in the single-level-set path, and its children (a > 0) or itself (a == 0, or nothing cuts it) are emitted
// at stage 0; the facets of `a` pass through the other stages unchanged but for the permutations.
//   k_cutm_count : lane = simplex.  Sign bits of every level set at the D+1 vertices -> the cutting stage, the case and
//                  the all-OUT mask in ONE word per simplex (sinfo: the fill kernel does not evaluate L level sets again),
//                  tile totals per channel with warp reductions; simplices with two or more mixed-sign level sets are
//                  recorded for the tree walk.
//   k_cutm_fill  : the single-level-set phases: A lane = simplex (points into a per-lane slice of shared memory, facet
//                  normals, warp scans of the counters -> offsets), F lane = subfacet element, P warp copy-out of the
//                  subcell points through a point map, C lane = subcell (connectivity row, bg cell, measure, L state
//                  bytes: neighbouring lanes write neighbouring addresses).
// ================================================================================================
constexpr int CM_T = 64;  // threads per block of the count kernel: 2 independent warp tiles
// fill kernel: 9 independent warp tiles per block, 2 blocks = 18 warps per SM (2 x (9 x 12.2 + 3) KB of shared memory).  The
// warps of a block start together and run the same phases at about the same time, so they share the instruction lines they
// fetch: the kernel is ~64 KB of code, with 2-warp blocks the 16 warps of an SM were spread all over it and instruction
// fetch was the largest stall (ncu r2_c1_final: no_instruction 2.1 per issue).  512^3 CSG: 11.9 ms with 64 threads,
// 10.3 with 256, 9.8 with 288, 11.3 with 512 (one block per SM: the SM idles while the block's last warps finish).
#ifndef CMF_T
#define CMF_T 288
#endif

template <int D>
struct CmLane {
  int ijk[3];
  uint32_t lv4;  // local vertex ids of the n-cube, one byte per simplex vertex (first two swapped when det J < 0)
  uint32_t rv4;  // reference coordinates of the vertices in the bg cell, one bit per direction
  int32_t bgcell1;
};

// Jacobian sign of every stage-0 simplex type (_ensure_positive_jacobians! in physical space, see k_cut1_fill)
template <int D>
__device__ __forceinline__ void cm_swap_flags(const GcGrid& g, const uint8_t* vtab, uint8_t* s_swap) {
  if ((int)threadIdx.x < (g.simplexified ? g.cpc : g.nt0)) {
    double x[D + 1][D];
#pragma unroll
    for (int m = 0; m <= D; ++m) {
      const int lvid = vtab[threadIdx.x * 4 + m];
#pragma unroll
      for (int d = 0; d < D; ++d) x[m][d] = gc_coord(g, d, (d == D - 1 ? g.k0 : 0) + ((lvid >> d) & 1));
    }
    double J[3][3];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) J[a][b] = __dsub_rn(x[a + 1][b], x[0][b]);
    double dV;
    if (D == 2) {
      dV = __dsub_rn(__dmul_rn(J[0][0], J[1][1]), __dmul_rn(J[0][1], J[1][0]));
    } else {
      const double p1 = __dmul_rn(__dmul_rn(J[0][0], J[1][1]), J[2][2]);
      const double p2 = __dmul_rn(__dmul_rn(J[0][1], J[1][2]), J[2][0]);
      const double p3 = __dmul_rn(__dmul_rn(J[0][2], J[1][0]), J[2][1]);
      const double q1 = __dmul_rn(__dmul_rn(J[0][0], J[1][2]), J[2][1]);
      const double q2 = __dmul_rn(__dmul_rn(J[0][1], J[1][0]), J[2][2]);
      const double q3 = __dmul_rn(__dmul_rn(J[0][2], J[1][1]), J[2][0]);
      dV = __dsub_rn(__dadd_rn(__dadd_rn(p1, p2), p3), __dadd_rn(__dadd_rn(q1, q2), q3));
    }
    s_swap[threadIdx.x] = dV < 0.0 ? 1 : 0;
  }
}

template <int D>
__device__ __forceinline__ void cm_lane_setup(const GcGrid& g, const uint8_t* vtab, const uint8_t* s_swap,
                                              const int32_t* __restrict__ cutcells, int64_t s, CmLane<D>& o) {
  const int64_t kcut = s / g.nt0;
  const int t0 = (int)(s - kcut * g.nt0);
  o.bgcell1 = cutcells[kcut];
  const uint32_t cell = (uint32_t)(o.bgcell1 - 1);  // slab-local cell ids fit in 32 bits (checked at cut time)
  const uint32_t cube = cell / (uint32_t)g.cpc;
  const int row_id = g.simplexified ? (int)(cell - cube * (uint32_t)g.cpc) : t0;
  const uint32_t row = cube / (uint32_t)g.n[0];
  o.ijk[0] = (int)(cube - row * (uint32_t)g.n[0]);
  if (D == 3) {
    const uint32_t lay3 = row / (uint32_t)g.n[1];
    o.ijk[1] = (int)(row - lay3 * (uint32_t)g.n[1]);
    o.ijk[2] = (int)lay3 + g.k0;
  } else {
    o.ijk[1] = (int)row + g.k0;
    o.ijk[2] = 0;
  }
  o.lv4 = *reinterpret_cast<const uint32_t*>(vtab + row_id * 4);
  o.rv4 = g.simplexified ? 0x04020100u : o.lv4;  // simplex cells: origin + unit axes; n-cube cells: the cube vertex
  if (s_swap[row_id]) {                          // swap the first two vertices
    o.lv4 = __byte_perm(o.lv4, 0u, 0x3201);
    o.rv4 = __byte_perm(o.rv4, 0u, 0x3201);
  }
}

// vertex order in which a simplex arrives at stage `lo` after the all-IN / all-OUT stages hi-1 ... lo+1 (each a vertex
// permutation = the single child of that table row), as 2-bit entries.  One stage = one look-up: the host tabulates, for
// the all-IN and the all-OUT row of every simplex table, packed order -> packed order after the stage (GC_PERM_LUT_BYTES
// bytes behind the tables); the kernels keep the two tables they need in shared memory.  (The first version composed byte
// arrays in local memory: 240 M sectors of local loads in k_cutm_fill on the 512^3 CSG model.)
constexpr int GC_PERM_LUT_BYTES = 3 * 512;  // [simplex dimension - 1][all-IN, all-OUT][256]

template <int D>
__device__ __forceinline__ void cm_load_perm_luts(const GcSimplexTable* __restrict__ tables, uint8_t* s_plut, uint8_t* s_flut) {
  const uint32_t* base = reinterpret_cast<const uint32_t*>(tables + 3);
  const uint32_t* pc = base + (D - 1) * 128;
  const uint32_t* pf = base + (D - 2) * 128;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) reinterpret_cast<uint32_t*>(s_plut)[i] = pc[i];
  // facets have at most 3 vertices: 64 packed orders per row
  for (int i = threadIdx.x; i < 32; i += blockDim.x)
    reinterpret_cast<uint32_t*>(s_flut)[i] = pf[(i >> 4) * 64 + (i & 15)];
}

template <int D>
__device__ __forceinline__ uint32_t cm_skip_cell_stages(const uint8_t* lut, uint32_t outm, int hi, int lo) {
  uint32_t q = (D == 3) ? 0xE4u : 0x24u;  // identity: entry i = i
  for (int k = hi - 1; k > lo; --k) q = lut[((outm >> k) & 1u) * 256u + q];
  return q;
}

template <int D>
__device__ __forceinline__ uint32_t cm_skip_facet_stages(const uint8_t* lut, uint32_t outm, int hi, int lo, int skip) {
  uint32_t q = (D == 3) ? 0x24u : 0x04u;
  for (int k = hi - 1; k > lo; --k) {
    if (k == skip) continue;
    q = lut[((outm >> k) & 1u) * 64u + q];
  }
  return q;
}

template <int D>
__global__ void __launch_bounds__(CM_T)
k_cutm_count(const GcGrid g, const GcLeaves lv, const GcSimplexTable* __restrict__ tables,
             const uint8_t* __restrict__ simp_raw, const uint8_t* __restrict__ simp_pos,
             const int32_t* __restrict__ cutcells, int64_t nsimp, GcMulti mu) {
  __shared__ GcSimplexTable s_tab;
  __shared__ uint8_t s_swap[8];
  __shared__ __align__(4) uint8_t s_plut[512], s_flut[128];
  cm_load_perm_luts<D>(tables, s_plut, s_flut);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables + (D - 1));
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
    for (int i = threadIdx.x; i < (int)(sizeof(GcSimplexTable) / 4); i += blockDim.x) dst[i] = src[i];
  }
  const uint8_t* vtab = (g.simplexified ? simp_raw : simp_pos) + (D - 2) * 24;
  cm_swap_flags<D>(g, vtab, s_swap);
  __syncthreads();
  const GcSimplexTable& TC = s_tab;
  const int L = lv.L;
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * (CM_T / 32) + (threadIdx.x >> 5);
  if (tile * 32 >= nsimp) return;  // whole warp beyond the end (no block barrier below)
  const int64_t s = tile * 32 + lane;
  uint32_t ncell = 0, npt = 0, nfac = 0;
  int a = -1;  // cutting stage of a fast simplex with facets
  int rec_bucket = -1;
  uint32_t rec_mask = 0u;
  if (s < nsimp) {
    CmLane<D> ln;
    cm_lane_setup<D>(g, vtab, s_swap, cutcells, s, ln);
    double X[D + 1][3];
    int64_t node[D + 1];
#pragma unroll
    for (int m = 0; m <= D; ++m) {
      const int lvid = (ln.lv4 >> (8 * m)) & 0xFFu;
      const int nijk[3] = {ln.ijk[0] + (lvid & 1), ln.ijk[1] + ((lvid >> 1) & 1), ln.ijk[2] + ((lvid >> 2) & 1)};
      X[m][0] = gc_coord(g, 0, nijk[0]);
      X[m][1] = gc_coord(g, 1, nijk[1]);
      X[m][2] = D == 3 ? gc_coord(g, 2, nijk[2]) : 0.0;
      node[m] = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                         : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
    }
    // sign bits of every level set at the vertices (in stage-0 order): one nibble per level set is enough for the case
    uint32_t mixed = 0u, outm = 0u, sa = 0u;
    for (int k = 0; k < L; ++k) {
      uint32_t sb = 0u;
#pragma unroll
      for (int m = 0; m <= D; ++m) {
        const double v = lv.leaf[k].kind == GECUT_LEAF_NODAL ? __ldg(lv.nodal[k] + (node[m] - lv.nodal_base))
                                                              : gc_eval_leaf(lv.leaf[k], X[m], D);
        sb |= (v > 0.0 ? 1u : 0u) << m;
      }
      const uint32_t full = (1u << (D + 1)) - 1u;
      if (sb == full) outm |= 1u << k;
      else if (sb != 0u) {
        mixed |= 1u << k;
        sa = sb;  // the highest mixed level set stays
      }
    }
    if (__popc(mixed) > 1) {
      // recorded: exactly two mixed level sets -> the stackless depth-2 kernel; more -> the generic tree walk, bucketed by
      // the number of level sets it has to track
      int bucket = 0;
      if (__popc(mixed) > 2) {
        const int nw = __popc((mixed | 3u) & ((1u << L) - 1u));
        while (nw > walk_nw(bucket)) bucket++;
        bucket++;
      }
      rec_bucket = bucket;
      rec_mask = mixed | (outm << 16);
    } else {
      const int aa = mixed ? (31 - __clz(mixed)) : 0;  // the only stage that can cut (stage 0 is visited in any case)
      const uint32_t vl = cm_skip_cell_stages<D>(s_plut, outm, L, aa);
      if (!mixed) sa = ((outm >> aa) & 1u) ? (1u << (D + 1)) - 1u : 0u;
      int c = 0;  // case in arrival order
#pragma unroll
      for (int i = 0; i <= D; ++i) c |= (int)((sa >> ((vl >> (2 * i)) & 3u)) & 1u) << i;
      ncell = TC.nsub[c];
      nfac = TC.nfac[c];
      npt = aa == 0 ? TC.npts[c] : ncell * (D + 1);
      if (nfac) a = aa;
      mu.sinfo[s] = outm | ((uint32_t)aa << 16) | ((uint32_t)c << 20) | (mixed ? 1u << 24 : 0u);
    }
  }
  // recorded simplices: one slot each; the lanes of a warp share one atomic per counter
  {
    const uint32_t recm = __ballot_sync(0xffffffffu, rec_bucket >= 0);
    if (recm) {
      uint32_t base = 0u;
      if (lane == __ffs(recm) - 1) base = atomicAdd(&mu.slow_count[CM_NBK], (uint32_t)__popc(recm));
      base = __shfl_sync(0xffffffffu, base, __ffs(recm) - 1);
      if (rec_bucket >= 0) {
        const uint32_t slot = base + (uint32_t)__popc(recm & ((1u << lane) - 1u));
        mu.slow_simp[slot] = (uint32_t)s;
        mu.slow_bucket[slot] = (uint8_t)rec_bucket;
        mu.slow_mask[slot] = rec_mask;
        mu.sinfo[s] = CM_SLOW | slot;
      }
      for (int bk = 0; bk < CM_NBK; ++bk) {
        const uint32_t bm = __ballot_sync(0xffffffffu, rec_bucket == bk);
        if (bm && lane == __ffs(bm) - 1) atomicAdd(&mu.slow_count[bk], (uint32_t)__popc(bm));
      }
    }
  }
  // tile totals of the fast simplices (the walk adds the recorded ones)
  const uint32_t tc = __reduce_add_sync(0xffffffffu, ncell), tp = __reduce_add_sync(0xffffffffu, npt);
  const uint32_t present = __reduce_or_sync(0xffffffffu, a >= 0 ? 1u << a : 0u);
  if (lane == 0) {
    mu.tile[tile] = tc;
    mu.tile[mu.ntiles + tile] = tp;
  }
  for (int k = 0; k < L; ++k) {
    uint32_t tf = 0u;
    if ((present >> k) & 1u) tf = __reduce_add_sync(0xffffffffu, a == k ? nfac : 0u);
    if (lane == 0) {
      mu.tile[(int64_t)(2 + k) * mu.ntiles + tile] = tf;
      mu.tile[(int64_t)(2 + L + k) * mu.ntiles + tile] = tf * D;  // a passed-through facet keeps its D points
    }
  }
}

// ---- n-cube cells: the count pass in two kernels.  The NT0 stage-0 simplices of a cut cell share the 2^D vertices of the
// cube, so ONE thread per cut cell evaluates the L level sets at the 2^D vertices (3x fewer leaf evaluations than D+1
// vertices per simplex: every cube vertex sits in 3 of the 6 TETs) and writes the info words / records of its NT0 simplices;
// the tile totals are a second, memory-light pass over the info words (lane = simplex).
template <int D>
__global__ void __launch_bounds__(128)
k_cutm_count_cells(const GcGrid g, const GcLeaves lv, const GcSimplexTable* __restrict__ tables,
                   const uint8_t* __restrict__ simp_pos, const int32_t* __restrict__ cutcells, int64_t ncut, GcMulti mu) {
  constexpr int NT0 = (D == 3) ? 6 : 2, NVC = 1 << D;
  __shared__ GcSimplexTable s_tab;
  __shared__ uint8_t s_swap[8];
  __shared__ uint32_t s_lv4[8];
  __shared__ __align__(4) uint8_t s_plut[512], s_flut[128];
  cm_load_perm_luts<D>(tables, s_plut, s_flut);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables + (D - 1));
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
    for (int i = threadIdx.x; i < (int)(sizeof(GcSimplexTable) / 4); i += blockDim.x) dst[i] = src[i];
  }
  const uint8_t* vtab = simp_pos + (D - 2) * 24;
  cm_swap_flags<D>(g, vtab, s_swap);
  __syncthreads();
  if (threadIdx.x < NT0) {  // vertex ids of every stage-0 simplex type, first two swapped when its Jacobian is negative
    uint32_t l4 = *reinterpret_cast<const uint32_t*>(vtab + threadIdx.x * 4);
    if (s_swap[threadIdx.x]) l4 = __byte_perm(l4, 0u, 0x3201);
    s_lv4[threadIdx.x] = l4;
  }
  __syncthreads();
  const GcSimplexTable& TC = s_tab;
  const int L = lv.L;
  const int64_t kcut = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (kcut >= ncut) return;
  const uint32_t cell = (uint32_t)(cutcells[kcut] - 1);
  int ijk[3];
  const uint32_t row = cell / (uint32_t)g.n[0];
  ijk[0] = (int)(cell - row * (uint32_t)g.n[0]);
  if (D == 3) {
    const uint32_t lay3 = row / (uint32_t)g.n[1];
    ijk[1] = (int)(row - lay3 * (uint32_t)g.n[1]);
    ijk[2] = (int)lay3 + g.k0;
  } else {
    ijk[1] = (int)row + g.k0;
    ijk[2] = 0;
  }
  double xc[3][2];  // the two node coordinates of the cube per direction
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    xc[d][0] = d < D ? gc_coord(g, d, ijk[d]) : 0.0;
    xc[d][1] = d < D ? gc_coord(g, d, ijk[d] + 1) : 0.0;
  }
  const int64_t node0 = (D == 3) ? ((int64_t)ijk[0] + (int64_t)g.nn[0] * ((int64_t)ijk[1] + (int64_t)g.nn[1] * ijk[2]))
                                 : ((int64_t)ijk[0] + (int64_t)g.nn[0] * ijk[1]);
  uint32_t mixed[NT0], outm[NT0], sa[NT0];
#pragma unroll
  for (int t = 0; t < NT0; ++t) mixed[t] = outm[t] = sa[t] = 0u;
  for (int k = 0; k < L; ++k) {
    const bool nodal = lv.leaf[k].kind == GECUT_LEAF_NODAL;
    uint32_t cube = 0u;
#pragma unroll
    for (int v = 0; v < NVC; ++v) {
      double val;
      if (nodal) {
        const int64_t node = node0 + (v & 1) + (int64_t)g.nn[0] * (((v >> 1) & 1) + (D == 3 ? (int64_t)g.nn[1] * ((v >> 2) & 1) : 0));
        val = __ldg(lv.nodal[k] + (node - lv.nodal_base));
      } else {
        const double x[3] = {xc[0][v & 1], xc[1][(v >> 1) & 1], D == 3 ? xc[2][(v >> 2) & 1] : 0.0};
        val = gc_eval_leaf(lv.leaf[k], x, D);
      }
      cube |= (val > 0.0 ? 1u : 0u) << v;
    }
    const uint32_t full = (1u << (D + 1)) - 1u;
#pragma unroll
    for (int t = 0; t < NT0; ++t) {
      const uint32_t l4 = s_lv4[t];
      uint32_t sb = 0u;
#pragma unroll
      for (int m = 0; m <= D; ++m) sb |= ((cube >> ((l4 >> (8 * m)) & 0xFFu)) & 1u) << m;
      if (sb == full) outm[t] |= 1u << k;
      else if (sb != 0u) {
        mixed[t] |= 1u << k;
        sa[t] = sb;  // the highest mixed level set stays
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NT0; ++t) {
    const int64_t s = kcut * NT0 + t;
    if (__popc(mixed[t]) > 1) {
      // recorded: exactly two mixed level sets -> the stackless depth-2 kernel; more -> the generic tree walk, bucketed by
      // the number of level sets it has to track
      int bucket = 0;
      if (__popc(mixed[t]) > 2) {
        const int nw = __popc((mixed[t] | 3u) & ((1u << L) - 1u));
        while (nw > walk_nw(bucket)) bucket++;
        bucket++;
      }
      const uint32_t slot = atomicAdd(&mu.slow_count[CM_NBK], 1u);
      atomicAdd(&mu.slow_count[bucket], 1u);
      mu.slow_simp[slot] = (uint32_t)s;
      mu.slow_bucket[slot] = (uint8_t)bucket;
      mu.slow_mask[slot] = mixed[t] | (outm[t] << 16);
      mu.sinfo[s] = CM_SLOW | slot;
    } else {
      const int aa = mixed[t] ? (31 - __clz(mixed[t])) : 0;  // the only stage that can cut (stage 0 is visited in any case)
      const uint32_t vl = cm_skip_cell_stages<D>(s_plut, outm[t], L, aa);
      uint32_t sat = sa[t];
      if (!mixed[t]) sat = ((outm[t] >> aa) & 1u) ? (1u << (D + 1)) - 1u : 0u;
      int c = 0;  // case in arrival order
#pragma unroll
      for (int i = 0; i <= D; ++i) c |= (int)((sat >> ((vl >> (2 * i)) & 3u)) & 1u) << i;
      mu.sinfo[s] = outm[t] | ((uint32_t)aa << 16) | ((uint32_t)c << 20) | (mixed[t] ? 1u << 24 : 0u);
    }
  }
  (void)TC;
}

// tile totals of the fast simplices from their info words (the walk adds the recorded ones)
template <int D>
__global__ void __launch_bounds__(256)
k_cutm_tiles(const GcSimplexTable* __restrict__ tables, int L, int64_t nsimp, GcMulti mu) {
  __shared__ GcSimplexTable s_tab;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables + (D - 1));
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
    for (int i = threadIdx.x; i < (int)(sizeof(GcSimplexTable) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const GcSimplexTable& TC = s_tab;
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (tile * 32 >= nsimp) return;
  const int64_t s = tile * 32 + lane;
  uint32_t ncell = 0, npt = 0, nfac = 0;
  int a = -1;
  if (s < nsimp) {
    const uint32_t info = mu.sinfo[s];
    if (!(info & CM_SLOW)) {
      const int aa = (int)((info >> 16) & 15u), c = (int)((info >> 20) & 15u);
      ncell = TC.nsub[c];
      nfac = TC.nfac[c];
      npt = aa == 0 ? TC.npts[c] : ncell * (D + 1);
      if (nfac) a = aa;
    }
  }
  const uint32_t tc = __reduce_add_sync(0xffffffffu, ncell), tp = __reduce_add_sync(0xffffffffu, npt);
  const uint32_t present = __reduce_or_sync(0xffffffffu, a >= 0 ? 1u << a : 0u);
  // channel ch of the tile goes out through lane ch (ch = 0 subcells, 1 points, 2 + k facets of level set k, 2 + L + k their
  // points): one store instruction per tile instead of 2 + 2 L by lane 0
  uint32_t v0 = lane == 0 ? tc : (lane == 1 ? tp : 0u), v1 = 0u;
  for (int k = 0; k < L; ++k) {
    uint32_t tf = 0u;
    if ((present >> k) & 1u) tf = __reduce_add_sync(0xffffffffu, a == k ? nfac : 0u);
    if (lane == 2 + k) v0 = tf;
    if (lane == 2 + L + k) v0 = tf * D;  // a passed-through facet keeps its D points
    if (lane + 32 == 2 + L + k) v1 = tf * D;
  }
  if (lane < 2 + 2 * L) mu.tile[(int64_t)lane * mu.ntiles + tile] = v0;
  if (lane + 32 < 2 + 2 * L) mu.tile[(int64_t)(lane + 32) * mu.ntiles + tile] = v1;
}

// shared memory of one warp of k_cutm_fill
template <int D>
struct CmSmem {
  static constexpr int NPC = WalkCfg<D>::NPC;
  // per lane: x of <= NPC points, the coefficients of the <= NPC-D-1 cut edges, one word of reference-coordinate CODES
  // (2 bits per point and direction: 0, 1, c, 1 - c with c the point's coefficient -- a vertex of a stage-0 simplex has
  // reference coordinates in {0, 1}, a cut point is gc_lerp of two of them); the reference coordinates themselves are
  // rebuilt where they are written (bit for bit: gc_lerp(0,1,c) = c, gc_lerp(1,0,c) = 1 - c).  29 instead of 49 doubles per
  // lane: 12.2 instead of 17.3 KB per warp = 16 instead of 12 warps per SM
  static constexpr int NCE = NPC - (D + 1);
  static constexpr int COFF = NPC * D;          // first coefficient
  static constexpr int RCOFF = COFF + NCE;      // the code word
  static constexpr int SSTR = ((RCOFF + 1) | 1);  // odd stride
  static constexpr int PCAP = 768;              // points of a tile that go through the staged copy-out (32 simplices x 6 children x 4;
                                                // more -- tiles with recorded simplices -- fall back to direct stores)
  static constexpr int NSUBMAX = (D == 3) ? 6 : 3;
  double sx[32 * SSTR];
  double fstash[32][8];   // <= 2 facets per lane: unit normal (D) + measure
  uint4 rec[32];          // per lane: case | a << 4 | qc << 8 | first fast subcell << 16, outm, first subcell, first point
  int32_t bg[32];
  uint16_t pmap[PCAP];    // tile point -> (lane << 3 | slot) of the staged copy
  uint8_t fslots[32][8];  // point slots of the <= 2 facets of a lane
  uint8_t own_c[32 * NSUBMAX];  // fast subcell of the tile -> lane
  uint8_t fitem[64];
};

// reference coordinate `d` of point slot `p` of a lane slice (see CmSmem)
template <int D>
__device__ __forceinline__ double cm_rcoord(const double* SXo, int p, int d) {
  using SM = CmSmem<D>;
  const unsigned long long rc = *reinterpret_cast<const unsigned long long*>(SXo + SM::RCOFF);
  const unsigned code = (unsigned)(rc >> (8 * p + 2 * d)) & 3u;
  if (code < 2u) return (double)code;
  const double c = SXo[SM::COFF + (p - (D + 1))];
  return code == 2u ? c : __dsub_rn(1.0, c);
}

template <int D>
__global__ void __launch_bounds__(CMF_T)
k_cutm_fill(const GcGrid g, const GcLeaves lv, const GcSimplexTable* __restrict__ tables,
            const uint8_t* __restrict__ simp_raw, const uint8_t* __restrict__ simp_pos,
            const int32_t* __restrict__ cutcells, int64_t nsimp, GcMulti mu, GcLayout lay,
            uint32_t* __restrict__ sc_ptr) {
  using SM = CmSmem<D>;
  constexpr int SSTR = SM::SSTR, PCAP = SM::PCAP;
  constexpr uint32_t FULLM = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char cm_smem[];
  __shared__ GcSimplexTable s_tab[2];  // [0]: D-simplex (cells), [1]: (D-1)-simplex (facet pass-through rows)
  __shared__ uint8_t s_swap[8];
  __shared__ __align__(4) uint8_t s_plut[512], s_flut[128];
  cm_load_perm_luts<D>(tables, s_plut, s_flut);
  {
    const uint32_t* src0 = reinterpret_cast<const uint32_t*>(tables + (D - 1));
    const uint32_t* src1 = reinterpret_cast<const uint32_t*>(tables + (D - 2));
    uint32_t* dst0 = reinterpret_cast<uint32_t*>(&s_tab[0]);
    uint32_t* dst1 = reinterpret_cast<uint32_t*>(&s_tab[1]);
    for (int i = threadIdx.x; i < (int)(sizeof(GcSimplexTable) / 4); i += blockDim.x) {
      dst0[i] = src0[i];
      dst1[i] = src1[i];
    }
  }
  const uint8_t* vtab = (g.simplexified ? simp_raw : simp_pos) + (D - 2) * 24;
  cm_swap_flags<D>(g, vtab, s_swap);
  __syncthreads();
  const GcSimplexTable& TC = s_tab[0];
  const GcSimplexTable& TF = s_tab[1];
  const int L = lv.L;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  SM& sm = *reinterpret_cast<SM*>(cm_smem + (size_t)wid * sizeof(SM));
  const int64_t tile = (int64_t)blockIdx.x * (CMF_T / 32) + wid;
  if (tile * 32 >= nsimp) return;  // whole warp beyond the end (no block barrier below)
  const int64_t s = tile * 32 + lane;
  const bool valid = s < nsimp;
  const uint32_t lt = (1u << lane) - 1u;
  const int c_allout = (1 << (D + 1)) - 1, fc_allout = (1 << D) - 1;

  // ======================================== phase A: lane = stage-0 simplex
  const uint32_t info = valid ? mu.sinfo[s] : 0u;
  const bool slow = valid && (info & CM_SLOW);
  const bool fast = valid && !slow;
  const uint32_t slot = info & ~CM_SLOW;
  const uint32_t outm = info & 0xFFFFu;
  const int a = fast ? (int)((info >> 16) & 15u) : 0;
  const int c = fast ? (int)((info >> 20) & 15u) : 0;
  uint32_t ncell = 0, npt = 0, nfac = 0;
  if (fast) {
    ncell = TC.nsub[c];
    nfac = TC.nfac[c];
    npt = a == 0 ? TC.npts[c] : ncell * (D + 1);
  } else if (slow) {
    ncell = mu.slow_val[slot];
    npt = mu.slow_val[mu.slow_n + slot];
  }
  // warp scans: all subcells / all points (recorded simplices included: they own ranges of the same arrays), and the
  // fast-only subcell index of the lane-per-subcell phase
  auto excl_scan = [&](uint32_t v, uint32_t* total) -> uint32_t {
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(FULLM, inc, o);
      if (lane >= o) inc += t;
    }
    *total = __shfl_sync(FULLM, inc, 31);
    return inc - v;
  };
  uint32_t tcells, tpts, tcf;
  const uint32_t lc = excl_scan(ncell, &tcells), lp = excl_scan(npt, &tpts);
  const uint32_t lcf = excl_scan(fast ? ncell : 0u, &tcf);
  const uint32_t cell_off = mu.tile[tile] + lc, pt_off = mu.tile[mu.ntiles + tile] + lp;
  const uint32_t wbase = pt_off - lp, wcount = tpts;  // point range of the tile
  const bool staged = wcount <= (uint32_t)PCAP;
  // facets: offsets inside the arrays of level set a
  uint32_t fac_off = 0, fpt_off = 0;
  const uint32_t any_slow = __ballot_sync(FULLM, slow);
  if (!any_slow) {
    const uint32_t peers = __match_any_sync(FULLM, nfac ? a : 31);
    const uint32_t b1 = __ballot_sync(FULLM, nfac >= 1), b2 = __ballot_sync(FULLM, nfac >= 2);
    const uint32_t before = (uint32_t)(__popc(peers & b1 & lt) + __popc(peers & b2 & lt));
    if (nfac) {
      fac_off = mu.fac_base[a] + mu.tile[(int64_t)(2 + a) * mu.ntiles + tile] + before;
      fpt_off = mu.fpt_base[a] + mu.tile[(int64_t)(2 + L + a) * mu.ntiles + tile] + before * D;
    }
  } else {
    // a tile with recorded simplices: they own facet ranges of several level sets -> one scan per level set and array
    for (int k = 0; k < L; ++k) {
      uint32_t tot;
      const uint32_t vf = fast ? (a == k ? nfac : 0u) : (slow ? mu.slow_val[(int64_t)(2 + k) * mu.slow_n + slot] : 0u);
      const uint32_t ef = excl_scan(vf, &tot);
      const uint32_t vq = fast ? (a == k ? nfac * D : 0u) : (slow ? mu.slow_val[(int64_t)(2 + L + k) * mu.slow_n + slot] : 0u);
      const uint32_t eq = excl_scan(vq, &tot);
      const uint32_t bf = mu.tile[(int64_t)(2 + k) * mu.ntiles + tile] + ef;
      const uint32_t bq = mu.tile[(int64_t)(2 + L + k) * mu.ntiles + tile] + eq;
      if (fast && a == k && nfac) {
        fac_off = mu.fac_base[k] + bf;
        fpt_off = mu.fpt_base[k] + bq;
      }
      if (slow) {  // counters -> start offsets (without the level-set base: the walk adds it)
        mu.slow_val[(int64_t)(2 + k) * mu.slow_n + slot] = bf;
        mu.slow_val[(int64_t)(2 + L + k) * mu.slow_n + slot] = bq;
      }
    }
  }
  if (slow) {
    mu.slow_val[slot] = cell_off;
    mu.slow_val[mu.slow_n + slot] = pt_off;
  }
  if (valid) {  // first subcell / first point of every cut cell (consumers: Laplacian, cell measures)
    const int64_t kcut = s / g.nt0;
    if (s - kcut * g.nt0 == 0) {
      sc_ptr[2 * kcut] = cell_off;
      sc_ptr[2 * kcut + 1] = pt_off;
    }
    if (s == nsimp - 1) {
      sc_ptr[2 * kcut + 2] = cell_off + ncell;
      sc_ptr[2 * kcut + 3] = pt_off + npt;
    }
  }
  if (staged) {
    uint4* pm4 = reinterpret_cast<uint4*>(sm.pmap) + lane * (PCAP / 256);
#pragma unroll
    for (int i = 0; i < PCAP / 256; ++i) pm4[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  }
  __syncwarp();
  double* SX = sm.sx + lane * SSTR;  // x of slot p at SX[p * D + d]; coefficients and codes behind (CmSmem)
  unsigned long long rcodes = 0ull;
  int w_fc = 0;
  if (fast) {
    CmLane<D> ln;
    cm_lane_setup<D>(g, vtab, s_swap, cutcells, s, ln);
    const bool mixed = (info >> 24) & 1u;
    const uint32_t vl = cm_skip_cell_stages<D>(s_plut, outm, L, a);
    // vertices in arrival order into slots 0..D; values of level set a (the only one whose values matter) alongside
    double W[D + 1];
#pragma unroll
    for (int i = 0; i <= D; ++i) {
      const int m = (vl >> (2 * i)) & 3u;
      const int lvid = (ln.lv4 >> (8 * m)) & 0xFFu;
      const int nijk[3] = {ln.ijk[0] + (lvid & 1), ln.ijk[1] + ((lvid >> 1) & 1), ln.ijk[2] + ((lvid >> 2) & 1)};
      const double x[3] = {gc_coord(g, 0, nijk[0]), gc_coord(g, 1, nijk[1]), D == 3 ? gc_coord(g, 2, nijk[2]) : 0.0};
#pragma unroll
      for (int d = 0; d < D; ++d) {
        SX[i * D + d] = x[d];
        rcodes |= (unsigned long long)((ln.rv4 >> (8 * m + d)) & 1u) << (8 * i + 2 * d);
      }
      W[i] = 0.0;
      if (mixed) {
        if (lv.leaf[a].kind == GECUT_LEAF_NODAL) {
          const int64_t node = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                                        : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
          W[i] = __ldg(lv.nodal[a] + (node - lv.nodal_base));
        } else {
          W[i] = gc_eval_leaf(lv.leaf[a], x, D);
        }
      }
    }
    int np = D + 1;
    if (TC.inoutcut[c] == GC_CUT) {
      for (int ed = 0; ed < TC.nedges; ++ed) {
        const int ia = TC.edges[ed][0], ib = TC.edges[ed][1];
        if (((c >> ia) ^ (c >> ib)) & 1) {
          double wa = W[0], wb = W[0];
#pragma unroll
          for (int m = 1; m <= D; ++m) {
            if (ia == m) wa = W[m];
            if (ib == m) wb = W[m];
          }
          const double w1 = fabs(wa), w2 = fabs(wb);
          const double coeff = __ddiv_rn(w1, __dadd_rn(w1, w2));
#pragma unroll
          for (int d = 0; d < D; ++d) {
            SX[np * D + d] = gc_lerp(SX[ia * D + d], SX[ib * D + d], coeff);
            const unsigned ra = (unsigned)(rcodes >> (8 * ia + 2 * d)) & 3u, rb = (unsigned)(rcodes >> (8 * ib + 2 * d)) & 3u;
            // the ends are vertices (codes 0 / 1): equal -> that value, 0 -> 1: c, 1 -> 0: 1 - c
            rcodes |= (unsigned long long)(ra == rb ? ra : (ra ? 3u : 2u)) << (8 * np + 2 * d);
          }
          SX[SM::COFF + (np - (D + 1))] = coeff;
          np++;
        }
      }
    }
    *reinterpret_cast<unsigned long long*>(SX + SM::RCOFF) = rcodes;
    // ---- facets of level set a: one final facet each, emitted at the last other stage
    if (nfac) {
      const int klast = a == 0 ? 1 : 0;
      const int fc = ((outm >> klast) & 1u) ? fc_allout : 0;
      w_fc = fc;
      const uint32_t qf = cm_skip_facet_stages<D>(s_flut, outm, L, klast, a);  // the same for every facet of the simplex
      for (int f = 0; f < (int)nfac; ++f) {
        const uint8_t* fp = TC.fac_pts[c][f];
        double q1[3], q2[3], q3[3], nrm[3];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          q1[d] = SX[fp[0] * D + d];
          q2[d] = SX[fp[1] * D + d];
          q3[d] = SX[fp[D - 1] * D + d];
        }
        unit_normal<D>(q1, q2, q3, (double)TC.fac_orient[c][f], nrm);
        uint8_t fvl[D];
#pragma unroll
        for (int i = 0; i < D; ++i) fvl[i] = fp[(qf >> (2 * i)) & 3u];
        double pf[D][3];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const int li = TF.sub_pts[fc][0][i];
#pragma unroll
          for (int d = 0; d < D; ++d) pf[i][d] = SX[fvl[li] * D + d];
        }
#pragma unroll
        for (int d = 0; d < D; ++d) sm.fstash[lane][f * 4 + d] = nrm[d];
        sm.fstash[lane][f * 4 + 3] = integrate_one<D - 1>(simplex_meas<D, D - 1>(pf));
#pragma unroll
        for (int i = 0; i < D; ++i) sm.fslots[lane][f * 4 + i] = fvl[i];
      }
    }
    // ---- subcell points: where every point of the final layout comes from (slot of this lane's slice)
    uint32_t qc = 0;
    if (a == 0) {
      for (int pnt = 0; pnt < np; ++pnt) {
        if (staged) {
          sm.pmap[pt_off + pnt - wbase] = (uint16_t)((lane << 3) | pnt);
        } else {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            lay.sc_X[(int64_t)(pt_off + pnt) * D + d] = SX[pnt * D + d];
            lay.sc_R[(int64_t)(pt_off + pnt) * D + d] = cm_rcoord<D>(SX, pnt, d);
          }
        }
      }
    } else {
      // cut at stage a > 0: every child goes down to stage 0 alone (composed permutations of the stages a-1 ... 1, the
      // same for every child) and is emitted there as an uncut simplex with its own copy of its D+1 points
      qc = cm_skip_cell_stages<D>(s_plut, outm, a, 0);
      for (int j = 0; j < (int)ncell; ++j) {
        const uint32_t pb = pt_off + (uint32_t)j * (D + 1);
#pragma unroll
        for (int pnt = 0; pnt <= D; ++pnt) {
          const int sl = TC.sub_pts[c][j][(qc >> (2 * pnt)) & 3u];
          if (staged) {
            sm.pmap[pb + pnt - wbase] = (uint16_t)((lane << 3) | sl);
          } else {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              lay.sc_X[(int64_t)(pb + pnt) * D + d] = SX[sl * D + d];
              lay.sc_R[(int64_t)(pb + pnt) * D + d] = cm_rcoord<D>(SX, sl, d);
            }
          }
        }
      }
    }
    sm.rec[lane] = make_uint4((uint32_t)c | ((uint32_t)a << 4) | (qc << 8) | (lcf << 16), outm, cell_off, pt_off);
    sm.bg[lane] = ln.bgcell1;
    for (int j = 0; j < (int)ncell; ++j) sm.own_c[lcf + j] = (uint8_t)lane;
  }
  __syncwarp();
  const double* SW = sm.sx;
  // ======================================== phase F: lane = element of the tile's subfacets
  // the facets of the warp's simplices form (per level set) runs of consecutive ids, so one facet entity per lane writes
  // point blocks, normals, connectivity, bg cells, measures and state rows with neighbouring lanes on neighbouring addresses
  {
    const uint32_t b1 = __ballot_sync(FULLM, fast && nfac >= 1), b2 = __ballot_sync(FULLM, fast && nfac >= 2);
    const int T = __popc(b1) + __popc(b2);
    if (T) {
      const int first = __popc(b1 & lt) + __popc(b2 & lt);
      if (fast)
        for (int f = 0; f < (int)nfac; ++f) sm.fitem[first + f] = (uint8_t)(lane * 2 + f);
      __syncwarp();
      // point blocks: D points x D coordinates per facet
      for (int q0 = 0; q0 < T * D * D; q0 += 32) {
        const int q = q0 + lane;
        const bool on = q < T * D * D;
        const int item = on ? q / (D * D) : 0;
        const int r = q - item * (D * D), pnt = r / D, d = r - pnt * D;
        const int it = sm.fitem[item], o = it >> 1, f = it & 1;
        const uint32_t fpo = __shfl_sync(FULLM, fpt_off, o);
        if (on) {
          const double* SXo = SW + o * SSTR;
          const int slot = sm.fslots[o][f * 4 + pnt];
          const int64_t dst = (int64_t)(fpo + (uint32_t)(f * D + pnt)) * D + d;
          lay.sf_X[dst] = SXo[slot * D + d];
          lay.sf_R[dst] = cm_rcoord<D>(SXo, slot, d);
        }
      }
      // normals and connectivity: D entries per facet
      for (int q0 = 0; q0 < T * D; q0 += 32) {
        const int q = q0 + lane;
        const bool on = q < T * D;
        const int item = on ? q / D : 0;
        const int d = q - item * D;
        const int it = sm.fitem[item], o = it >> 1, f = it & 1;
        const uint32_t fpo = __shfl_sync(FULLM, fpt_off, o), fao = __shfl_sync(FULLM, fac_off, o);
        const int fco = __shfl_sync(FULLM, w_fc, o);
        if (on) {
          const int64_t dst = (int64_t)(fao + (uint32_t)f) * D + d;
          lay.sf_N[dst] = sm.fstash[o][f * 4 + d];
          lay.sf_c2p[dst] = (int32_t)(fpo + (uint32_t)(f * D) + TF.sub_pts[fco][0][d] + 1);
        }
      }
      // per facet: bg cell, measure, state rows
      for (int q0 = 0; q0 < T; q0 += 32) {
        const int q = q0 + lane;
        const bool on = q < T;
        const int it = sm.fitem[on ? q : 0], o = it >> 1, f = it & 1;
        const uint32_t fao = __shfl_sync(FULLM, fac_off, o), om = __shfl_sync(FULLM, outm, o);
        const int fco = __shfl_sync(FULLM, w_fc, o), ao = __shfl_sync(FULLM, a, o);
        if (on) {
          const int64_t fid = (int64_t)fao + f;
          lay.sf_c2bg[fid] = sm.bg[o];
          lay.sf_meas[fid] = sm.fstash[o][f * 4 + 3];
          const int klast = ao == 0 ? 1 : 0;
          for (int kk = 0; kk < L; ++kk)
            lay.sf_state[(int64_t)kk * lay.sf_pitch + fid] =
                (kk == ao) ? (int8_t)GECUT_INTERFACE
                           : ((kk == klast) ? TF.sub_inout[fco][0] : (((om >> kk) & 1u) ? (int8_t)GC_OUT : (int8_t)GC_IN));
        }
      }
    }
  }
  // ======================================== phase P: subcell points of the tile leave through the point map
  if (staged) {
    double* gX = lay.sc_X + (int64_t)wbase * D;
    double* gR = lay.sc_R + (int64_t)wbase * D;
    for (uint32_t q = lane; q < wcount * D; q += 32) {
      const uint32_t pnt = q / D, d = q - pnt * D;
      const uint32_t m = sm.pmap[pnt];
      if (m != 0xFFFFu) {  // points of recorded simplices are written by the walk
        const double* SXo = SW + (m >> 3) * SSTR;
        gX[q] = SXo[(m & 7u) * D + d];
        gR[q] = cm_rcoord<D>(SXo, (int)(m & 7u), (int)d);
      }
    }
  }
  // ======================================== phase C: lane = subcell of a fast simplex
  for (uint32_t idx = lane; idx < tcf; idx += 32) {
    const int o = sm.own_c[idx];
    const uint4 r = sm.rec[o];
    const int cc = r.x & 15u, aa = (r.x >> 4) & 15u;
    const uint32_t qc = (r.x >> 8) & 0xFFu;
    const int j = (int)(idx - ((r.x >> 16) & 0xFFu));
    const uint32_t om = r.y;
    const int64_t cid = (int64_t)r.z + j;
    const double* SXo = SW + o * SSTR;
    const uint32_t ptsj = *reinterpret_cast<const uint32_t*>(TC.sub_pts[cc][j]);
    int32_t row[D + 1];
    double pc[D + 1][3];
    int8_t s0;
    if (aa == 0) {
#pragma unroll
      for (int i = 0; i <= D; ++i) {
        const uint32_t li = (ptsj >> (8 * i)) & 0xFFu;
        row[i] = (int32_t)(r.w + li + 1u);
#pragma unroll
        for (int d = 0; d < D; ++d) pc[i][d] = SXo[li * D + d];
      }
      s0 = TC.sub_inout[cc][j];
    } else {
      const int c0 = (om & 1u) ? c_allout : 0;
      const uint32_t pts0 = *reinterpret_cast<const uint32_t*>(TC.sub_pts[c0][0]);
      const uint32_t pb = r.w + (uint32_t)j * (D + 1);
#pragma unroll
      for (int i = 0; i <= D; ++i) {
        const uint32_t li = (pts0 >> (8 * i)) & 0xFFu;
        row[i] = (int32_t)(pb + li + 1u);
        const uint32_t sl = (ptsj >> (8 * ((qc >> (2 * li)) & 3u))) & 0xFFu;
#pragma unroll
        for (int d = 0; d < D; ++d) pc[i][d] = SXo[sl * D + d];
      }
      s0 = TC.sub_inout[c0][0];
    }
    store_c2p_row<D>(lay.sc_c2p, (uint32_t)cid, row);
    lay.sc_c2bg[cid] = sm.bg[o];
    lay.sc_meas[cid] = integrate_one<D>(simplex_meas<D, D>(pc));
    lay.sc_state[cid] = s0;
    for (int kk = 1; kk < L; ++kk)
      lay.sc_state[(int64_t)kk * lay.sc_pitch + cid] =
          (kk == aa) ? TC.sub_inout[cc][j] : (((om >> kk) & 1u) ? (int8_t)GC_OUT : (int8_t)GC_IN);
  }
}

// ================================================================================================
// Recorded simplices cut by EXACTLY TWO level sets a > b (all but a handful of the recorded ones: two surfaces meet
// along curves, three only at points).  The refinement tree has a fixed shape, so it is walked without a stack, without
// a point pool in local memory and with a fraction of the generic walk's code:
//   stage a cuts the simplex (<= NPC points PA, shared by the 8 lanes of the simplex), its <= 2 facets are re-cut by b;
//   every child of stage a (one lane each) is cut by b (points PB of the lane), the <= 2 facets of that cut are re-cut
//   by a -- their vertices carry the ROUNDED values of level set a (zero up to round-off on the cut points of stage a),
//   exactly like the reference's set_point_data! interpolates every pending level set (CutTriangulations.jl:123-135);
//   everything below passes through the remaining all-IN / all-OUT stages as vertex permutations.
// Same work split, counters and offsets as k_cut_walk (lane j < 6: child j; lanes 6, 7: the facets of stage a).
// ================================================================================================
constexpr int TWO_T = 192;  // 24 simplices per block (6 warps that start together share their instruction fetches; 2 blocks per SM)

template <int D>
struct TwoSmem {
  static constexpr int NPC = WalkCfg<D>::NPC, NPF = WalkCfg<D>::NPF;
  static constexpr int PAS = 2 * D + 2;  // x, r, wa, wb
  static constexpr int PBS = 2 * D + 1;  // x, r, wa
  double pa[TWO_T / WALK_ITEMS][NPC * PAS];
  // cut points of stage b of the child items (6 of the 8 lanes of a simplex); a child's vertices stay in PA
  static constexpr int NEWB = NPC - (D + 1);
  double pb[TWO_T / WALK_ITEMS * 6][NEWB * PBS + 1];
  double fe[TWO_T][NPF * 2 * D + 1];
};

template <int D, bool FILL>
__global__ void __launch_bounds__(TWO_T)
k_cut_two(const GcGrid g, const GcLeaves lv, const GcSimplexTable* __restrict__ tables,
          const uint8_t* __restrict__ simp_raw, const uint8_t* __restrict__ simp_pos,
          const int32_t* __restrict__ cutcells, GcMulti mu, GcLayout lay,
          const uint32_t* __restrict__ slots, int64_t nslots) {
  using SM = TwoSmem<D>;
  constexpr int PAS = SM::PAS, PBS = SM::PBS;
  constexpr uint32_t FULLM = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char two_smem[];
  SM& sm = *reinterpret_cast<SM*>(two_smem);
  __shared__ GcSimplexTable s_tab[2];
  __shared__ uint8_t s_swap[8];
  __shared__ __align__(4) uint8_t s_plut[512], s_flut[128];
  cm_load_perm_luts<D>(tables, s_plut, s_flut);
  {
    const uint32_t* src0 = reinterpret_cast<const uint32_t*>(tables + (D - 1));
    const uint32_t* src1 = reinterpret_cast<const uint32_t*>(tables + (D - 2));
    uint32_t* dst0 = reinterpret_cast<uint32_t*>(&s_tab[0]);
    uint32_t* dst1 = reinterpret_cast<uint32_t*>(&s_tab[1]);
    for (int i = threadIdx.x; i < (int)(sizeof(GcSimplexTable) / 4); i += blockDim.x) {
      dst0[i] = src0[i];
      dst1[i] = src1[i];
    }
  }
  const uint8_t* vtab = (g.simplexified ? simp_raw : simp_pos) + (D - 2) * 24;
  cm_swap_flags<D>(g, vtab, s_swap);
  __syncthreads();
  const GcSimplexTable& TC = s_tab[0];
  const GcSimplexTable& TF = s_tab[1];
  const int L = lv.L;
  const int lane = threadIdx.x & 31;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t idx = gtid / WALK_ITEMS;
  const int item = (int)(gtid % WALK_ITEMS);
  const bool valid = idx < nslots;  // idle lanes still take part in the shuffles / warp barriers
  const uint32_t slot = valid ? slots[idx] : 0u;
  const int64_t s = valid ? (int64_t)mu.slow_simp[slot] : 0;
  const uint32_t mask = valid ? mu.slow_mask[slot] : 3u;
  const uint32_t mixed = mask & 0xFFFFu, outm = mask >> 16;
  const int a = 31 - __clz(mixed), b = __ffs(mixed) - 1;
  const int c_allout = (1 << (D + 1)) - 1, fc_allout = (1 << D) - 1;
  double* PA = sm.pa[threadIdx.x / WALK_ITEMS];
  double* PBN = sm.pb[(threadIdx.x / WALK_ITEMS) * 6 + (threadIdx.x % WALK_ITEMS < 6 ? threadIdx.x % WALK_ITEMS : 0)];
  double* FE = sm.fe[threadIdx.x];
  auto cst = [&](int k) -> int8_t { return ((outm >> k) & 1u) ? (int8_t)GC_OUT : (int8_t)GC_IN; };

  // ---- stage-0 simplex in arrival order at stage a: lane i <= D of the group sets up vertex i
  CmLane<D> ln;
  cm_lane_setup<D>(g, vtab, s_swap, cutcells, s, ln);
  const int32_t bgcell1 = ln.bgcell1;
  const uint32_t vl = cm_skip_cell_stages<D>(s_plut, outm, L, a);
  if (item <= D) {
    const int m = (vl >> (2 * item)) & 3u;
    const int lvid = (ln.lv4 >> (8 * m)) & 0xFFu;
    const int nijk[3] = {ln.ijk[0] + (lvid & 1), ln.ijk[1] + ((lvid >> 1) & 1), ln.ijk[2] + ((lvid >> 2) & 1)};
    const double x[3] = {gc_coord(g, 0, nijk[0]), gc_coord(g, 1, nijk[1]), D == 3 ? gc_coord(g, 2, nijk[2]) : 0.0};
    const int64_t node = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                                  : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
    double* p = PA + item * PAS;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      p[d] = x[d];
      p[D + d] = (double)((ln.rv4 >> (8 * m + d)) & 1u);
    }
    p[2 * D] = lv.leaf[a].kind == GECUT_LEAF_NODAL ? __ldg(lv.nodal[a] + (node - lv.nodal_base)) : gc_eval_leaf(lv.leaf[a], x, D);
    p[2 * D + 1] = lv.leaf[b].kind == GECUT_LEAF_NODAL ? __ldg(lv.nodal[b] + (node - lv.nodal_base)) : gc_eval_leaf(lv.leaf[b], x, D);
  }
  __syncwarp();
  // ---- stage a: case, cut points (lane D+1+q of the group computes cut point q)
  int ca = 0;
#pragma unroll
  for (int i = 0; i <= D; ++i)
    if (PA[i * PAS + 2 * D] > 0.0) ca |= 1 << i;
  if (item > D && TC.inoutcut[ca] == GC_CUT) {
    int q = D + 1;
    for (int ed = 0; ed < TC.nedges; ++ed) {
      const int ia = TC.edges[ed][0], ib = TC.edges[ed][1];
      if (((ca >> ia) ^ (ca >> ib)) & 1) {
        if (q == item) {
          const double* p1 = PA + ia * PAS;
          const double* p2 = PA + ib * PAS;
          const double w1 = fabs(p1[2 * D]), w2 = fabs(p2[2 * D]);
          const double coeff = __ddiv_rn(w1, __dadd_rn(w1, w2));
          double* po = PA + q * PAS;
#pragma unroll
          for (int k = 0; k < PAS; ++k) po[k] = gc_lerp(p1[k], p2[k], coeff);  // x, r and BOTH pending values
        }
        q++;
      }
    }
  }
  __syncwarp();

  uint32_t n_cells = 0, n_pts = 0, n_fa = 0, n_fpa = 0, n_fb = 0, n_fpb = 0;  // counters of this item
  uint32_t cell_off = 0, pt_off = 0, fa_off = 0, fpa_off = 0, fb_off = 0, fpb_off = 0;
  // emission helpers (FILL only)
  auto put_facet = [&](uint32_t fid, uint32_t pb0, const uint8_t* lp, const double* nrm, const double (*pf)[3], int ls,
                       int k1, int8_t s1, int k2, int8_t s2) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      lay.sf_c2p[(int64_t)fid * D + i] = (int32_t)(pb0 + lp[i] + 1);
      lay.sf_N[(int64_t)fid * D + i] = nrm[i];
    }
    lay.sf_c2bg[fid] = bgcell1;
    lay.sf_meas[fid] = integrate_one<D - 1>(simplex_meas<D, D - 1>(pf));
    for (int kk = 0; kk < L; ++kk)
      lay.sf_state[(int64_t)kk * lay.sf_pitch + fid] =
          (kk == ls) ? (int8_t)GECUT_INTERFACE : (kk == k1 ? s1 : (kk == k2 ? s2 : cst(kk)));
  };
  // inclusive scan over the WALK_ITEMS lanes of the simplex
  auto seg_scan = [&](uint32_t v) -> uint32_t {
#pragma unroll
    for (int o = 1; o < WALK_ITEMS; o <<= 1) {
      const uint32_t up = __shfl_up_sync(FULLM, v, o);
      if ((lane & (WALK_ITEMS - 1)) >= o) v += up;
    }
    return v;
  };
  // The count launch leaves the six counters of every item (one byte each) in mu.two_items; the fill launch turns them
  // into offsets with the 8-lane scans and runs the body once, writing.  (The launches of one cut use the same slot
  // list, so gtid addresses the same item in both.)
  if (FILL) {
    const unsigned long long pk = mu.two_items[gtid];
    n_cells = (uint32_t)(pk & 0xFFu);
    n_pts = (uint32_t)((pk >> 8) & 0xFFu);
    n_fa = (uint32_t)((pk >> 16) & 0xFFu);
    n_fpa = (uint32_t)((pk >> 24) & 0xFFu);
    n_fb = (uint32_t)((pk >> 32) & 0xFFu);
    n_fpb = (uint32_t)((pk >> 40) & 0xFFu);
    const uint32_t ic = seg_scan(n_cells), ip = seg_scan(n_pts), ifa = seg_scan(n_fa), ipa = seg_scan(n_fpa),
                   ifb = seg_scan(n_fb), ipb = seg_scan(n_fpb);
    if (valid) {
      // start offsets of the simplex, left by the fill kernel of its tile (k_cutm_fill)
      const uint32_t* val = mu.slow_val + slot;  // channel ch at val[ch * slow_n]
      cell_off = val[0] + ic - n_cells;
      pt_off = val[mu.slow_n] + ip - n_pts;
      fa_off = mu.fac_base[a] + val[(int64_t)(2 + a) * mu.slow_n] + ifa - n_fa;
      fpa_off = mu.fpt_base[a] + val[(int64_t)(2 + L + a) * mu.slow_n] + ipa - n_fpa;
      fb_off = mu.fac_base[b] + val[(int64_t)(2 + b) * mu.slow_n] + ifb - n_fb;
      fpb_off = mu.fpt_base[b] + val[(int64_t)(2 + L + b) * mu.slow_n] + ipb - n_fpb;
    }
  }
  {
    constexpr bool wr = FILL;
    if (valid && item < (int)TC.nsub[ca]) {
      // ================= child `item` of stage a
      const int j = item;
      const int8_t sta = TC.sub_inout[ca][j];
      const uint32_t q1 = cm_skip_cell_stages<D>(s_plut, outm, a, b);
      // points of the child: its D+1 vertices are points of stage a (slots of PA, 4 bits each in vmap: the layouts of a PA
      // and a PBN row agree on x, r, wa), the cut points of stage b go to the lane's PBN rows
      double wbv[D + 1];
      uint32_t vmap = 0u;
#pragma unroll
      for (int i = 0; i <= D; ++i) {
        const uint32_t sl = TC.sub_pts[ca][j][(q1 >> (2 * i)) & 3u];
        vmap |= sl << (4 * i);
        wbv[i] = PA[sl * PAS + 2 * D + 1];
      }
      auto pbp = [&](int li) -> const double* {
        return li <= D ? PA + ((vmap >> (4 * li)) & 15u) * PAS : PBN + (li - (D + 1)) * PBS;
      };
      int cb = 0;
#pragma unroll
      for (int i = 0; i <= D; ++i)
        if (wbv[i] > 0.0) cb |= 1 << i;
      int np = D + 1;
      if (TC.inoutcut[cb] == GC_CUT) {
        for (int ed = 0; ed < TC.nedges; ++ed) {
          const int ia = TC.edges[ed][0], ib = TC.edges[ed][1];
          if (((cb >> ia) ^ (cb >> ib)) & 1) {
            double wa1 = wbv[0], wb1 = wbv[0];
#pragma unroll
            for (int m = 1; m <= D; ++m) {
              if (ia == m) wa1 = wbv[m];
              if (ib == m) wb1 = wbv[m];
            }
            const double w1 = fabs(wa1), w2 = fabs(wb1);
            const double coeff = __ddiv_rn(w1, __dadd_rn(w1, w2));
#pragma unroll
            for (int k = 0; k < PBS; ++k) PBN[(np - (D + 1)) * PBS + k] = gc_lerp(pbp(ia)[k], pbp(ib)[k], coeff);
            np++;
          }
        }
      }
      // ---- facets of stage b, re-cut by a (then constant stages down to klast)
      const int klast = b == 0 ? 1 : 0;
      const int nfb = TC.nfac[cb];
      uint32_t fo = fb_off, fpo = fpb_off;
      for (int f = 0; f < nfb; ++f) {
        const uint8_t* fp = TC.fac_pts[cb][f];
        double nrm[3] = {0.0, 0.0, 0.0};
        if (wr) unit_normal<D>(pbp(fp[0]), pbp(fp[1]), pbp(fp[D - 1]), (double)TC.fac_orient[cb][f], nrm);
        const uint32_t qf = cm_skip_facet_stages<D>(s_flut, outm, L, a, b);
        int fca = 0;
        double wav[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const double* p = pbp(fp[(qf >> (2 * i)) & 3u]);
#pragma unroll
          for (int k = 0; k < 2 * D; ++k) FE[i * 2 * D + k] = p[k];
          wav[i] = p[2 * D];
          if (wav[i] > 0.0) fca |= 1 << i;
        }
        int fn = D;
        if (TF.inoutcut[fca] == GC_CUT) {
          for (int ed = 0; ed < TF.nedges; ++ed) {
            const int ia = TF.edges[ed][0], ib = TF.edges[ed][1];
            if (((fca >> ia) ^ (fca >> ib)) & 1) {
              double v1 = wav[0], v2 = wav[0];
#pragma unroll
              for (int m = 1; m < D; ++m) {
                if (ia == m) v1 = wav[m];
                if (ib == m) v2 = wav[m];
              }
              const double w1 = fabs(v1), w2 = fabs(v2);
              const double coeff = __ddiv_rn(w1, __dadd_rn(w1, w2));
#pragma unroll
              for (int k = 0; k < 2 * D; ++k) FE[fn * 2 * D + k] = gc_lerp(FE[ia * 2 * D + k], FE[ib * 2 * D + k], coeff);
              fn++;
            }
          }
        }
        const int nsf = TF.nsub[fca];
        if (a == klast) {
          // emitted at stage a itself: one point block, nsf children
          if (wr) {
            for (int p = 0; p < fn; ++p)
#pragma unroll
              for (int d = 0; d < D; ++d) {
                lay.sf_X[(int64_t)(fpo + p) * D + d] = FE[p * 2 * D + d];
                lay.sf_R[(int64_t)(fpo + p) * D + d] = FE[p * 2 * D + D + d];
              }
            for (int jj = 0; jj < nsf; ++jj) {
              double pf[D][3];
#pragma unroll
              for (int i = 0; i < D; ++i)
#pragma unroll
                for (int d = 0; d < D; ++d) pf[i][d] = FE[TF.sub_pts[fca][jj][i] * 2 * D + d];
              put_facet(fo + jj, fpo, TF.sub_pts[fca][jj], nrm, pf, b, a, TF.sub_inout[fca][jj], -1, 0);
            }
          }
          fo += nsf;
          fpo += fn;
        } else {
          const uint32_t q2 = cm_skip_facet_stages<D>(s_flut, outm, a, klast, b);
          const int fc0 = ((outm >> klast) & 1u) ? fc_allout : 0;
          if (wr) {
            for (int jj = 0; jj < nsf; ++jj) {
              uint8_t fv[D];
#pragma unroll
              for (int i = 0; i < D; ++i) fv[i] = TF.sub_pts[fca][jj][(q2 >> (2 * i)) & 3u];
              const uint32_t pb0 = fpo + (uint32_t)jj * D;
              double pf[D][3];
#pragma unroll
              for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                  lay.sf_X[(int64_t)(pb0 + i) * D + d] = FE[fv[i] * 2 * D + d];
                  lay.sf_R[(int64_t)(pb0 + i) * D + d] = FE[fv[i] * 2 * D + D + d];
                  pf[i][d] = FE[fv[TF.sub_pts[fc0][0][i]] * 2 * D + d];
                }
              }
              put_facet(fo + jj, pb0, TF.sub_pts[fc0][0], nrm, pf, b, a, TF.sub_inout[fca][jj], klast, TF.sub_inout[fc0][0]);
            }
          }
          fo += nsf;
          fpo += (uint32_t)nsf * D;
        }
      }
      // ---- cells of this child
      const int nsb = TC.nsub[cb];
      uint32_t ncell_out, npt_out;
      if (b == 0) {
        if (wr) {
          for (int p = 0; p < np; ++p) {
            const double* src = pbp(p);
#pragma unroll
            for (int d = 0; d < D; ++d) {
              lay.sc_X[(int64_t)(pt_off + p) * D + d] = src[d];
              lay.sc_R[(int64_t)(pt_off + p) * D + d] = src[D + d];
            }
          }
          for (int jj = 0; jj < nsb; ++jj) {
            const uint32_t cid = cell_off + jj;
            int32_t row[D + 1];
            double pc[D + 1][3];
#pragma unroll
            for (int i = 0; i <= D; ++i) {
              const int li = TC.sub_pts[cb][jj][i];
              row[i] = (int32_t)(pt_off + li + 1);
#pragma unroll
              for (int d = 0; d < D; ++d) pc[i][d] = pbp(li)[d];
            }
            store_c2p_row<D>(lay.sc_c2p, cid, row);
            lay.sc_c2bg[cid] = bgcell1;
            lay.sc_meas[cid] = integrate_one<D>(simplex_meas<D, D>(pc));
            lay.sc_state[cid] = TC.sub_inout[cb][jj];
            for (int kk = 1; kk < L; ++kk) lay.sc_state[(int64_t)kk * lay.sc_pitch + cid] = kk == a ? sta : cst(kk);
          }
        }
        ncell_out = nsb;
        npt_out = np;
      } else {
        const uint32_t qc = cm_skip_cell_stages<D>(s_plut, outm, b, 0);
        const int c0 = (outm & 1u) ? c_allout : 0;
        if (wr) {
          for (int jj = 0; jj < nsb; ++jj) {
            uint8_t cv[D + 1];
#pragma unroll
            for (int i = 0; i <= D; ++i) cv[i] = TC.sub_pts[cb][jj][(qc >> (2 * i)) & 3u];
            const uint32_t pb0 = pt_off + (uint32_t)jj * (D + 1), cid = cell_off + jj;
            int32_t row[D + 1];
            double pc[D + 1][3];
#pragma unroll
            for (int i = 0; i <= D; ++i) {
#pragma unroll
              for (int d = 0; d < D; ++d) {
                lay.sc_X[(int64_t)(pb0 + i) * D + d] = pbp(cv[i])[d];
                lay.sc_R[(int64_t)(pb0 + i) * D + d] = pbp(cv[i])[D + d];
              }
              const int li = TC.sub_pts[c0][0][i];
              row[i] = (int32_t)(pb0 + li + 1);
#pragma unroll
              for (int d = 0; d < D; ++d) pc[i][d] = pbp(cv[li])[d];
            }
            store_c2p_row<D>(lay.sc_c2p, cid, row);
            lay.sc_c2bg[cid] = bgcell1;
            lay.sc_meas[cid] = integrate_one<D>(simplex_meas<D, D>(pc));
            lay.sc_state[cid] = TC.sub_inout[c0][0];
            const int8_t stb = TC.sub_inout[cb][jj];
            for (int kk = 1; kk < L; ++kk)
              lay.sc_state[(int64_t)kk * lay.sc_pitch + cid] = kk == a ? sta : (kk == b ? stb : cst(kk));
          }
        }
        ncell_out = nsb;
        npt_out = (uint32_t)nsb * (D + 1);
      }
      if (!FILL) {
        n_cells = ncell_out;
        n_pts = npt_out;
        n_fb = fo - fb_off;
        n_fpb = fpo - fpb_off;
      }
    } else if (valid && item >= 6 && item - 6 < (int)TC.nfac[ca]) {
      // ================= facet `item - 6` of stage a, re-cut by b (then constant stages down to stage 0)
      const int f = item - 6;
      const uint8_t* fp = TC.fac_pts[ca][f];
      double nrm[3] = {0.0, 0.0, 0.0};
      if (wr) unit_normal<D>(PA + fp[0] * PAS, PA + fp[1] * PAS, PA + fp[D - 1] * PAS, (double)TC.fac_orient[ca][f], nrm);
      const uint32_t qf = cm_skip_facet_stages<D>(s_flut, outm, L, b, a);
      int fcb = 0;
      double wv[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double* p = PA + fp[(qf >> (2 * i)) & 3u] * PAS;
#pragma unroll
        for (int k = 0; k < 2 * D; ++k) FE[i * 2 * D + k] = p[k];
        wv[i] = p[2 * D + 1];
        if (wv[i] > 0.0) fcb |= 1 << i;
      }
      int fn = D;
      if (TF.inoutcut[fcb] == GC_CUT) {
        for (int ed = 0; ed < TF.nedges; ++ed) {
          const int ia = TF.edges[ed][0], ib = TF.edges[ed][1];
          if (((fcb >> ia) ^ (fcb >> ib)) & 1) {
            double v1 = wv[0], v2 = wv[0];
#pragma unroll
            for (int m = 1; m < D; ++m) {
              if (ia == m) v1 = wv[m];
              if (ib == m) v2 = wv[m];
            }
            const double w1 = fabs(v1), w2 = fabs(v2);
            const double coeff = __ddiv_rn(w1, __dadd_rn(w1, w2));
#pragma unroll
            for (int k = 0; k < 2 * D; ++k) FE[fn * 2 * D + k] = gc_lerp(FE[ia * 2 * D + k], FE[ib * 2 * D + k], coeff);
            fn++;
          }
        }
      }
      const int nsf = TF.nsub[fcb];
      uint32_t nfo, npo;
      if (b == 0) {
        if (wr) {
          for (int p = 0; p < fn; ++p)
#pragma unroll
            for (int d = 0; d < D; ++d) {
              lay.sf_X[(int64_t)(fpa_off + p) * D + d] = FE[p * 2 * D + d];
              lay.sf_R[(int64_t)(fpa_off + p) * D + d] = FE[p * 2 * D + D + d];
            }
          for (int jj = 0; jj < nsf; ++jj) {
            double pf[D][3];
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
              for (int d = 0; d < D; ++d) pf[i][d] = FE[TF.sub_pts[fcb][jj][i] * 2 * D + d];
            put_facet(fa_off + jj, fpa_off, TF.sub_pts[fcb][jj], nrm, pf, a, 0, TF.sub_inout[fcb][jj], -1, 0);
          }
        }
        nfo = nsf;
        npo = fn;
      } else {
        const uint32_t q2 = cm_skip_facet_stages<D>(s_flut, outm, b, 0, a);
        const int fc0 = (outm & 1u) ? fc_allout : 0;
        if (wr) {
          for (int jj = 0; jj < nsf; ++jj) {
            uint8_t fv[D];
#pragma unroll
            for (int i = 0; i < D; ++i) fv[i] = TF.sub_pts[fcb][jj][(q2 >> (2 * i)) & 3u];
            const uint32_t pb0 = fpa_off + (uint32_t)jj * D;
            double pf[D][3];
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
              for (int d = 0; d < D; ++d) {
                lay.sf_X[(int64_t)(pb0 + i) * D + d] = FE[fv[i] * 2 * D + d];
                lay.sf_R[(int64_t)(pb0 + i) * D + d] = FE[fv[i] * 2 * D + D + d];
                pf[i][d] = FE[fv[TF.sub_pts[fc0][0][i]] * 2 * D + d];
              }
            }
            put_facet(fa_off + jj, pb0, TF.sub_pts[fc0][0], nrm, pf, a, b, TF.sub_inout[fcb][jj], 0, TF.sub_inout[fc0][0]);
          }
        }
        nfo = nsf;
        npo = (uint32_t)nsf * D;
      }
      if (!FILL) {
        n_fa = nfo;
        n_fpa = npo;
      }
    }
  }
  if (!FILL) {
    mu.two_items[gtid] = (unsigned long long)n_cells | ((unsigned long long)n_pts << 8) | ((unsigned long long)n_fa << 16) |
                         ((unsigned long long)n_fpa << 24) | ((unsigned long long)n_fb << 32) |
                         ((unsigned long long)n_fpb << 40);
    const uint32_t ic = seg_scan(n_cells), ip = seg_scan(n_pts), ifa = seg_scan(n_fa), ipa = seg_scan(n_fpa),
                   ifb = seg_scan(n_fb), ipb = seg_scan(n_fpb);
    if (valid && item == WALK_ITEMS - 1) {
      uint32_t* val = mu.slow_val + slot;  // channel ch at val[ch * slow_n]
      const int64_t tile = s / 32;
      val[0] = ic;
      val[mu.slow_n] = ip;
      val[(int64_t)(2 + a) * mu.slow_n] = ifa;
      val[(int64_t)(2 + L + a) * mu.slow_n] = ipa;
      val[(int64_t)(2 + b) * mu.slow_n] = ifb;
      val[(int64_t)(2 + L + b) * mu.slow_n] = ipb;
      atomicAdd(mu.tile + tile, ic);
      atomicAdd(mu.tile + mu.ntiles + tile, ip);
      if (ifa) atomicAdd(mu.tile + (int64_t)(2 + a) * mu.ntiles + tile, ifa);
      if (ipa) atomicAdd(mu.tile + (int64_t)(2 + L + a) * mu.ntiles + tile, ipa);
      if (ifb) atomicAdd(mu.tile + (int64_t)(2 + b) * mu.ntiles + tile, ifb);
      if (ipb) atomicAdd(mu.tile + (int64_t)(2 + L + b) * mu.ntiles + tile, ipb);
    }
  }
}

// ================================================================================================
// L == 1 fast path (one level set: BASELINE configs 2 and 3).  The refinement tree has depth one, so a stage-0 simplex
// needs no stack: case -> table row.  Blocks own tiles of CUT1_TILE consecutive stage-0 simplices:
//   k_cut1_count : per-tile totals (subcells, points, facets)                      -> exclusive scan over the tiles
//   k_cut1_fill  : recomputes the case, block-scans the three counters (packed in one word), stages the tile's output
//                  in shared memory and writes every array of the layout with fully coalesced stores.
// No per-simplex offset ever goes to HBM.  The facet point arrays of a single level set are the subcell point arrays
// (CutTriangulations.jl:365-378 aliases them; merge_sub_face_data copies them): on the device they stay aliased.
// ================================================================================================
constexpr int CUT1_TILE = 32;    // one warp = one tile: warps run their phases independently (no block barriers)
constexpr int CUT1_BLOCK = 128;  // threads per block of the fill kernel (4 independent tiles)
#ifndef CUT1_MINB
#define CUT1_MINB 6  // 24 warps per SM (80 registers, 34 KB of shared memory per block): 0.66 -> 0.60 ms on 512^3 tets vs 5
#endif

// shared memory of one warp of the fill kernel: staged points (+ phase pad), cut-edge coefficients, 16-byte records,
// subcell measures, owner maps; 32-byte multiple
template <int D, bool SX = false>
__host__ __device__ constexpr int cut1_warp_bytes() {
  // measure slots: simplex bg cells (closed-form cell measures) only park (IN, OUT) of the lane's cell, n-cube cells none
  return (((CUT1_TILE * ((D == 3) ? 8 : 5) * D + 4) * 8 + CUT1_TILE * ((D == 3) ? 4 : 2) * 8 + CUT1_TILE * 16 +
           (SX ? CUT1_TILE * 2 : 0) * 8 +
           CUT1_TILE * ((D == 3) ? 6 : 3) + CUT1_TILE * ((D == 3) ? 2 : 1)) + 31) / 32 * 32;
}

__device__ __forceinline__ uint32_t cut1_pack(int ncell, int npt, int nfac) {
  return (uint32_t)nfac | ((uint32_t)npt << 10) | ((uint32_t)ncell << 21);
}

// exclusive scan of one packed word per lane across the warp (= the tile); *total = tile sum
__device__ __forceinline__ uint32_t cut1_warp_scan(uint32_t v, uint32_t* total) {
  const int lane = threadIdx.x & 31;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  *total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// Tile totals.  The three counters of a stage-0 simplex depend only on HOW MANY of its vertices are OUT (the table
// rows of a k-vs-(D+1-k) split all have the same sizes), so one thread per cut bg cell evaluates the cell's nodes once,
// takes the popcount of each of its stage-0 simplices and adds the packed counters to the tile the simplex falls in.
// A block owns BC cells with BC*nt0 a multiple of CUT1_TILE, i.e. a whole number of tiles.
template <int D>
__global__ void __launch_bounds__(128)
k_cut1_count(const GcGrid g, const GcLeaves lv, const uint8_t* __restrict__ simp_raw, const uint8_t* __restrict__ simp_pos,
             const int32_t* __restrict__ cutcells, int64_t ncut, int cells_per_block, uint32_t* __restrict__ tile_cells,
             uint32_t* __restrict__ tile_points, uint32_t* __restrict__ tile_facets, int64_t ntiles) {
  __shared__ uint32_t s_tile[32];
  if (threadIdx.x < 32) s_tile[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t k = (int64_t)blockIdx.x * cells_per_block + threadIdx.x;
  const int tiles_per_block = cells_per_block * g.nt0 / CUT1_TILE;
  if ((int)threadIdx.x < cells_per_block && k < ncut) {
    // slab-local cell ids fit in 32 bits (checked at cut time)
    const uint32_t cell = (uint32_t)(cutcells[k] - 1);
    const uint32_t cube = cell / (uint32_t)g.cpc;
    const int typ = (int)(cell - cube * (uint32_t)g.cpc);
    int ijk[3];
    const uint32_t row = cube / (uint32_t)g.n[0];
    ijk[0] = (int)(cube - row * (uint32_t)g.n[0]);
    if (D == 3) {
      const uint32_t lay3 = row / (uint32_t)g.n[1];
      ijk[1] = (int)(row - lay3 * (uint32_t)g.n[1]);
      ijk[2] = (int)lay3 + g.k0;
    } else {
      ijk[1] = (int)row + g.k0;
      ijk[2] = 0;
    }
    // sign bits of the bg cell's vertices (local vertex ids of the n-cube)
    uint32_t bits = 0;
    const int nvb = g.simplexified ? D + 1 : (1 << D);
    for (int m = 0; m < nvb; ++m) {
      const int lvid = g.simplexified ? simp_raw[(D - 2) * 24 + typ * 4 + m] : m;
      const int nijk[3] = {ijk[0] + (lvid & 1), ijk[1] + ((lvid >> 1) & 1), ijk[2] + ((lvid >> 2) & 1)};
      double v;
      if (lv.leaf[0].kind == GECUT_LEAF_NODAL) {
        const int64_t node = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                                      : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
        v = __ldg(lv.nodal[0] + (node - lv.nodal_base));
      } else {
        const double xx[3] = {gc_coord(g, 0, nijk[0]), gc_coord(g, 1, nijk[1]), D == 3 ? gc_coord(g, 2, nijk[2]) : 0.0};
        v = gc_eval_leaf(lv.leaf[0], xx, D);
      }
      if (v > 0.0) bits |= 1u << m;
    }
    for (int t = 0; t < g.nt0; ++t) {
      int nout;
      if (g.simplexified) {
        nout = __popc(bits);
      } else {
        nout = 0;
        for (int m = 0; m <= D; ++m) nout += (bits >> simp_pos[(D - 2) * 24 + t * 4 + m]) & 1u;
      }
      const int kk = min(nout, D + 1 - nout);  // 0: uncut, 1: one vertex split off, 2: two-two split (TET only)
      const uint32_t pk = (D == 3) ? (kk == 0 ? cut1_pack(1, 4, 0) : kk == 1 ? cut1_pack(4, 7, 1) : cut1_pack(6, 8, 2))
                                   : (kk == 0 ? cut1_pack(1, 3, 0) : cut1_pack(3, 5, 1));
      const int64_t sidx = k * g.nt0 + t;
      atomicAdd(&s_tile[(int)(sidx / CUT1_TILE - (int64_t)blockIdx.x * tiles_per_block)], pk);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < tiles_per_block) {
    const int64_t tile = (int64_t)blockIdx.x * tiles_per_block + threadIdx.x;
    if (tile < ntiles) {
      const uint32_t total = s_tile[threadIdx.x];
      tile_cells[tile] = total >> 21;
      tile_points[tile] = (total >> 10) & 0x7FFu;
      tile_facets[tile] = total & 0x3FFu;
    }
  }
}

// warp copy of n doubles from the staging buffer to global memory: every instruction moves 512 contiguous bytes
// (16 per lane), which is conflict-free on the shared-memory side and whole sectors on the global side.  The caller
// staged element k at src[k] with src = stage + (global index of element 0 mod 4), so both sides share the 16-byte phase.
__device__ __forceinline__ void cut1_copy_doubles(double* __restrict__ dst, const double* __restrict__ src, uint32_t n,
                                                  uint32_t phase, int lane) {
  const uint32_t head = min(n, phase & 1u);  // element before the first 16-byte boundary
  if (head && lane == 0) dst[0] = src[0];
  const uint32_t nvec = (n - head) >> 1;
  const double2* s2 = reinterpret_cast<const double2*>(src + head);
  double2* d2 = reinterpret_cast<double2*>(dst + head);
#pragma unroll 2
  for (uint32_t i = lane; i < nvec; i += 32) d2[i] = s2[i];
  if (((n - head) & 1u) && lane == 0) dst[n - 1] = src[n - 1];
}

// The same copy handed to the TMA engine: one elected lane issues a bulk shared -> global copy of the 16-byte aligned body
// (cp.async.bulk, SASS UBLKCP), so the tile's points leave the staging buffer without passing through the LSU data pipe
// that bounds this kernel (ncu: 84 % of its peak; the LDS.128 + STG.128 pairs of the warp copy were 10 % of its shared-
// memory wavefronts and half of its global-store sectors).  Every lane fences its own staging writes towards the async
// proxy first; the caller waits with cut1_bulk_wait_read before the staging buffer is written again (or the warp exits).
__device__ __forceinline__ void cut1_bulk_copy_doubles(double* __restrict__ dst, const double* __restrict__ src, uint32_t n,
                                                       uint32_t phase, int lane) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    const uint32_t head = min(n, phase & 1u);  // element before the first 16-byte boundary
    if (head) dst[0] = src[0];
    const uint32_t nvec = (n - head) >> 1;
    if (nvec) {
      const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(src + head);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + head), "r"(saddr), "r"(nvec * 16u)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if ((n - head) & 1u) dst[n - 1] = src[n - 1];
  }
}
__device__ __forceinline__ void cut1_bulk_wait_read(int lane) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}

// edge e of the simplex in Gridap order: TRI [1,2],[1,3],[2,3]; TET [1,2],[1,3],[2,3],[1,4],[2,4],[3,4] (0-based ends)
template <int D>
__device__ __forceinline__ constexpr int cut1_edge_a(int e) {
  return (D == 3) ? (e == 0 ? 0 : e == 1 ? 0 : e == 2 ? 1 : e == 3 ? 0 : e == 4 ? 1 : 2) : (e == 0 ? 0 : e == 1 ? 0 : 1);
}
template <int D>
__device__ __forceinline__ constexpr int cut1_edge_b(int e) {
  return (D == 3) ? (e == 0 ? 1 : e == 1 ? 2 : e == 2 ? 2 : 3) : (e == 0 ? 1 : 2);
}

// Fill of the single-level-set layout.  One warp = one tile of 32 consecutive stage-0 simplices, four phases:
//   A  lane = simplex : stage-0 set-up (vertices, level-set values), case, warp scan of the three counters, edge
//                       coefficients, physical points into the staging buffer, one 16-byte record per simplex
//   X  warp           : point_to_coords of the tile leaves the staging buffer with 32-byte stores
//   C  lane = subcell : connectivity row, bg cell, IN/OUT state and measure, written directly (lane-contiguous)
//   D  lane = subfacet: connectivity row, unit normal, bg cell, measure
//   R  lane = simplex : reference points into the staging buffer (an edge point is 0, 1, c or 1-c per direction),
//                       then the same warp copy
// Phases C and D have no divergent trip counts: every lane owns one output entity and finds its simplex through a
// byte map written in phase A.
template <int D, bool SX, int MB, bool WM = true>
__global__ void __launch_bounds__(CUT1_BLOCK, MB)
k_cut1_fill(const GcGrid g, const GcLeaves lv, const GcSimplexTable* __restrict__ tables,
            const uint8_t* __restrict__ simp_raw, const uint8_t* __restrict__ simp_pos,
            const int32_t* __restrict__ cutcells, int64_t nsimp, const uint32_t* __restrict__ tile_cells,
            const uint32_t* __restrict__ tile_points, const uint32_t* __restrict__ tile_facets, GcLayout lay,
            uint32_t* __restrict__ sc_ptr, double4* __restrict__ tile_stats, double2* __restrict__ cell_vio) {
  constexpr int NT0 = SX ? 1 : (D == 3 ? 6 : 2);   // stage-0 simplices per bg cell
  constexpr int CPC = SX ? (D == 3 ? 6 : 2) : 1;   // bg cells per n-cube
  constexpr int NPC = WalkCfg<D>::NPC;
  constexpr int NSUB = (D == 3) ? 6 : 3, NFAC = (D == 3) ? 2 : 1, NE = (D == 3) ? 6 : 3;
  constexpr int STG = CUT1_TILE * NPC * D + 4;     // staged doubles per warp (+ phase pad)
  constexpr int WBYTES = cut1_warp_bytes<D, SX && !WM>();
  extern __shared__ __align__(32) unsigned char fill_smem[];
  __shared__ GcSimplexTable s_tab;
  __shared__ uint8_t s_swap[8];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tables + (D - 1));
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_tab);
    for (int i = threadIdx.x; i < (int)(sizeof(GcSimplexTable) / 4); i += CUT1_BLOCK) dst[i] = src[i];
  }
  const uint8_t* vtab = (SX ? simp_raw : simp_pos) + (D - 2) * 24;  // rows of 4 local vertex ids of the n-cube
  if (threadIdx.x < (SX ? CPC : NT0)) {
    // _ensure_positive_jacobians! (CutTriangulations.jl:512-514): on a Cartesian grid the sign of the stage-0 Jacobian
    // depends on the simplex type only (every entry of J is 0 or +-(one grid step)), so it is decided once per block
    double x[D + 1][D];
#pragma unroll
    for (int m = 0; m <= D; ++m) {
      const int lvid = vtab[threadIdx.x * 4 + m];
#pragma unroll
      for (int d = 0; d < D; ++d) x[m][d] = gc_coord(g, d, (d == D - 1 ? g.k0 : 0) + ((lvid >> d) & 1));
    }
    double J[3][3];
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = 0; b < D; ++b) J[a][b] = __dsub_rn(x[a + 1][b], x[0][b]);
    double dV;
    if (D == 2) {
      dV = __dsub_rn(__dmul_rn(J[0][0], J[1][1]), __dmul_rn(J[0][1], J[1][0]));
    } else {
      const double p1 = __dmul_rn(__dmul_rn(J[0][0], J[1][1]), J[2][2]);
      const double p2 = __dmul_rn(__dmul_rn(J[0][1], J[1][2]), J[2][0]);
      const double p3 = __dmul_rn(__dmul_rn(J[0][2], J[1][0]), J[2][1]);
      const double q1 = __dmul_rn(__dmul_rn(J[0][0], J[1][2]), J[2][1]);
      const double q2 = __dmul_rn(__dmul_rn(J[0][1], J[1][0]), J[2][2]);
      const double q3 = __dmul_rn(__dmul_rn(J[0][2], J[1][1]), J[2][0]);
      dV = __dsub_rn(__dadd_rn(__dadd_rn(p1, p2), p3), __dadd_rn(__dadd_rn(q1, q2), q3));
    }
    s_swap[threadIdx.x] = dV < 0.0 ? 1 : 0;
  }
  __syncthreads();
  const GcSimplexTable& TC = s_tab;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* wbase = fill_smem + (size_t)wid * WBYTES;
  double* stage = reinterpret_cast<double*>(wbase);
  constexpr int NCE = (D == 3) ? 4 : 2;  // a cut simplex has at most NCE cut edges
  double* cstash = stage + STG;          // [NCE][32] coefficients of the cut edges, parked here between phases A and R
  uint4* rec = reinterpret_cast<uint4*>(wbase + sizeof(double) * (STG + NCE * CUT1_TILE));
  double* vstage = reinterpret_cast<double*>(rec + CUT1_TILE);  // simplex bg cells: [2][32] (IN, OUT) measure of the lane's cell
  static_assert(!SX || !WM, "simplex bg cells use the closed-form cell measures");
  uint8_t* own_c = reinterpret_cast<uint8_t*>(vstage + CUT1_TILE * (SX ? 2 : 0));
  uint8_t* own_f = own_c + CUT1_TILE * NSUB;
  const int64_t tile = (int64_t)blockIdx.x * (CUT1_BLOCK / CUT1_TILE) + wid;
  if (tile * CUT1_TILE >= nsimp) return;  // whole warp beyond the end (no block barrier below)
  const int64_t s = tile * CUT1_TILE + lane;
  const bool valid = s < nsimp;

  // ======================================== phase A: lane = stage-0 simplex
  double X[D + 1][D], W[D + 1];
  uint32_t rv4 = 0u;  // byte m: reference coordinates of vertex m, one bit per direction
  int c = 0, nsub = 0, np = 0, nfac = 0;
  int32_t bgcell1 = 0;
  int64_t kcut = 0;
  int t0 = 0;
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    W[i] = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) X[i][d] = 0.0;
  }
  if (valid) {
    kcut = s / NT0;
    t0 = (int)(s - kcut * NT0);
    bgcell1 = cutcells[kcut];
    // slab-local cell ids fit in 32 bits (checked at cut time)
    const uint32_t cell = (uint32_t)(bgcell1 - 1);
    const uint32_t cube = cell / (uint32_t)CPC;
    const int row_id = SX ? (int)(cell - cube * (uint32_t)CPC) : t0;  // row of the vertex table
    int ijk[3];
    const uint32_t row = cube / (uint32_t)g.n[0];
    ijk[0] = (int)(cube - row * (uint32_t)g.n[0]);
    if (D == 3) {
      const uint32_t lay3 = row / (uint32_t)g.n[1];
      ijk[1] = (int)(row - lay3 * (uint32_t)g.n[1]);
      ijk[2] = (int)lay3 + g.k0;
    } else {
      ijk[1] = (int)row + g.k0;
      ijk[2] = 0;
    }
    uint32_t lv4 = *reinterpret_cast<const uint32_t*>(vtab + row_id * 4);
    rv4 = SX ? 0x04020100u : lv4;  // simplex cells: origin + unit axes; n-cube cells: the cube vertex itself
    if (s_swap[row_id]) {          // swap the first two vertices
      lv4 = __byte_perm(lv4, 0u, 0x3201);
      rv4 = __byte_perm(rv4, 0u, 0x3201);
    }
    const double RR = __dmul_rn(lv.leaf[0].R, lv.leaf[0].R);
#pragma unroll
    for (int m = 0; m <= D; ++m) {
      const int lvid = (lv4 >> (8 * m)) & 0xFFu;
      const int nijk[3] = {ijk[0] + (lvid & 1), ijk[1] + ((lvid >> 1) & 1), ijk[2] + ((lvid >> 2) & 1)};
      double xx[3] = {gc_coord(g, 0, nijk[0]), gc_coord(g, 1, nijk[1]), D == 3 ? gc_coord(g, 2, nijk[2]) : 0.0};
#pragma unroll
      for (int d = 0; d < D; ++d) X[m][d] = xx[d];
      const int kind = lv.leaf[0].kind;
      if (kind == GECUT_LEAF_NODAL) {
        const int64_t node = (D == 3) ? ((int64_t)nijk[0] + (int64_t)g.nn[0] * ((int64_t)nijk[1] + (int64_t)g.nn[1] * nijk[2]))
                                      : ((int64_t)nijk[0] + (int64_t)g.nn[0] * nijk[1]);
        W[m] = __ldg(lv.nodal[0] + (node - lv.nodal_base));
      } else if (kind == GECUT_LEAF_SPHERE) {  // gc_eval_leaf with R*R hoisted
        double w[3] = {__dsub_rn(xx[0], lv.leaf[0].x0[0]), __dsub_rn(xx[1], lv.leaf[0].x0[1]),
                       D == 3 ? __dsub_rn(xx[2], lv.leaf[0].x0[2]) : 0.0};
        W[m] = __dsub_rn(gc_dot3(w, w, D), RR);
      } else {
        W[m] = gc_eval_leaf(lv.leaf[0], xx, D);
      }
    }
#pragma unroll
    for (int i = 0; i <= D; ++i)
      if (W[i] > 0.0) c |= 1 << i;
    nsub = TC.nsub[c];
    np = TC.npts[c];
    nfac = TC.nfac[c];
  }
  uint32_t total;
  const uint32_t excl = cut1_warp_scan(valid ? cut1_pack(nsub, np, nfac) : 0u, &total);
  const uint32_t lc = excl >> 21, lp = (excl >> 10) & 0x7FFu, lf = excl & 0x3FFu;    // tile-local offsets
  const uint32_t tc = total >> 21, tp = (total >> 10) & 0x7FFu, tf = total & 0x3FFu;  // tile totals
  // tile bases.  n-cube cells: from the count pass + scan.  Simplex cells: every listed cell is cut, so the sizes are linear
  // in the number of cells (c0 = 32*tile) and of 2-2 TET splits (tile_cells[tile], written by k_cutmask_emit) before the tile
  uint32_t gc, gp, gf;
  if (SX) {
    const uint32_t c0 = (uint32_t)(tile * CUT1_TILE);
    if (D == 3) {
      const uint32_t b = tile_cells[tile];
      gc = 4u * c0 + 2u * b;
      gp = 7u * c0 + b;
      gf = c0 + b;
    } else {
      gc = 3u * c0;
      gp = 5u * c0;
      gf = c0;
    }
  } else {
    gc = tile_cells[tile];
    gp = tile_points[tile];
    gf = tile_facets[tile];
  }
  // staged doubles start at `phase` so that shared and global memory have the same 32-byte phase (vector copy-out)
  const uint32_t phase = (uint32_t)(((uint64_t)gp * D) & 3u);
  double* stage_p = stage + phase;
  double coeff[NE];
  const bool iscut = valid && TC.inoutcut[c] == GC_CUT;
#pragma unroll
  for (int e = 0; e < NE; ++e) {
    const double v1 = W[cut1_edge_a<D>(e)], v2 = W[cut1_edge_b<D>(e)];
    coeff[e] = -1.0;  // edge not cut
    if (iscut && ((v1 > 0.0) != (v2 > 0.0))) {
      const double w1 = fabs(v1), w2 = fabs(v2);
      coeff[e] = __ddiv_rn(w1, __dadd_rn(w1, w2));  // (CutTriangulations.jl:213-215)
    }
  }
  // Simplex bg cells without per-entity measures (ANALYTIC): the IN / OUT measure of the cut cell in closed form from the
  // cut fractions -- the crossing points are the zeros of the LINEAR interpolant of the vertex values, so the pieces are a
  // corner simplex + its complement (1 vertex split off) or two wedges (2-2 split) whose barycentric volumes are products of
  // the fractions; every term is a product of numbers in [0, 1] and only sums occur (no cancellation), so the result is
  // within a few ulp of the exact measure of the piece -- the sum of the subcell measures the reference forms differs from
  // it by the rounding of its own determinants (same order).  Saves the gather of 4 staged points per subcell.
  constexpr bool ANALYTIC = SX && !WM;
  double vi = 0.0, vo = 0.0;
  if (ANALYTIC && valid) {
    double x3[D + 1][3];
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
      for (int d = 0; d < 3; ++d) x3[i][d] = d < D ? X[i][d < D ? d : 0] : 0.0;
    const double m0 = integrate_one<D>(simplex_meas<D, D>(x3));
    const int nout = __popc((unsigned)c);
    if (!iscut) {
      if (nout == 0) vi = m0; else vo = m0;
    } else if (D == 2 || nout != 2) {
      // one vertex `iso` on its own side: fractions f_k from iso along its D cut edges; corner = prod f_k,
      // rest = 1 - prod f_k = (1 - f_1) + f_1 (1 - f_2) + f_1 f_2 (1 - f_3)
      const bool iso_out = nout == 1;
      const int iso = __ffs(iso_out ? c : (~c & ((1 << (D + 1)) - 1))) - 1;
      double prod = 1.0, rest = 0.0;
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (coeff[e] >= 0.0) {
          const double cf = coeff[e], omc = __dsub_rn(1.0, cf);
          const bool from_a = cut1_edge_a<D>(e) == iso;
          const double f = from_a ? cf : omc, g1 = from_a ? omc : cf;
          rest += prod * g1;
          prod *= f;
        }
      }
      const double corner = m0 * prod, other = m0 * rest;
      vi = iso_out ? other : corner;
      vo = iso_out ? corner : other;
    } else {
      // 2-2 split: OUT vertices p < q, IN vertices r < s; fractions from the OUT end: a = p->r, b = p->s, c = q->r, d = q->s.
      // OUT wedge = (p,q,Ppr,Pps) + (q,Ppr,Pps,Pqs) + (q,Ppr,Pqs,Pqr) = ab + ad(1-b) + (1-a)cd, IN wedge by symmetry
      const int pp = __ffs(c) - 1, rr = __ffs(~c & 15) - 1;
      double fa = 0.0, fb = 0.0, fc = 0.0, fd = 0.0, ga = 0.0, gb = 0.0, gc_ = 0.0, gd = 0.0;
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        if (coeff[e] >= 0.0) {
          const int ea = cut1_edge_a<D>(e), eb = cut1_edge_b<D>(e);
          const bool a_out = (c >> ea) & 1;
          const int vo_ = a_out ? ea : eb, vi_ = a_out ? eb : ea;
          const double cf = coeff[e], omc = __dsub_rn(1.0, cf);
          const double f = a_out ? cf : omc, g1 = a_out ? omc : cf;
          const int role = (vo_ == pp ? 0 : 2) + (vi_ == rr ? 0 : 1);
          if (role == 0) { fa = f; ga = g1; }
          else if (role == 1) { fb = f; gb = g1; }
          else if (role == 2) { fc = f; gc_ = g1; }
          else { fd = f; gd = g1; }
        }
      }
      vo = m0 * (fa * fb + fa * fd * gb + ga * fc * fd);
      vi = m0 * (ga * gc_ + ga * gd * fc + fa * gb * gd);
    }
  }
  if (ANALYTIC) {  // parked in the (otherwise unused) subcell-measure slots until the end of the tile: lane-private
    vstage[lane] = vi;
    vstage[CUT1_TILE + lane] = vo;
  }
  if (valid) {
    if (t0 == 0) {  // first subcell / first point of this cut cell
      sc_ptr[2 * kcut] = gc + lc;
      sc_ptr[2 * kcut + 1] = gp + lp;
    }
    if (s == nsimp - 1) {
      sc_ptr[2 * kcut + 2] = gc + lc + nsub;
      sc_ptr[2 * kcut + 3] = gp + lp + np;
    }
    // physical coordinates: vertex copies, then one point per cut edge in edge order (:199-219)
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
      for (int d = 0; d < D; ++d) stage_p[(lp + i) * D + d] = X[i][d];
    int q = D + 1;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      if (coeff[e] >= 0.0) {
#pragma unroll
        for (int d = 0; d < D; ++d)
          stage_p[(lp + q) * D + d] = gc_lerp(X[cut1_edge_a<D>(e)][d], X[cut1_edge_b<D>(e)][d], coeff[e]);
        cstash[(q - (D + 1)) * CUT1_TILE + lane] = coeff[e];
        q++;
      }
    }
    rec[lane] = make_uint4((uint32_t)c | (lp << 4) | (lf << 16), lc, gp + lp + 1u, (uint32_t)bgcell1);
#pragma unroll
    for (int j = 0; j < NSUB; ++j)
      if (j < nsub) own_c[lc + j] = (uint8_t)lane;
#pragma unroll
    for (int f = 0; f < NFAC; ++f)
      if (f < nfac) own_f[lf + f] = (uint8_t)lane;
  }
  // ======================================== phase X: point_to_coords of the tile (asynchronous: the staged points are only
  // read until phase R)
  cut1_bulk_copy_doubles(lay.sc_X + (int64_t)gp * D, stage_p, tp * D, phase, lane);
  // ======================================== phase C: lane = subcell
  // tile statistics for the selections of a single-leaf geometry (no pass over the subcells later): measure of the IN and
  // of the OUT subcells, measure of the subfacets, number of IN subcells
  double st_in = 0.0, st_out = 0.0, st_surf = 0.0;
  uint32_t st_nin = 0u;
  for (uint32_t idx = lane; idx < tc; idx += CUT1_TILE) {
    const uint4 r = rec[own_c[idx]];
    const int cc = r.x & 15u;
    const uint32_t plo = (r.x >> 4) & 0xFFFu;
    const int j = (int)(idx - r.y);
    const uint32_t pts = *reinterpret_cast<const uint32_t*>(TC.sub_pts[cc][j]);  // D+1 local point ids
    const int64_t o = (int64_t)gc + idx;
    if (D == 3) {
      reinterpret_cast<int4*>(lay.sc_c2p)[o] =
          make_int4((int)(r.z + (pts & 0xFFu)), (int)(r.z + ((pts >> 8) & 0xFFu)), (int)(r.z + ((pts >> 16) & 0xFFu)),
                    (int)(r.z + (pts >> 24)));
    } else {
#pragma unroll
      for (int v = 0; v <= D; ++v) lay.sc_c2p[o * (D + 1) + v] = (int)(r.z + ((pts >> (8 * v)) & 0xFFu));
    }
    lay.sc_c2bg[o] = (int32_t)r.w;
    const int8_t io = TC.sub_inout[cc][j];
    lay.sc_state[o] = io;
    if (ANALYTIC) {
      if (io == GC_IN) st_nin++;
    } else {
      double pc[D + 1][3];
#pragma unroll
      for (int v = 0; v <= D; ++v) {
        const double* pv = stage_p + (plo + ((pts >> (8 * v)) & 0xFFu)) * D;
#pragma unroll
        for (int d = 0; d < D; ++d) pc[v][d] = pv[d];
      }
      const double ms = integrate_one<D>(simplex_meas<D, D>(pc));
      if (WM) lay.sc_meas[o] = ms;
      if (io == GC_IN) {
        st_in += ms;
        st_nin++;
      } else {
        st_out += ms;
      }
    }
  }
  // ======================================== phase D: lane = subfacet
  for (uint32_t idx = lane; idx < tf; idx += CUT1_TILE) {
    const uint4 r = rec[own_f[idx]];
    const int cc = r.x & 15u;
    const uint32_t plo = (r.x >> 4) & 0xFFFu;
    const int f = (int)(idx - (r.x >> 16));
    const uint8_t* fp = TC.fac_pts[cc][f];
    double pf[D][3];
#pragma unroll
    for (int v = 0; v < D; ++v)
#pragma unroll
      for (int d = 0; d < D; ++d) pf[v][d] = stage_p[(plo + fp[v]) * D + d];
    double nrm[3];
    unit_normal<D>(pf[0], pf[1], pf[D - 1], (double)TC.fac_orient[cc][f], nrm);
    const int64_t o = (int64_t)gf + idx;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      lay.sf_c2p[o * D + d] = (int)(r.z + fp[d]);
      lay.sf_N[o * D + d] = nrm[d];
    }
    lay.sf_c2bg[o] = (int32_t)r.w;
    lay.sf_state[o] = (int8_t)GECUT_INTERFACE;
    const double mf = integrate_one<D - 1>(simplex_meas<D, D - 1>(pf));
    if (WM) lay.sf_meas[o] = mf;
    st_surf += mf;
  }
  if (ANALYTIC) {  // lane = simplex
    st_in = vstage[lane];
    st_out = vstage[CUT1_TILE + lane];
    if (valid) cell_vio[s] = make_double2(st_in, st_out);
  }
  // fixed lane assignment + fixed shuffle tree: the tile sums are reproducible
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    st_in += __shfl_xor_sync(0xffffffffu, st_in, o);
    st_out += __shfl_xor_sync(0xffffffffu, st_out, o);
    st_surf += __shfl_xor_sync(0xffffffffu, st_surf, o);
    st_nin += __shfl_xor_sync(0xffffffffu, st_nin, o);
  }
  if (lane == 0) tile_stats[tile] = make_double4(st_in, st_out, st_surf, (double)st_nin);
  __syncwarp();
  // ======================================== phase R: reference coordinates (lane = simplex), same staging buffer
  cut1_bulk_wait_read(lane);  // the physical points have left the staging buffer
  if (valid) {
#pragma unroll
    for (int i = 0; i <= D; ++i)
#pragma unroll
      for (int d = 0; d < D; ++d) stage_p[(lp + i) * D + d] = (double)((rv4 >> (8 * i + d)) & 1u);
    int q = D + 1;
    const bool iscut_r = TC.inoutcut[c] == GC_CUT;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      if (iscut_r && (((c >> cut1_edge_a<D>(e)) ^ (c >> cut1_edge_b<D>(e))) & 1)) {  // edge e is cut
        // gc_lerp(ra, rb, c) with ra, rb in {0, 1}: ra (equal ends), c (0 -> 1) or 1 - c (1 -> 0), bit for bit
        const double cf = cstash[(q - (D + 1)) * CUT1_TILE + lane], omc = __dsub_rn(1.0, cf);
        const uint32_t ra = (rv4 >> (8 * cut1_edge_a<D>(e))) & 0xFFu, rb = (rv4 >> (8 * cut1_edge_b<D>(e))) & 0xFFu;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const uint32_t a = (ra >> d) & 1u, b = (rb >> d) & 1u;
          stage_p[(lp + q) * D + d] = a == b ? (double)a : (a ? omc : cf);
        }
        q++;
      }
    }
  }
  cut1_bulk_copy_doubles(lay.sc_R + (int64_t)gp * D, stage_p, tp * D, phase, lane);
  cut1_bulk_wait_read(lane);  // the block's shared memory must stay until the engine has read it
}

template <int D, int NW>
void launch_walk(gecut_ctx* ctx, const gecut_result* res, const GcMulti& mu, bool fill, int64_t nslots,
                 const uint32_t* slots, int64_t items_base, cudaStream_t stream) {
  const int T = WALK_T;
  const unsigned nb = (unsigned)gc_div_up(nslots * WALK_ITEMS, T);
  if (nb == 0) return;
  if (fill)
    GC_LAUNCH(ctx, "k_cut_walk_fill", k_cut_walk<D, NW, true><<<nb, T, 0, stream>>>(res->grid, res->leaves, ctx->d_tables, ctx->d_simplexify_raw,
                                                       ctx->d_simplexify_pos, res->lay.cutcells, mu, res->lay, slots, nslots, items_base));
  else
    GC_LAUNCH(ctx, "k_cut_walk_count", k_cut_walk<D, NW, false><<<nb, T, 0, stream>>>(res->grid, res->leaves, ctx->d_tables, ctx->d_simplexify_raw,
                                                        ctx->d_simplexify_pos, res->lay.cutcells, mu, res->lay, slots, nslots, items_base));
}

void launch_fast_count(gecut_ctx* ctx, const gecut_result* res, int64_t nsimp, const GcMulti& mu) {
  if (!res->grid.simplexified && !getenv("GECUT_COUNT_PER_SIMPLEX")) {  // n-cube cells: thread per cut cell, then the tile totals
    const int64_t ncut = res->sizes.num_cutcells;
    const unsigned nbc = (unsigned)gc_div_up(ncut, 128), nbt = (unsigned)gc_div_up(mu.ntiles, 256 / 32);
    if (res->grid.D == 2) {
      GC_LAUNCH(ctx, "k_cutm_count", k_cutm_count_cells<2><<<nbc, 128, 0, ctx->stream>>>(res->grid, res->leaves, ctx->d_tables,
                                                                                        ctx->d_simplexify_pos, res->lay.cutcells, ncut, mu));
      GC_LAUNCH(ctx, "k_cutm_tiles", k_cutm_tiles<2><<<nbt, 256, 0, ctx->stream>>>(ctx->d_tables, res->L, nsimp, mu));
    } else {
      GC_LAUNCH(ctx, "k_cutm_count", k_cutm_count_cells<3><<<nbc, 128, 0, ctx->stream>>>(res->grid, res->leaves, ctx->d_tables,
                                                                                        ctx->d_simplexify_pos, res->lay.cutcells, ncut, mu));
      GC_LAUNCH(ctx, "k_cutm_tiles", k_cutm_tiles<3><<<nbt, 256, 0, ctx->stream>>>(ctx->d_tables, res->L, nsimp, mu));
    }
    return;
  }
  const unsigned nb = (unsigned)gc_div_up(mu.ntiles, CM_T / 32);
  if (res->grid.D == 2)
    GC_LAUNCH(ctx, "k_cutm_count", k_cutm_count<2><<<nb, CM_T, 0, ctx->stream>>>(res->grid, res->leaves, ctx->d_tables, ctx->d_simplexify_raw,
                                                                               ctx->d_simplexify_pos, res->lay.cutcells, nsimp, mu));
  else
    GC_LAUNCH(ctx, "k_cutm_count", k_cutm_count<3><<<nb, CM_T, 0, ctx->stream>>>(res->grid, res->leaves, ctx->d_tables, ctx->d_simplexify_raw,
                                                                               ctx->d_simplexify_pos, res->lay.cutcells, nsimp, mu));
}

void launch_fast_fill(gecut_ctx* ctx, const gecut_result* res, int64_t nsimp, const GcMulti& mu) {
  const unsigned nb = (unsigned)gc_div_up(mu.ntiles, CMF_T / 32);
  // per device (a process may drive several): set on every launch, it is a host-side table write
  if (res->grid.D == 2)
    cudaFuncSetAttribute(k_cutm_fill<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(CmSmem<2>) * (CMF_T / 32)));
  else
    cudaFuncSetAttribute(k_cutm_fill<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(CmSmem<3>) * (CMF_T / 32)));
  if (res->grid.D == 2)
    GC_LAUNCH(ctx, "k_cutm_fill", k_cutm_fill<2><<<nb, CMF_T, sizeof(CmSmem<2>) * (CMF_T / 32), ctx->stream>>>(
        res->grid, res->leaves, ctx->d_tables, ctx->d_simplexify_raw, ctx->d_simplexify_pos, res->lay.cutcells, nsimp, mu, res->lay,
        res->cut_sc_ptr));
  else
    GC_LAUNCH(ctx, "k_cutm_fill", k_cutm_fill<3><<<nb, CMF_T, sizeof(CmSmem<3>) * (CMF_T / 32), ctx->stream>>>(
        res->grid, res->leaves, ctx->d_tables, ctx->d_simplexify_raw, ctx->d_simplexify_pos, res->lay.cutcells, nsimp, mu, res->lay,
        res->cut_sc_ptr));
}

void launch_two(gecut_ctx* ctx, const gecut_result* res, const GcMulti& mu, bool fill, int64_t nslots, const uint32_t* slots,
                cudaStream_t stream) {
  const unsigned nb = (unsigned)gc_div_up(nslots * WALK_ITEMS, TWO_T);
  if (nb == 0) return;
#define GC_TWO(DD, FF, NAME)                                                                                              \
  do {                                                                                                                    \
    /* per device (a process may drive several): set on every launch, it is a host-side table write */                   \
    cudaFuncSetAttribute(k_cut_two<DD, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TwoSmem<DD>));         \
    GC_LAUNCH(ctx, NAME, k_cut_two<DD, FF><<<nb, TWO_T, sizeof(TwoSmem<DD>), stream>>>(                                      \
                             res->grid, res->leaves, ctx->d_tables, ctx->d_simplexify_raw, ctx->d_simplexify_pos,           \
                             res->lay.cutcells, mu, res->lay, slots, nslots));                                              \
  } while (0)
  if (res->grid.D == 2) {
    if (fill) GC_TWO(2, true, "k_cut_two_fill"); else GC_TWO(2, false, "k_cut_two_count");
  } else {
    if (fill) GC_TWO(3, true, "k_cut_two_fill"); else GC_TWO(3, false, "k_cut_two_count");
  }
#undef GC_TWO
}

// one dense launch per bucket of recorded simplices (nslow[b] slots starting at lists + base[b]): bucket 0 = the depth-2
// kernel, 1 + b = the generic walk tracking walk_nw(b) level sets.  The buckets are independent and each launch lasts as
// long as its slowest thread, so they run concurrently on side streams.
int dispatch_walk(gecut_ctx* ctx, const gecut_result* res, const GcMulti& mu, bool fill, const int64_t* nslow,
                  const uint32_t* lists) {
  const int D = res->grid.D;
  int nonempty = 0;
  for (int b = 0; b < CM_NBK; ++b) nonempty += nslow[b] > 0 ? 1 : 0;
  if (nonempty == 0) return GECUT_OK;
  const bool fork = nonempty > 1 && !ctx->prof_on;  // per-kernel profiling times launches on the main stream
  if (fork) {
    if (!ctx->ev_fork) GC_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    for (int b = 0; b < CM_NBK; ++b) {
      if (!ctx->side[b]) GC_CUDA(cudaStreamCreateWithFlags(&ctx->side[b], cudaStreamNonBlocking));
      if (!ctx->ev_join[b]) GC_CUDA(cudaEventCreateWithFlags(&ctx->ev_join[b], cudaEventDisableTiming));
    }
    GC_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  }
  auto stream_of = [&](int b) -> cudaStream_t {
    if (!fork || nslow[b] == 0) return ctx->stream;
    cudaStreamWaitEvent(ctx->side[b], ctx->ev_fork, 0);
    return ctx->side[b];
  };
  int64_t base[CM_NBK];
  base[0] = 0;
  for (int b = 1; b < CM_NBK; ++b) base[b] = base[b - 1] + nslow[b - 1];
  launch_two(ctx, res, mu, fill, nslow[0], lists + base[0], stream_of(0));
#define GC_WALK(DD, B) \
  launch_walk<DD, walk_nw(B)>(ctx, res, mu, fill, nslow[1 + B], lists + base[1 + B], (base[1 + B] - base[1]) * WALK_ITEMS, stream_of(1 + B))
  if (D == 2) {
    GC_WALK(2, 0);
    GC_WALK(2, 1);
    GC_WALK(2, 2);
    GC_WALK(2, 3);
    GC_WALK(2, 4);
  } else {
    GC_WALK(3, 0);
    GC_WALK(3, 1);
    GC_WALK(3, 2);
    GC_WALK(3, 3);
    GC_WALK(3, 4);
  }
#undef GC_WALK
  if (fork) {
    for (int b = 0; b < CM_NBK; ++b) {
      if (nslow[b] == 0) continue;
      GC_CUDA(cudaEventRecord(ctx->ev_join[b], ctx->side[b]));
      GC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[b], 0));
    }
  }
  return GECUT_OK;
}

}  // namespace

// deterministic sum of the tile statistics: fixed element -> thread -> tree order inside a block, the block that finishes
// last adds the block partials in index order
__global__ void __launch_bounds__(256)
k_tile_stats_reduce(const double4* __restrict__ tile_stats, int64_t ntiles, double4* __restrict__ partial,
                    unsigned int* __restrict__ ticket, double4* __restrict__ out) {
  __shared__ double4 sh[256];
  __shared__ bool s_last;
  double4 a = make_double4(0.0, 0.0, 0.0, 0.0);
  const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi = lo + per < ntiles ? lo + per : ntiles;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const double4 v = tile_stats[i];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      double4 u = sh[threadIdx.x], v = sh[threadIdx.x + o];
      u.x += v.x; u.y += v.y; u.z += v.z; u.w += v.w;
      sh[threadIdx.x] = u;
    }
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
  double4 t = make_double4(0.0, 0.0, 0.0, 0.0);
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    const volatile double* pv = reinterpret_cast<const volatile double*>(partial + i);
    t.x += pv[0]; t.y += pv[1]; t.z += pv[2]; t.w += pv[3];
  }
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      double4 u = sh[threadIdx.x], v = sh[threadIdx.x + o];
      u.x += v.x; u.y += v.y; u.z += v.z; u.w += v.w;
      sh[threadIdx.x] = u;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

static inline int64_t pad_pitch(int64_t n) { return ((n + 255) / 256) * 256; }

// Per-entity measures on demand.  Simplex bg cells with one level set: everything the default consumers need (the three
// global sums, the (IN, OUT) measure of every cut cell) is summed by the fill kernel while the vertices are on chip, so the
// 8 bytes per subcell / subfacet are not written at cut time; whoever asks for per-entity values (gecut_integrate_measure
// with `values`, non-leaf selections, get_cell_measure) gets them from this pass -- the same simplex_meas / integrate_one
// on the same stored coordinates, hence the same bits as the fill kernel would have written.
template <int D, int DS>
__global__ void __launch_bounds__(256)
k_entity_measures(const int32_t* __restrict__ c2p, const double* __restrict__ X, int64_t n, double* __restrict__ meas) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double pc[DS + 1][3];
#pragma unroll
  for (int v = 0; v <= DS; ++v) {
    const int64_t pid = (int64_t)c2p[e * (DS + 1) + v] - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) pc[v][d] = d < D ? X[pid * D + d] : 0.0;
  }
  meas[e] = integrate_one<DS>(simplex_meas<D, DS>(pc));
}

int gc_ensure_measures(gecut_ctx* ctx, const gecut_result* res_c) {
  gecut_result* res = const_cast<gecut_result*>(res_c);
  GcLayout& lay = res->lay;
  const int D = res->grid.D;
  const int64_t nsc = res->sizes.num_subcells, nsf = res->sizes.num_subfacets;
  const int T = 256;
  if (!lay.sc_meas && nsc > 0) {
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_meas, sizeof(double) * nsc));
    const unsigned nb = (unsigned)gc_div_up(nsc, T);
    if (D == 2)
      GC_LAUNCH(ctx, "k_entity_measures", k_entity_measures<2, 2><<<nb, T, 0, ctx->stream>>>(lay.sc_c2p, lay.sc_X, nsc, lay.sc_meas));
    else
      GC_LAUNCH(ctx, "k_entity_measures", k_entity_measures<3, 3><<<nb, T, 0, ctx->stream>>>(lay.sc_c2p, lay.sc_X, nsc, lay.sc_meas));
  }
  if (!lay.sf_meas && nsf > 0) {
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_meas, sizeof(double) * nsf));
    const unsigned nb = (unsigned)gc_div_up(nsf, T);
    if (D == 2)
      GC_LAUNCH(ctx, "k_entity_measures", k_entity_measures<2, 1><<<nb, T, 0, ctx->stream>>>(lay.sf_c2p, lay.sf_X, nsf, lay.sf_meas));
    else
      GC_LAUNCH(ctx, "k_entity_measures", k_entity_measures<3, 2><<<nb, T, 0, ctx->stream>>>(lay.sf_c2p, lay.sf_X, nsf, lay.sf_meas));
  }
  return gc_launch_check(ctx, "k_entity_measures");
}

// L == 1: tile totals -> scan over tiles -> allocate -> fill (see the fast-path kernels above)
static int gc_cut_cells_l1(gecut_ctx* ctx, gecut_result* res) {
  const GcGrid& g = res->grid;
  const int D = g.D;
  const int64_t ncut = res->sizes.num_cutcells;
  const int64_t nsimp = ncut * g.nt0;
  const int64_t ntiles = gc_div_up(nsimp, CUT1_TILE);
  gecut_sizes& sz = res->sizes;
  GcScratch scratch(ctx);  // tiles, totals, tile statistics: released on every exit path
  uint32_t* tiles = nullptr;
  uint64_t* totals_dev = nullptr;
  uint32_t *t_cells = nullptr, *t_points = nullptr, *t_facets = nullptr;
  uint64_t nsc = 0, nscp = 0, nsf = 0;
  if (g.simplexified) {
    // every listed cell is cut: TRI -> (3, 5, 1); TET -> (4, 7, 1), or (6, 8, 2) for the n22 cells that split 2-2 (counted by
    // k_classify).  No count pass, no scan, no read-back; the tile bases come from k_cutmask_emit.
    if (D == 3 && !res->want_n22) {
      ctx->err = "internal: the 2-2 split count of the classification is missing";
      return GECUT_ERR_INVALID;
    }
    const uint64_t nc = (uint64_t)ncut, b = (uint64_t)res->n22;
    nsc = D == 3 ? 4 * nc + 2 * b : 3 * nc;
    nscp = D == 3 ? 7 * nc + b : 5 * nc;
    nsf = D == 3 ? nc + b : nc;
    if (D == 3) t_cells = res->tile_n22;  // 2-2 TETs before every tile (k_cutmask_emit)
  } else {
    GC_CHECK(scratch.alloc(&tiles, sizeof(uint32_t) * ntiles * 3));
    GC_CHECK(scratch.alloc(&totals_dev, sizeof(uint64_t) * 3));
    t_cells = tiles;
    t_points = tiles + ntiles;
    t_facets = tiles + 2 * ntiles;
  }
  if (!g.simplexified) {
    // 128 cells per block = 128*nt0/32 whole tiles (nt0 in {1, 2, 6})
    const int cpb = 128;
    const unsigned nbc = (unsigned)gc_div_up(ncut, cpb);
    if (D == 2)
      GC_LAUNCH(ctx, "k_cut1_count", k_cut1_count<2><<<nbc, 128, 0, ctx->stream>>>(
          g, res->leaves, ctx->d_simplexify_raw, ctx->d_simplexify_pos, res->lay.cutcells, ncut, cpb, t_cells, t_points, t_facets, ntiles));
    else
      GC_LAUNCH(ctx, "k_cut1_count", k_cut1_count<3><<<nbc, 128, 0, ctx->stream>>>(
          g, res->leaves, ctx->d_simplexify_raw, ctx->d_simplexify_pos, res->lay.cutcells, ncut, cpb, t_cells, t_points, t_facets, ntiles));
  }
  if (!g.simplexified) {
    GC_CHECK(gc_launch_check(ctx, "k_cut1_count"));
    GC_CHECK(gc_exclusive_scan_u32_multi(ctx, tiles, tiles, ntiles, 3, totals_dev));
    uint64_t* h = (uint64_t*)ctx->h_pinned;
    GC_CUDA(cudaMemcpyAsync(h, totals_dev, sizeof(uint64_t) * 3, cudaMemcpyDeviceToHost, ctx->stream));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    nsc = h[0];
    nscp = h[1];
    nsf = h[2];
  }
  const uint64_t lim = 2147483647ull;
  if (nsc * (uint64_t)(D + 1) >= lim || nscp > lim || nsf > lim) {
    ctx->err = "Int32 ids of the SubCellData/SubFacetData layout would overflow";
    return GECUT_ERR_OVERFLOW;
  }
  sz.num_subcells = (int64_t)nsc;
  sz.num_subcell_points = (int64_t)nscp;
  sz.num_subfacets = (int64_t)nsf;
  sz.num_subfacet_points = (int64_t)nscp;  // the facets of a single level set share the subcell points
  GcLayout& lay = res->lay;
  lay.sc_pitch = pad_pitch((int64_t)nsc);
  lay.sf_pitch = pad_pitch((int64_t)nsf);
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_c2p, sizeof(int32_t) * nsc * (D + 1)));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_c2bg, sizeof(int32_t) * nsc));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_X, sizeof(double) * nscp * D));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_R, sizeof(double) * nscp * D));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_state, (size_t)lay.sc_pitch));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_c2p, sizeof(int32_t) * nsf * D));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_c2bg, sizeof(int32_t) * nsf));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_N, sizeof(double) * nsf * D));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_state, (size_t)lay.sf_pitch));
  // simplex bg cells: per-entity measures are not written at cut time (gc_ensure_measures computes them on demand)
  const bool lazy_meas = g.simplexified != 0;
  if (!lazy_meas) {
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_meas, sizeof(double) * nsc));
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_meas, sizeof(double) * nsf));
  }
  lay.sf_X = lay.sc_X;  // device-side alias (read-only results); the host copy-out writes both destinations
  lay.sf_R = lay.sc_R;
  res->sf_points_alias = true;
  sz.subfacet_points_alias = 1;
  GC_CHECK(gc_malloc(ctx, (void**)&res->cut_sc_ptr, sizeof(uint32_t) * 2 * (ncut + 1)));
  if (g.simplexified) GC_CHECK(gc_malloc(ctx, (void**)&res->cell_vio, sizeof(double2) * ncut));
  // tile statistics -> four sums (IN / OUT subcell measure, subfacet measure, IN subcells), read back with the result
  const int NBR = (int)std::min<int64_t>(592, gc_div_up(ntiles, 256));
  double4* tile_stats = nullptr;
  GC_CHECK(scratch.alloc(&tile_stats, sizeof(double4) * (ntiles + NBR + 2)));
  double4* st_partial = tile_stats + ntiles;
  double4* st_out = st_partial + NBR;
  unsigned int* st_ticket = reinterpret_cast<unsigned int*>(st_out + 1);
  GC_CUDA(cudaMemsetAsync(st_ticket, 0, sizeof(unsigned int), ctx->stream));
  {
    const unsigned nbf = (unsigned)gc_div_up(ntiles, CUT1_BLOCK / CUT1_TILE);
#define GC_FILL1(DD, SXV, MB)                                                                                          \
  GC_LAUNCH(ctx, "k_cut1_fill",                                                                                        \
            k_cut1_fill<DD, SXV, MB, !SXV><<<nbf, CUT1_BLOCK, (size_t)cut1_warp_bytes<DD, SXV>() * (CUT1_BLOCK / CUT1_TILE), ctx->stream>>>( \
                g, res->leaves, ctx->d_tables, ctx->d_simplexify_raw, ctx->d_simplexify_pos, lay.cutcells, nsimp, t_cells,   \
                t_points, t_facets, lay, res->cut_sc_ptr, tile_stats, res->cell_vio))
#define GC_FILL2(DD, SXV) GC_FILL1(DD, SXV, (SXV ? CUT1_MINB : 5))  // n-cube cells: 5 blocks (6 spill and lose 4 %)
    if (D == 2) {
      if (g.simplexified) GC_FILL2(2, true); else GC_FILL2(2, false);
    } else {
      if (g.simplexified) GC_FILL2(3, true); else GC_FILL2(3, false);
    }
#undef GC_FILL2
#undef GC_FILL1
  }
  int st = gc_launch_check(ctx, "k_cut1_fill");
  if (res->tile_n22) {  // the tile bases have served (stream-ordered cache: not handed out before the fill has run)
    gc_free(ctx, res->tile_n22);
    res->tile_n22 = nullptr;
  }
  if (st == GECUT_OK) {
    GC_LAUNCH(ctx, "k_tile_stats_reduce", k_tile_stats_reduce<<<NBR, 256, 0, ctx->stream>>>(tile_stats, ntiles, st_partial, st_ticket, st_out));
    st = gc_launch_check(ctx, "k_tile_stats_reduce");
    // the copy completes before gecut_cut returns (stream synchronisation at the end of the cut)
    if (st == GECUT_OK && cudaMemcpyAsync(res->l1_stats_pinned(), st_out, sizeof(double4), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
      ctx->err = "read-back of the tile statistics failed";
      st = GECUT_ERR_CUDA;
    }
    res->has_l1_stats = st == GECUT_OK;
  }
  return st;
}

int gc_cut_cells(gecut_ctx* ctx, gecut_result* res) {
  const GcGrid& g = res->grid;
  const int L = res->L, D = g.D;
  const int64_t ncut = res->sizes.num_cutcells;
  const int64_t nsimp = ncut * g.nt0;
  gecut_sizes& sz = res->sizes;
  sz.num_subcells = sz.num_subcell_points = sz.num_subfacets = sz.num_subfacet_points = 0;
  if (nsimp == 0) return GECUT_OK;
  if (L == 1) return gc_cut_cells_l1(ctx, res);
  // ---- L > 1: tile totals (k_cutm_count + k_cut_walk<count> for the recorded simplices) -> scan over tiles -> fill
  GcScratch scratch(ctx);  // every block below is released on every exit path
  GcMulti mu{};
  mu.nch = 2 + 2 * L;
  mu.ntiles = gc_div_up(nsimp, 32);
  uint64_t* totals_dev = nullptr;
  GC_CHECK(scratch.alloc(&mu.tile, sizeof(uint32_t) * mu.ntiles * mu.nch));
  GC_CHECK(scratch.alloc(&mu.sinfo, sizeof(uint32_t) * nsimp));
  GC_CHECK(scratch.alloc(&mu.slow_simp, sizeof(uint32_t) * nsimp));
  GC_CHECK(scratch.alloc(&mu.slow_bucket, (size_t)nsimp));
  GC_CHECK(scratch.alloc(&mu.slow_mask, sizeof(uint32_t) * nsimp));
  GC_CHECK(scratch.alloc(&mu.slow_count, sizeof(uint32_t) * 16));
  GC_CHECK(scratch.alloc(&totals_dev, sizeof(uint64_t) * mu.nch));
  GC_CUDA(cudaMemsetAsync(mu.slow_count, 0, sizeof(uint32_t) * 16, ctx->stream));
  launch_fast_count(ctx, res, nsimp, mu);
  GC_CHECK(gc_launch_check(ctx, "k_cutm_count"));
  int64_t nslow[CM_NBK], nslow_total = 0;
  {
    uint32_t* hs = (uint32_t*)ctx->h_pinned;
    GC_CUDA(cudaMemcpyAsync(hs, mu.slow_count, sizeof(uint32_t) * (CM_NBK + 1), cudaMemcpyDeviceToHost, ctx->stream));
    GC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < CM_NBK; ++b) nslow[b] = (int64_t)hs[b];
    nslow_total = (int64_t)hs[CM_NBK];
  }
  if (getenv("GECUT_WALK_DEBUG")) fprintf(stderr, "[gecut] recorded simplices: two level sets %lld, tree walk %lld %lld %lld %lld %lld, of %lld\n", (long long)nslow[0], (long long)nslow[1], (long long)nslow[2], (long long)nslow[3], (long long)nslow[4], (long long)nslow[5], (long long)nsimp);
  uint32_t* lists = nullptr;
  if (nslow_total > 0) {
    // slot-indexed counters of the recorded simplices, one dense slot list per launch bucket
    mu.slow_n = nslow_total;
    uint32_t* cursor = nullptr;
    GC_CHECK(scratch.alloc(&mu.slow_val, sizeof(uint32_t) * nslow_total * mu.nch));
    // the depth-2 kernel writes only the channels of its two level sets
    GC_CUDA(cudaMemsetAsync(mu.slow_val, 0, sizeof(uint32_t) * nslow_total * mu.nch, ctx->stream));
    GC_CHECK(scratch.alloc(&lists, sizeof(uint32_t) * nslow_total));
    if (nslow_total > nslow[0]) {
      mu.walk_items_n = (nslow_total - nslow[0]) * WALK_ITEMS;
      GC_CHECK(scratch.alloc(&mu.walk_items, sizeof(uint32_t) * (size_t)mu.walk_items_n * mu.nch));
    }
    if (nslow[0] > 0)  // rounded up to whole blocks: every thread of the launches stores / loads its word
      GC_CHECK(scratch.alloc(&mu.two_items, sizeof(unsigned long long) * (size_t)gc_div_up(nslow[0] * WALK_ITEMS, TWO_T) * TWO_T));
    GC_CHECK(scratch.alloc(&cursor, sizeof(uint32_t) * 2 * CM_NBK));
    uint32_t hb[2 * CM_NBK];
    uint32_t acc = 0;
    for (int b = 0; b < CM_NBK; ++b) {
      hb[b] = acc;           // base of bucket b
      hb[CM_NBK + b] = 0u;   // cursor
      acc += (uint32_t)nslow[b];
    }
    uint32_t* hp = (uint32_t*)ctx->h_pinned + 512;  // the pinned page is reused right away: keep clear of the read-back area
    memcpy(hp, hb, sizeof(hb));
    GC_CUDA(cudaMemcpyAsync(cursor, hp, sizeof(hb), cudaMemcpyHostToDevice, ctx->stream));
    GC_LAUNCH(ctx, "k_cutm_bucketize", k_cutm_bucketize<<<(unsigned)gc_div_up(nslow_total, 256), 256, 0, ctx->stream>>>(
        mu.slow_bucket, nslow_total, cursor, cursor + CM_NBK, lists));
    GC_CHECK(dispatch_walk(ctx, res, mu, false, nslow, lists));
    GC_CHECK(gc_launch_check(ctx, "k_cut_walk(count)"));
  }
  GC_CHECK(gc_exclusive_scan_u32_multi(ctx, mu.tile, mu.tile, mu.ntiles, mu.nch, totals_dev));
  uint64_t* h = (uint64_t*)ctx->h_pinned;
  GC_CUDA(cudaMemcpyAsync(h, totals_dev, sizeof(uint64_t) * mu.nch, cudaMemcpyDeviceToHost, ctx->stream));
  GC_CUDA(cudaStreamSynchronize(ctx->stream));
  uint64_t nsc = h[0], nscp = h[1], nsf = 0, nsfp = 0;
  for (int k = 0; k < L; ++k) {
    mu.fac_base[k] = (uint32_t)nsf;
    mu.fpt_base[k] = (uint32_t)nsfp;
    nsf += h[2 + k];
    nsfp += h[2 + L + k];
  }
  const uint64_t lim = 2147483647ull;
  if (nsc * (uint64_t)(D + 1) >= lim || nscp > lim || nsf > lim || nsfp > lim || nsc > lim) {
    ctx->err = "Int32 ids of the SubCellData/SubFacetData layout would overflow";
    return GECUT_ERR_OVERFLOW;
  }
  sz.num_subcells = (int64_t)nsc;
  sz.num_subcell_points = (int64_t)nscp;
  sz.num_subfacets = (int64_t)nsf;
  sz.num_subfacet_points = (int64_t)nsfp;
  GcLayout& lay = res->lay;
  lay.sc_pitch = pad_pitch((int64_t)nsc);
  lay.sf_pitch = pad_pitch((int64_t)nsf);
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_c2p, sizeof(int32_t) * nsc * (D + 1)));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_c2bg, sizeof(int32_t) * nsc));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_X, sizeof(double) * nscp * D));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_R, sizeof(double) * nscp * D));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_state, (size_t)lay.sc_pitch * L));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sc_meas, sizeof(double) * nsc));
  GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_meas, sizeof(double) * nsf));
  if (nsf > 0) {
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_c2p, sizeof(int32_t) * nsf * D));
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_c2bg, sizeof(int32_t) * nsf));
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_N, sizeof(double) * nsf * D));
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_state, (size_t)lay.sf_pitch * L));
  }
  if (nsfp > 0) {
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_X, sizeof(double) * nsfp * D));
    GC_CHECK(gc_malloc(ctx, (void**)&lay.sf_R, sizeof(double) * nsfp * D));
  }
  GC_CHECK(gc_malloc(ctx, (void**)&res->cut_sc_ptr, sizeof(uint32_t) * 2 * (ncut + 1)));
  launch_fast_fill(ctx, res, nsimp, mu);
  GC_CHECK(gc_launch_check(ctx, "k_cutm_fill"));
  if (nslow_total > 0) GC_CHECK(dispatch_walk(ctx, res, mu, true, nslow, lists));
  return gc_launch_check(ctx, "k_cut_walk(fill)");
}
