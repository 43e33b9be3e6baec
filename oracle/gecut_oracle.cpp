// =============================================================================
// ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Single-threaded CPU restatement of the cutting-and-integration hot path of
// gridap/GridapEmbedded.jl v0.9.11 (`cut(LevelSetCutter(), bgmodel, geo)` plus
// the selectors and the cut-cell quadrature).  It follows the reference source
// stage by stage (same loops, same stage order, same 1-based layouts) and cites
// the reference file:line next to every function.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The shipped CUDA path never calls into it.
//
// PARITY STATUS: the reference is Julia and cannot run in this image (no julia
// binary, Gridap.jl / MiniQhull sources are not vendored).  The marching-simplex
// connectivity inside a lookup-table case therefore comes from
// oracle/gen_lookup_tables.py (scipy's qhull 2020.2, same flags) and is
// "parity unpinned"; everything the reference's own tests pin (compute_case
// known answers, disk area/perimeter, quadrilateral areas, divergence theorem)
// is asserted against this oracle in tests/test_oracle_*.py.
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC  (see oracle/build.sh)
// FP contract is OFF because Julia never fuses a*b+c; IN/OUT decisions on nodes
// that sit on a surface depend on it.
// =============================================================================
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

namespace {

constexpr int IN = -1, OUT = 1, CUT = 0, INTERFACE = 0;  // Interfaces.jl:70-73

// LookupTable{D,T} (LookupTables.jl:2-13), 1-based ids as in the reference.
struct OTable {
  int D;
  int num_cases;
  std::vector<std::vector<std::vector<int>>> case_to_subcell_to_points;
  std::vector<std::vector<int>> case_to_subcell_to_inout;
  std::vector<std::vector<std::vector<int>>> case_to_subfacet_to_points;
  std::vector<std::vector<double>> case_to_subfacet_to_orientation;
  std::vector<int> case_to_num_points;  // length(case_to_point_to_coordinates[case])
  std::vector<int> case_to_inoutcut;
  std::vector<std::vector<int>> edge_to_points;  // MiniCell (CutTriangulations.jl:1-12)
};

#include "oracle_tables.inc"

const OTable& simplex_table(int ds) {
  return ds == 1 ? TABLE_SEGMENT : (ds == 2 ? TABLE_TRI : TABLE_TET);
}

inline bool isout(double v) { return v > 0; }  // LookupTables.jl:40-42

// ----------------------------------------------------------------------------
// Leaves (AnalyticalGeometries.jl:75-78,146-198).  No FMA, left-assoc sums.
// ----------------------------------------------------------------------------
enum LeafKind { SPHERE = 0, CYLINDER = 1, PLANE = 2, DOUGHNUT = 3, NODAL = 4 };
struct OLeaf {
  int32_t kind;
  int32_t pad;
  double x0[3];
  double v[3];
  double R;
  double r;
};

inline double dotD(const double* a, const double* b, int D) {
  double s = a[0] * b[0];
  for (int d = 1; d < D; ++d) s = s + a[d] * b[d];
  return s;
}

double eval_leaf(const OLeaf& lf, const double* x, int D) {
  double w[3] = {0, 0, 0};
  for (int d = 0; d < D; ++d) w[d] = x[d] - lf.x0[d];
  switch (lf.kind) {
    case SPHERE: {  // _sphere: w.w - R^2   (AnalyticalGeometries.jl:146-150)
      return dotD(w, w, D) - lf.R * lf.R;
    }
    case CYLINDER: {  // _cylinder (AnalyticalGeometries.jl:176-182)
      double A = dotD(w, lf.v, D);
      double H2 = dotD(w, w, D);
      double B = H2 - A * A;
      return B - lf.R * lf.R;
    }
    case PLANE: {  // _plane (AnalyticalGeometries.jl:194-198)
      return dotD(w, lf.v, D);
    }
    case DOUGHNUT: {  // _doughnut_fun (AnalyticalGeometries.jl:75-78)
      double t = lf.R - std::sqrt(w[0] * w[0] + w[1] * w[1]);
      return (t * t + w[2] * w[2]) - lf.r * lf.r;
    }
    default:
      return std::numeric_limits<double>::quiet_NaN();
  }
}

// ----------------------------------------------------------------------------
// Background grid: CartesianDiscreteModel(pmin,pmax,partition) optionally
// simplexified (Gridap conventions A1,A2,A4 of SURVEY.md section 8c).
// ----------------------------------------------------------------------------
struct OGrid {
  int32_t D;
  int32_t simplexified;
  double pmin[3];
  double pmax[3];
  int32_t partition[3];
  int32_t pad;
};

struct BgGrid {
  int D;
  bool simplexified;
  double pmin[3], dx[3];
  int64_t n[3];   // cells per direction
  int64_t nn[3];  // nodes per direction
  int nv;         // vertices per bg cell
  int nt_hex;     // simplices per n-cube (2 or 6); bg cells per n-cube if simplexified
  int64_t num_cubes, num_cells, num_nodes;
  const std::vector<std::vector<int>>* simplexify_raw;  // of the n-cube
  const std::vector<std::vector<int>>* lt_pos;          // stage-0 ltcell_to_lpoints of the bg polytope
  const std::vector<std::vector<double>>* vertices;     // bg polytope reference vertices

  explicit BgGrid(const OGrid& g) {
    D = g.D;
    simplexified = g.simplexified != 0;
    num_cubes = 1;
    num_nodes = 1;
    for (int d = 0; d < 3; ++d) {
      n[d] = d < D ? g.partition[d] : 1;
      nn[d] = d < D ? n[d] + 1 : 1;
      pmin[d] = d < D ? g.pmin[d] : 0.0;
      // CartesianDescriptor: sizes = (pmax-pmin)/partition
      dx[d] = d < D ? (g.pmax[d] - g.pmin[d]) / (double)g.partition[d] : 0.0;
      num_cubes *= n[d];
      num_nodes *= nn[d];
    }
    simplexify_raw = D == 2 ? &SIMPLEXIFY_RAW_QUAD : &SIMPLEXIFY_RAW_HEX;
    nt_hex = (int)simplexify_raw->size();
    if (simplexified) {
      nv = D + 1;
      num_cells = num_cubes * nt_hex;
      lt_pos = D == 2 ? &SIMPLEXIFY_POS_TRI : &SIMPLEXIFY_POS_TET;
      vertices = D == 2 ? &VERTICES_TRI : &VERTICES_TET;
    } else {
      nv = 1 << D;
      num_cells = num_cubes;
      lt_pos = D == 2 ? &SIMPLEXIFY_POS_QUAD : &SIMPLEXIFY_POS_HEX;
      vertices = D == 2 ? &VERTICES_QUAD : &VERTICES_HEX;
    }
  }
  // CartesianCoordinates getindex: x0[d] + (I[d]-1)*dx[d]
  void node_coords(int64_t node0, double* x) const {
    int64_t i = node0 % nn[0], j = (node0 / nn[0]) % nn[1], k = node0 / (nn[0] * nn[1]);
    int64_t I[3] = {i, j, k};
    for (int d = 0; d < D; ++d) x[d] = pmin[d] + (double)I[d] * dx[d];
  }
  // 0-based node ids of 0-based bg cell (vertex order x-fastest; simplexified:
  // cell = nt*cube + ltcell, nodes = cube_nodes[simplexify(p)[ltcell]]).
  void cell_nodes(int64_t cell0, int64_t* nodes) const {
    int64_t cube = simplexified ? cell0 / nt_hex : cell0;
    int64_t i = cube % n[0], j = (cube / n[0]) % n[1], k = cube / (n[0] * n[1]);
    int64_t cn[8];
    int nvc = 1 << D;
    for (int lv = 0; lv < nvc; ++lv) {
      int64_t ii = i + (lv & 1), jj = j + ((lv >> 1) & 1), kk = k + ((lv >> 2) & 1);
      cn[lv] = ii + nn[0] * (jj + nn[1] * kk);
    }
    if (!simplexified) {
      for (int lv = 0; lv < nvc; ++lv) nodes[lv] = cn[lv];
    } else {
      const std::vector<int>& lt = (*simplexify_raw)[cell0 % nt_hex];
      for (int lv = 0; lv <= D; ++lv) nodes[lv] = cn[lt[lv] - 1];
    }
  }
};

// ----------------------------------------------------------------------------
// Working mesh = `CutTriangulation` wrapping a SubCellData or a SubFacetData
// (CutTriangulations.jl:14-48; SubCellTriangulations.jl:2-7;
// SubFacetTriangulations.jl:2-8).
// ----------------------------------------------------------------------------
struct Mesh {
  int ds = 0;  // dimension of the simplices (Dc for subcells, Dp-1 for subfacets)
  int dp = 0;  // dimension of the points
  int dr = -1; // dimension of the reference coordinates (-1: dp; ds for the sub-mesh of a bg FACET grid)
  int rdim() const { return dr < 0 ? dp : dr; }
  bool is_facets = false;
  std::vector<int32_t> c2p;   // Table data, (ds+1) ids per cell, 1-based
  std::vector<int32_t> c2bg;  // cell_to_bgcell / facet_to_bgcell, 1-based
  std::vector<double> c2n;    // facet_to_normal (dp per facet), facets only
  std::vector<double> X;      // point_to_coords (dp per point)
  std::vector<double> R;      // point_to_rcoords (dp per point)
  std::vector<std::vector<int8_t>> done;     // done_ls_to_cell_to_inoutcut
  std::vector<std::vector<double>> pending;  // pending_ls_to_point_to_value
  int64_t ncells() const { return (int64_t)c2bg.size(); }
  int64_t npoints() const { return (int64_t)X.size() / dp; }
};

// compute_case (LookupTables.jl:19-33)
inline int compute_case(const Mesh& m, const std::vector<double>& val, int64_t mcell0) {
  int c = 1;
  const int32_t* p = &m.c2p[mcell0 * (m.ds + 1)];
  for (int i = 0; i <= m.ds; ++i)
    if (isout(val[p[i] - 1])) c += 1 << i;
  return c;
}

// _normal_vector / _orthogonal_vector (LookupTables.jl:339-360)
void setup_normal(const std::vector<int>& subpoints, const std::vector<double>& X, int dp, int64_t pointoffset,
                  double* n) {
  if (dp == 2) {
    const double* p1 = &X[(subpoints[0] + pointoffset - 1) * 2];
    const double* p2 = &X[(subpoints[1] + pointoffset - 1) * 2];
    double v1 = p2[0] - p1[0], v2 = p2[1] - p1[1];
    double w1 = v2, w2 = -v1;
    double m = std::sqrt(w1 * w1 + w2 * w2);
    if (m < std::numeric_limits<double>::epsilon()) {
      n[0] = 0;
      n[1] = 0;
    } else {
      n[0] = w1 / m;
      n[1] = w2 / m;
    }
  } else {
    const double* p1 = &X[(subpoints[0] + pointoffset - 1) * 3];
    const double* p2 = &X[(subpoints[1] + pointoffset - 1) * 3];
    const double* p3 = &X[(subpoints[2] + pointoffset - 1) * 3];
    double a[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    double b[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
    double w1 = a[1] * b[2] - a[2] * b[1];
    double w2 = a[2] * b[0] - a[0] * b[2];
    double w3 = a[0] * b[1] - a[1] * b[0];
    double m = std::sqrt(w1 * w1 + w2 * w2 + w3 * w3);
    if (m < std::numeric_limits<double>::epsilon()) {
      n[0] = n[1] = n[2] = 0;
    } else {
      n[0] = w1 / m;
      n[1] = w2 / m;
      n[2] = w3 / m;
    }
  }
}

// One stage: count -> allocate -> fill.
// cut_sub_triangulation (CutTriangulations.jl:144-177) and
// cut_sub_triangulation_with_boundary (CutTriangulations.jl:231-266).
// `val` is the level set being cut; `m.pending` are the arrays carried along.
void cut_stage(const Mesh& m, const std::vector<double>& val, bool with_boundary, Mesh& s,
               std::vector<int8_t>& scell_to_inoutcut, Mesh* sb) {
  const OTable& table = simplex_table(m.ds);
  const int nlp = m.ds + 1;
  // count_sub_triangulation(_with_boundary) (CutTriangulations.jl:157-166,238-251)
  int64_t n_scells = 0, n_spoints = 0, n_sfacets = 0;
  for (int64_t mcell = 0; mcell < m.ncells(); ++mcell) {
    int c = compute_case(m, val, mcell) - 1;
    n_scells += (int64_t)table.case_to_subcell_to_inout[c].size();
    n_spoints += table.case_to_num_points[c];
    n_sfacets += (int64_t)table.case_to_subfacet_to_points[c].size();
  }
  // allocate_sub_triangulation (CutTriangulations.jl:50-65,350-363,405-423)
  s = Mesh();
  s.ds = m.ds;
  s.dp = m.dp;
  s.dr = m.dr;
  s.is_facets = m.is_facets;
  s.c2p.assign(n_scells * nlp, 0);
  s.c2bg.assign(n_scells, 0);
  if (m.is_facets) s.c2n.assign(n_scells * m.dp, 0.0);
  s.X.assign(n_spoints * m.dp, 0.0);
  s.R.assign(n_spoints * m.rdim(), 0.0);
  s.done.assign(m.done.size(), std::vector<int8_t>(n_scells, 0));
  s.pending.assign(m.pending.size(), std::vector<double>(n_spoints, 0.0));
  scell_to_inoutcut.assign(n_scells, 0);
  if (with_boundary) {
    // allocate_boundary_triangulation (CutTriangulations.jl:365-378): the facet
    // struct aliases the new stage's point arrays; copied after the fill below.
    *sb = Mesh();
    sb->ds = m.ds - 1;
    sb->dp = m.dp;
    sb->is_facets = true;
    sb->c2p.assign(n_sfacets * m.ds, 0);
    sb->c2bg.assign(n_sfacets, 0);
    sb->c2n.assign(n_sfacets * m.dp, 0.0);
  }
  // cut_sub_triangulation(_with_boundary)! (CutTriangulations.jl:168-177,253-266)
  int64_t scell = 0, spoint = 0, q = 0, sfacet = 0, z = 0;
  const int dp = m.dp, dr = m.rdim();
  auto copy_point = [&](int64_t sp1, int64_t mp1) {  // set_point_data! (CutTriangulations.jl:93-101,113-121)
    for (size_t l = 0; l < s.pending.size(); ++l) s.pending[l][sp1 - 1] = m.pending[l][mp1 - 1];
    for (int d = 0; d < dp; ++d) s.X[(sp1 - 1) * dp + d] = m.X[(mp1 - 1) * dp + d];
    for (int d = 0; d < dr; ++d) s.R[(sp1 - 1) * dr + d] = m.R[(mp1 - 1) * dr + d];
  };
  auto lerp_point = [&](int64_t sp1, int64_t mp1, int64_t mp2, double coeff) {  // CutTriangulations.jl:123-135
    for (size_t l = 0; l < s.pending.size(); ++l) {
      double a = m.pending[l][mp1 - 1], b = m.pending[l][mp2 - 1];
      s.pending[l][sp1 - 1] = a + coeff * (b - a);
    }
    for (int d = 0; d < dp; ++d) {
      double a = m.X[(mp1 - 1) * dp + d], b = m.X[(mp2 - 1) * dp + d];
      s.X[(sp1 - 1) * dp + d] = a + coeff * (b - a);
    }
    for (int d = 0; d < dr; ++d) {
      double ra = m.R[(mp1 - 1) * dr + d], rb = m.R[(mp2 - 1) * dr + d];
      s.R[(sp1 - 1) * dr + d] = ra + coeff * (rb - ra);
    }
  };
  for (int64_t mcell = 0; mcell < m.ncells(); ++mcell) {
    int c = compute_case(m, val, mcell) - 1;
    int64_t pointoffset = spoint;
    // _cut_cell! (CutTriangulations.jl:179-222)
    int nsubcells = (int)table.case_to_subcell_to_inout[c].size();
    for (int sc = 0; sc < nsubcells; ++sc) {
      scell += 1;
      scell_to_inoutcut[scell - 1] = (int8_t)table.case_to_subcell_to_inout[c][sc];
      // set_cell_data! (CutTriangulations.jl:79-87,392-394,433-436)
      for (size_t l = 0; l < s.done.size(); ++l) s.done[l][scell - 1] = m.done[l][mcell];
      s.c2bg[scell - 1] = m.c2bg[mcell];
      if (m.is_facets)
        for (int d = 0; d < dp; ++d) s.c2n[(scell - 1) * dp + d] = m.c2n[mcell * dp + d];
      for (int subpoint : table.case_to_subcell_to_points[c][sc]) {
        q += 1;
        s.c2p[q - 1] = (int32_t)(subpoint + spoint);
      }
    }
    const int32_t* mp = &m.c2p[mcell * nlp];
    for (int lpoint = 0; lpoint < nlp; ++lpoint) {
      spoint += 1;
      copy_point(spoint, mp[lpoint]);
    }
    if (CUT == table.case_to_inoutcut[c]) {
      for (const std::vector<int>& lpoints : table.edge_to_points) {
        int64_t mpoint1 = mp[lpoints[0] - 1], mpoint2 = mp[lpoints[1] - 1];
        double v1 = val[mpoint1 - 1], v2 = val[mpoint2 - 1];
        if (isout(v1) != isout(v2)) {
          spoint += 1;
          double w1 = std::fabs(v1), w2 = std::fabs(v2);
          double coeff = w1 / (w1 + w2);
          lerp_point(spoint, mpoint1, mpoint2, coeff);
        }
      }
    }
    if (with_boundary) {
      // _cut_cell_boundary! (CutTriangulations.jl:268-292)
      int nsubfacets = (int)table.case_to_subfacet_to_points[c].size();
      for (int sf = 0; sf < nsubfacets; ++sf) {
        sfacet += 1;
        for (int subpoint : table.case_to_subfacet_to_points[c][sf]) {
          z += 1;
          sb->c2p[z - 1] = (int32_t)(subpoint + pointoffset);
        }
        sb->c2bg[sfacet - 1] = m.c2bg[mcell];  // set_cell_data_boundary! (CutTriangulations.jl:401-403)
        double nrm[3];
        setup_normal(table.case_to_subfacet_to_points[c][sf], s.X, dp, pointoffset, nrm);
        double orientation = table.case_to_subfacet_to_orientation[c][sf];
        for (int d = 0; d < dp; ++d) sb->c2n[(sfacet - 1) * dp + d] = orientation * nrm[d];
      }
    }
  }
  if (with_boundary) {
    sb->X = s.X;
    sb->R = s.R;
  }
}

// cut_sub_triangulation_several_levelsets (CutTriangulations.jl:294-307)
void cut_several_levelsets(Mesh m, std::vector<std::vector<double>> pending, Mesh& out,
                           std::vector<std::vector<int8_t>>& done_out) {
  m.pending = std::move(pending);
  m.done.clear();
  while (!m.pending.empty()) {
    std::vector<double> point_to_value = std::move(m.pending.back());
    m.pending.pop_back();
    Mesh s;
    std::vector<int8_t> inout;
    cut_stage(m, point_to_value, false, s, inout, nullptr);
    s.done.insert(s.done.begin(), std::move(inout));  // pushfirst!
    m = std::move(s);
  }
  done_out = std::move(m.done);
  m.done.clear();
  out = std::move(m);
}

// ----------------------------------------------------------------------------
// Result = EmbeddedDiscretization (EmbeddedDiscretizations.jl:36-45)
// ----------------------------------------------------------------------------
struct Result {
  BgGrid grid;
  int L;
  std::vector<std::vector<int8_t>> ls_to_bgcell_to_inoutcut;
  std::vector<int32_t> cutcell_to_cell;  // 1-based, ascending
  Mesh subcells;
  std::vector<std::vector<int8_t>> ls_to_subcell_to_inout;
  Mesh subfacets;
  std::vector<std::vector<int8_t>> ls_to_subfacet_to_inout;
  // scratch returned to the caller (kept alive by the handle)
  std::map<std::string, std::vector<int8_t>> i8;
  std::map<std::string, std::vector<int32_t>> i32;
  std::map<std::string, std::vector<double>> f64;
  explicit Result(const OGrid& g) : grid(g), L(0) {}
};

inline int case_to_inoutcut_ncube(int c1, int nv) {  // LookupTables.jl:289-302 for any polytope
  if (c1 == 1) return IN;
  if (c1 == (1 << nv)) return OUT;
  return CUT;
}

Result* do_cut(const OGrid& og, const OLeaf* leaves, int L, const double* const* nodal_values) {
  Result* r = new Result(og);
  const BgGrid& g = r->grid;
  r->L = L;
  const int D = g.D, nv = g.nv;
  // ---- discretize (DiscreteGeometries.jl:48-63): every unique leaf at every node
  std::vector<std::vector<double>> ls_to_point_to_value(L);
  for (int ls = 0; ls < L; ++ls) {
    std::vector<double>& v = ls_to_point_to_value[ls];
    v.resize(g.num_nodes);
    if (leaves[ls].kind == NODAL) {
      std::memcpy(v.data(), nodal_values[ls], sizeof(double) * g.num_nodes);
    } else {
      for (int64_t node = 0; node < g.num_nodes; ++node) {
        double x[3];
        g.node_coords(node, x);
        v[node] = eval_leaf(leaves[ls], x, D);
      }
    }
  }
  // ---- _extract_grid_of_cut_cells (CutTriangulations.jl:462-479)
  // _compute_in_out_or_cut per ls (CutTriangulations.jl:619-640)
  r->ls_to_bgcell_to_inoutcut.assign(L, std::vector<int8_t>());
  for (int ls = 0; ls < L; ++ls) {
    std::vector<int8_t>& st = r->ls_to_bgcell_to_inoutcut[ls];
    st.resize(g.num_cells);
    const std::vector<double>& v = ls_to_point_to_value[ls];
    for (int64_t cell = 0; cell < g.num_cells; ++cell) {
      int64_t nodes[8];
      g.cell_nodes(cell, nodes);
      int c = 1;
      for (int i = 0; i < nv; ++i)
        if (isout(v[nodes[i]])) c += 1 << i;
      st[cell] = (int8_t)case_to_inoutcut_ncube(c, nv);
    }
  }
  // _find_cut_cells (CutTriangulations.jl:481-493)
  for (int64_t cell = 0; cell < g.num_cells; ++cell) {
    bool iscut = false;
    for (int ls = 0; ls < L; ++ls)
      if (r->ls_to_bgcell_to_inoutcut[ls][cell] == CUT) iscut = true;
    if (iscut) r->cutcell_to_cell.push_back((int32_t)(cell + 1));
  }
  // ---- _simplexify_and_isolate_cells_in_cutgrid + _simplexify (CutTriangulations.jl:505-617)
  const std::vector<std::vector<int>>& lt = *g.lt_pos;  // already positive in reference coords (:512-514)
  const int nlt = (int)lt.size();
  const int nsp = D + 1;
  const int64_t ncut = (int64_t)r->cutcell_to_cell.size();
  Mesh m0;
  m0.ds = D;
  m0.dp = D;
  m0.c2p.assign(ncut * nlt * nsp, 0);
  m0.c2bg.assign(ncut * nlt, 0);
  m0.X.assign(ncut * nv * D, 0.0);
  m0.R.assign(ncut * nv * D, 0.0);
  m0.pending.assign(L, std::vector<double>(ncut * nv, 0.0));
  {
    int64_t tpoint = 0, tcell = 0;
    for (int64_t k = 0; k < ncut; ++k) {
      int64_t cell = r->cutcell_to_cell[k] - 1;
      for (int ltc = 0; ltc < nlt; ++ltc) {
        tcell += 1;
        for (int j = 0; j < nsp; ++j) m0.c2p[(tcell - 1) * nsp + j] = (int32_t)(tpoint + lt[ltc][j]);
        m0.c2bg[tcell - 1] = (int32_t)(cell + 1);  // _setup_cell_to_bgcell (:550-560)
      }
      int64_t nodes[8];
      g.cell_nodes(cell, nodes);
      for (int lpoint = 0; lpoint < nv; ++lpoint) {
        tpoint += 1;
        double x[3];
        g.node_coords(nodes[lpoint], x);
        for (int d = 0; d < D; ++d) {
          m0.X[(tpoint - 1) * D + d] = x[d];
          m0.R[(tpoint - 1) * D + d] = (*g.vertices)[lpoint][d];
        }
        for (int ls = 0; ls < L; ++ls) m0.pending[ls][tpoint - 1] = ls_to_point_to_value[ls][nodes[lpoint]];
      }
    }
  }
  // _ensure_positive_jacobians!(tcell_to_tpoints,tpoint_to_coords,simplex) in PHYSICAL coords
  // (CutTriangulations.jl:532-533 -> LookupTables.jl:199-216): J = sum_i outer(gradN_i, x_i)
  for (int64_t tcell = 0; tcell < m0.ncells(); ++tcell) {
    int32_t* p = &m0.c2p[tcell * nsp];
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < nsp; ++i) {
      const double* x = &m0.X[(int64_t)(p[i] - 1) * D];
      for (int a = 0; a < D; ++a) {
        double ga = (i == 0) ? -1.0 : ((i - 1 == a) ? 1.0 : 0.0);
        for (int b = 0; b < D; ++b) J[a][b] += ga * x[b];
      }
    }
    double dV;
    if (D == 2)
      dV = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    else
      dV = J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
           (J[0][0] * J[1][2] * J[2][1] + J[0][1] * J[1][0] * J[2][2] + J[0][2] * J[1][1] * J[2][0]);
    if (dV < 0) std::swap(p[0], p[1]);
  }
  // ---- cut_sub_triangulation_with_boundary_several_levelsets (CutTriangulations.jl:309-343)
  Mesh m = std::move(m0);
  std::vector<Mesh> done_ls_to_boundary(L);
  std::vector<std::vector<std::vector<int8_t>>> done_ls_to_ls_to_facet_inoutcut(L);
  for (int ls = L - 1; ls >= 0; --ls) {
    Mesh s, mb;
    std::vector<int8_t> cell_to_inoutcut;
    cut_stage(m, m.pending[ls], true, s, cell_to_inoutcut, &mb);
    // _remove_index (:345-348)
    std::vector<std::vector<double>> others;
    for (int k = 0; k < L; ++k)
      if (k != ls) others.push_back(s.pending[k]);
    Mesh fb;
    std::vector<std::vector<int8_t>> ls_to_facet_to_inoutcut;
    cut_several_levelsets(std::move(mb), std::move(others), fb, ls_to_facet_to_inoutcut);
    int64_t n_facets = fb.ncells();
    ls_to_facet_to_inoutcut.insert(ls_to_facet_to_inoutcut.begin() + ls, std::vector<int8_t>(n_facets, (int8_t)INTERFACE));
    s.done.insert(s.done.begin(), std::move(cell_to_inoutcut));  // pushfirst!
    done_ls_to_boundary[ls] = std::move(fb);
    done_ls_to_ls_to_facet_inoutcut[ls] = std::move(ls_to_facet_to_inoutcut);
    m = std::move(s);
  }
  // merge_sub_face_data (SubFacetTriangulations.jl:161-208)
  Mesh fst;
  fst.ds = D - 1;
  fst.dp = D;
  fst.is_facets = true;
  r->ls_to_subfacet_to_inout.assign(L, std::vector<int8_t>());
  for (int i = 0; i < L; ++i) {
    const Mesh& b = done_ls_to_boundary[i];
    int64_t o = fst.npoints();
    fst.c2n.insert(fst.c2n.end(), b.c2n.begin(), b.c2n.end());
    fst.c2bg.insert(fst.c2bg.end(), b.c2bg.begin(), b.c2bg.end());
    fst.X.insert(fst.X.end(), b.X.begin(), b.X.end());
    fst.R.insert(fst.R.end(), b.R.begin(), b.R.end());
    for (int32_t id : b.c2p) fst.c2p.push_back((int32_t)(id + o));
    for (int j = 0; j < L; ++j) {
      const std::vector<int8_t>& f = done_ls_to_ls_to_facet_inoutcut[i][j];
      r->ls_to_subfacet_to_inout[j].insert(r->ls_to_subfacet_to_inout[j].end(), f.begin(), f.end());
    }
  }
  r->ls_to_subcell_to_inout = std::move(m.done);
  m.done.clear();
  m.pending.clear();
  r->subcells = std::move(m);
  r->subfacets = std::move(fst);
  return r;
}

// ----------------------------------------------------------------------------
// cut_facets(LevelSetCutter(), bgmodel, geo) (LevelSetCutters.jl:107-142):
// the same machinery on `Grid(ReferenceFE{D-1}, model)`, without boundary.
// Result = EmbeddedFacetDiscretization (EmbeddedFacetDiscretizations.jl:28-35).
//
// Assumption A8 (Gridap 0.20.x, not vendored, unpinned by the reference tests):
//  * facets of the model are numbered by first encounter over (cell ascending,
//    local facet ascending) (`generate_cell_to_faces`), and a facet takes the
//    node order of the local facet of the cell that meets it first;
//  * local facets of the n-cubes: QUAD [[1,2],[3,4],[1,3],[2,4]],
//    HEX [[1,2,3,4],[5,6,7,8],[1,2,5,6],[3,4,7,8],[1,3,5,7],[2,4,6,8]];
//  * simplexified models (cells = the simplices of simplexify(model), cube-major): the same rule with the local
//    facets of the simplices, TRI [[1,2],[1,3],[2,3]], TET [[1,2,3],[1,2,4],[1,3,4],[2,3,4]]; the bg facets are
//    SEGMENTs / TRIs (simplexify(TRI) = [[1,2,3]], reference vertices (0,0),(1,0),(0,1)).
// ----------------------------------------------------------------------------
static const std::vector<std::vector<int>> LFACETS_QUAD = {{1, 2}, {3, 4}, {1, 3}, {2, 4}};
static const std::vector<std::vector<int>> LFACETS_HEX = {{1, 2, 3, 4}, {5, 6, 7, 8}, {1, 2, 5, 6},
                                                          {3, 4, 7, 8}, {1, 3, 5, 7}, {2, 4, 6, 8}};
static const std::vector<std::vector<int>> LFACETS_TRI = {{1, 2}, {1, 3}, {2, 3}};
static const std::vector<std::vector<int>> LFACETS_TET = {{1, 2, 3}, {1, 2, 4}, {1, 3, 4}, {2, 3, 4}};
static const std::vector<std::vector<int>> SIMPLEXIFY_POS_SEGMENT = {{1, 2}};
static const std::vector<std::vector<int>> SIMPLEXIFY_TRI_FACET = {{1, 2, 3}};
static const std::vector<std::vector<double>> VERTICES_SEGMENT = {{0.0}, {1.0}};

struct FacetResult {
  BgGrid grid;
  int L;
  int nvf;                                    // nodes per bg facet
  std::vector<int64_t> facet_nodes;           // nvf 0-based node ids per bg facet
  std::vector<int8_t> facet_is_boundary;      // 1: the facet belongs to one cell only
  std::vector<std::vector<int8_t>> ls_to_facet_to_inoutcut;
  std::vector<int32_t> cutfacet_to_facet;     // 1-based, ascending
  Mesh subfacets;                             // SubCellData{D-1,D}
  std::vector<std::vector<int8_t>> ls_to_subfacet_to_inout;
  explicit FacetResult(const OGrid& g) : grid(g), L(0), nvf(0) {}
};

FacetResult* do_cut_facets(const OGrid& og, const OLeaf* leaves, int L, const double* const* nodal_values) {
  FacetResult* r = new FacetResult(og);
  const BgGrid& g = r->grid;
  r->L = L;
  const bool sx = g.simplexified;
  const int D = g.D, nvf = sx ? D : 1 << (D - 1);
  r->nvf = nvf;
  // ---- Grid(ReferenceFE{D-1}, model): first-encounter numbering (A8)
  const std::vector<std::vector<int>>& lfacets =
      sx ? (D == 2 ? LFACETS_TRI : LFACETS_TET) : (D == 2 ? LFACETS_QUAD : LFACETS_HEX);
  {
    std::map<std::array<int64_t, 4>, int32_t> seen;
    for (int64_t cell = 0; cell < g.num_cells; ++cell) {
      int64_t nodes[8];
      g.cell_nodes(cell, nodes);
      for (const std::vector<int>& lf : lfacets) {
        std::array<int64_t, 4> key = {-1, -1, -1, -1};
        for (int i = 0; i < nvf; ++i) key[i] = nodes[lf[i] - 1];
        std::array<int64_t, 4> sorted = key;
        std::sort(sorted.begin(), sorted.end());
        auto it = seen.find(sorted);
        if (it == seen.end()) {
          seen.emplace(sorted, (int32_t)r->facet_is_boundary.size());
          for (int i = 0; i < nvf; ++i) r->facet_nodes.push_back(key[i]);
          r->facet_is_boundary.push_back(1);
        } else {
          r->facet_is_boundary[it->second] = 0;
        }
      }
    }
  }
  const int64_t nfacets = (int64_t)r->facet_is_boundary.size();
  // ---- discretize (DiscreteGeometries.jl:48-63)
  std::vector<std::vector<double>> ls_to_point_to_value(L);
  for (int ls = 0; ls < L; ++ls) {
    std::vector<double>& v = ls_to_point_to_value[ls];
    v.resize(g.num_nodes);
    if (leaves[ls].kind == NODAL) {
      std::memcpy(v.data(), nodal_values[ls], sizeof(double) * g.num_nodes);
    } else {
      for (int64_t node = 0; node < g.num_nodes; ++node) {
        double x[3];
        g.node_coords(node, x);
        v[node] = eval_leaf(leaves[ls], x, D);
      }
    }
  }
  // ---- _extract_grid_of_cut_cells on the facet grid (CutTriangulations.jl:462-493)
  r->ls_to_facet_to_inoutcut.assign(L, std::vector<int8_t>(nfacets, 0));
  for (int ls = 0; ls < L; ++ls)
    for (int64_t f = 0; f < nfacets; ++f) {
      int c = 1;
      for (int i = 0; i < nvf; ++i)
        if (isout(ls_to_point_to_value[ls][r->facet_nodes[f * nvf + i]])) c += 1 << i;
      r->ls_to_facet_to_inoutcut[ls][f] = (int8_t)case_to_inoutcut_ncube(c, nvf);
    }
  for (int64_t f = 0; f < nfacets; ++f) {
    bool iscut = false;
    for (int ls = 0; ls < L; ++ls)
      if (r->ls_to_facet_to_inoutcut[ls][f] == CUT) iscut = true;
    if (iscut) r->cutfacet_to_facet.push_back((int32_t)(f + 1));
  }
  // ---- _simplexify_and_isolate_cells_in_cutgrid (CutTriangulations.jl:505-548), Dc = D-1, Dp = D
  const std::vector<std::vector<int>>& lt = D == 2 ? SIMPLEXIFY_POS_SEGMENT : (sx ? SIMPLEXIFY_TRI_FACET : SIMPLEXIFY_POS_QUAD);
  const std::vector<std::vector<double>>& verts = D == 2 ? VERTICES_SEGMENT : (sx ? VERTICES_TRI : VERTICES_QUAD);
  const int nlt = (int)lt.size(), nsp = D, ds = D - 1;
  const int64_t ncut = (int64_t)r->cutfacet_to_facet.size();
  Mesh m0;
  m0.ds = ds;
  m0.dp = D;
  m0.dr = ds;
  m0.c2p.assign(ncut * nlt * nsp, 0);
  m0.c2bg.assign(ncut * nlt, 0);
  m0.X.assign(ncut * nvf * D, 0.0);
  m0.R.assign(ncut * nvf * ds, 0.0);
  m0.pending.assign(L, std::vector<double>(ncut * nvf, 0.0));
  int64_t tpoint = 0, tcell = 0;
  for (int64_t k = 0; k < ncut; ++k) {
    const int64_t f = r->cutfacet_to_facet[k] - 1;
    for (int ltc = 0; ltc < nlt; ++ltc) {
      tcell += 1;
      for (int j = 0; j < nsp; ++j) m0.c2p[(tcell - 1) * nsp + j] = (int32_t)(tpoint + lt[ltc][j]);
      m0.c2bg[tcell - 1] = (int32_t)(f + 1);
    }
    for (int lpoint = 0; lpoint < nvf; ++lpoint) {
      tpoint += 1;
      const int64_t node = r->facet_nodes[f * nvf + lpoint];
      double x[3];
      g.node_coords(node, x);
      for (int d = 0; d < D; ++d) m0.X[(tpoint - 1) * D + d] = x[d];
      for (int d = 0; d < ds; ++d) m0.R[(tpoint - 1) * ds + d] = verts[lpoint][d];
      for (int ls = 0; ls < L; ++ls) m0.pending[ls][tpoint - 1] = ls_to_point_to_value[ls][node];
    }
  }
  // _ensure_positive_jacobians_facets! (CutTriangulations.jl:642-683): swap the first two ids when the Jacobian of the
  // sub-simplex and the Jacobian of the bg facet map (both at the first reference node) have a negative inner product
  for (int64_t t = 0; t < m0.ncells(); ++t) {
    int32_t* p = &m0.c2p[t * nsp];
    const int64_t k = t / nlt;  // tcell_to_cutcell
    const double* q0 = &m0.X[(k * nvf + 0) * D];
    double dot = 0.0;
    for (int a = 0; a < ds; ++a) {
      const double* xa = &m0.X[(int64_t)(p[a + 1] - 1) * D];
      const double* x0 = &m0.X[(int64_t)(p[0] - 1) * D];
      const double* qa = &m0.X[(k * nvf + (1 << a)) * D];  // d x / d xi_a of the n-cube map at the origin
      for (int d = 0; d < D; ++d) dot += (xa[d] - x0[d]) * (qa[d] - q0[d]);
    }
    if (dot < 0) std::swap(p[0], p[1]);
  }
  // ---- cut_sub_triangulation_several_levelsets (CutTriangulations.jl:294-307)
  std::vector<std::vector<double>> pending = std::move(m0.pending);
  m0.pending.clear();
  Mesh out;
  cut_several_levelsets(std::move(m0), std::move(pending), out, r->ls_to_subfacet_to_inout);
  r->subfacets = std::move(out);
  return r;
}

// ----------------------------------------------------------------------------
// Selectors (EmbeddedDiscretizations.jl:91-245, 297-339, 408-463).
// A geometry (sub)tree is passed as a postfix program: token >= 0 pushes the
// state row of level set `token` (0-based), negative tokens are operators.
// ----------------------------------------------------------------------------
enum Op { OP_UNION = -1, OP_INTERSECT = -2, OP_SETDIFF = -3, OP_COMPLEMENT = -4 };

inline int8_t io_union(int a, int b) {  // _compute_inoutcut_union (:136-145)
  if (a == OUT && b == OUT) return OUT;
  if ((a == CUT && b == CUT) || (a == CUT && b == OUT) || (a == OUT && b == CUT)) return CUT;
  return IN;
}
inline int8_t io_intersection(int a, int b) {  // (:147-156)
  if (a == IN && b == IN) return IN;
  if ((a == CUT && b == CUT) || (a == CUT && b == IN) || (a == IN && b == CUT)) return CUT;
  return OUT;
}
inline int8_t io_setdiff(int a, int b) {  // (:158-167)
  if (a == IN && b == OUT) return IN;
  if ((a == CUT && b == CUT) || (a == CUT && b == OUT) || (a == IN && b == CUT)) return CUT;
  return OUT;
}
inline int8_t io_complementary(int a) {  // (:169-177)
  if (a == OUT) return IN;
  if (a == IN) return OUT;
  return CUT;
}

// compute_inoutcut(tree) (:107-134) on element e of the given rows
int8_t eval_inoutcut(const int32_t* prog, int len, const std::vector<std::vector<int8_t>>& rows, int64_t e) {
  int8_t st[64];
  int sp = 0;
  for (int t = 0; t < len; ++t) {
    int tok = prog[t];
    if (tok >= 0) {
      st[sp++] = rows[tok][e];
    } else if (tok == OP_COMPLEMENT) {
      st[sp - 1] = io_complementary(st[sp - 1]);
    } else {
      int8_t b = st[--sp], a = st[--sp];
      st[sp++] = tok == OP_UNION ? io_union(a, b) : (tok == OP_INTERSECT ? io_intersection(a, b) : io_setdiff(a, b));
    }
  }
  return st[0];
}

// compute_inoutboundary(tree) (:179-245): state + orientation
void eval_inoutboundary(const int32_t* prog, int len, const std::vector<std::vector<int8_t>>& rows, int64_t e,
                        int8_t* state, int8_t* orientation) {
  int8_t st[64], orr[64];
  int sp = 0;
  for (int t = 0; t < len; ++t) {
    int tok = prog[t];
    if (tok >= 0) {
      st[sp] = rows[tok][e];
      orr[sp] = 1;
      sp++;
    } else if (tok == OP_COMPLEMENT) {
      // _compute_orientation_complementary (:239-245)
      if (st[sp - 1] == INTERFACE) orr[sp - 1] = (int8_t)-orr[sp - 1];
      st[sp - 1] = io_complementary(st[sp - 1]);
    } else {
      int8_t b = st[sp - 1], a = st[sp - 2];
      int8_t d2 = orr[sp - 1], d1 = orr[sp - 2];
      sp -= 2;
      int8_t rs, ro;
      if (tok == OP_SETDIFF) {  // _compute_orientation_setdiff (:229-237)
        rs = io_setdiff(a, b);
        ro = (a == INTERFACE) ? d1 : ((b == INTERFACE) ? (int8_t)-d2 : d1);
      } else {  // _compute_orientation_union / intersect (:217-227)
        rs = tok == OP_UNION ? io_union(a, b) : io_intersection(a, b);
        ro = (a == INTERFACE) ? d1 : ((b == INTERFACE) ? d2 : d1);
      }
      st[sp] = rs;
      orr[sp] = ro;
      sp++;
    }
  }
  *state = st[0];
  *orientation = orr[0];
}

// ----------------------------------------------------------------------------
// Quadrature (Gridap `Measure(trian,2)` semantics; assumption A5)
// ----------------------------------------------------------------------------
struct Rule {
  int n;
  double xi[8][3];
  double w[8];
};
Rule simplex_rule_deg2(int ds) {
  Rule q{};
  if (ds == 1) {  // SEGMENT is an n-cube in Gridap: 2-point Gauss-Legendre on [0,1]
    double a = 0.5 - 0.5 / std::sqrt(3.0), b = 0.5 + 0.5 / std::sqrt(3.0);
    q.n = 2;
    q.xi[0][0] = a;
    q.xi[1][0] = b;
    q.w[0] = q.w[1] = 0.5;
  } else if (ds == 2) {  // Strang-Fix 3-point degree 2
    q.n = 3;
    double p[3][2] = {{1.0 / 6, 1.0 / 6}, {2.0 / 3, 1.0 / 6}, {1.0 / 6, 2.0 / 3}};
    for (int i = 0; i < 3; ++i) {
      q.xi[i][0] = p[i][0];
      q.xi[i][1] = p[i][1];
      q.w[i] = 1.0 / 6;
    }
  } else {  // 4-point symmetric degree 2
    q.n = 4;
    double a = (5.0 - std::sqrt(5.0)) / 20.0, b = (5.0 + 3.0 * std::sqrt(5.0)) / 20.0;
    double p[4][3] = {{a, a, a}, {b, a, a}, {a, b, a}, {a, a, b}};
    for (int i = 0; i < 4; ++i) {
      for (int d = 0; d < 3; ++d) q.xi[i][d] = p[i][d];
      q.w[i] = 1.0 / 24;
    }
  }
  return q;
}
Rule ncube_rule_deg2(int D) {  // tensor Gauss-Legendre, 2 points per direction, x fastest
  Rule q{};
  double g[2] = {0.5 - 0.5 / std::sqrt(3.0), 0.5 + 0.5 / std::sqrt(3.0)};
  q.n = 1 << D;
  for (int i = 0; i < q.n; ++i) {
    double w = 1.0;
    for (int d = 0; d < D; ++d) {
      q.xi[i][d] = g[(i >> d) & 1];
      w *= 0.5;
    }
    q.w[i] = w;
  }
  return q;
}

inline double det2(const double J[3][3]) { return J[0][0] * J[1][1] - J[0][1] * J[1][0]; }
inline double det3(const double J[3][3]) {
  return J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
         (J[0][0] * J[1][2] * J[2][1] + J[0][1] * J[1][0] * J[2][2] + J[0][2] * J[1][1] * J[2][0]);
}

// |det J| of a D-simplex in D dims, J rows = x_{i+1}-x_1
double simplex_meas(const Mesh& m, int64_t cell0) {
  const int D = m.dp;
  const int32_t* p = &m.c2p[cell0 * (m.ds + 1)];
  const double* x0 = &m.X[(int64_t)(p[0] - 1) * D];
  if (m.ds == D) {
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < D; ++i) {
      const double* xi = &m.X[(int64_t)(p[i + 1] - 1) * D];
      for (int d = 0; d < D; ++d) J[i][d] = xi[d] - x0[d];
    }
    return std::fabs(D == 2 ? det2(J) : det3(J));
  }
  // facets: sqrt(det(J^T J)) (Gridap `meas` for 1x2 and 2x3 Jacobians)
  if (D == 2) {
    const double* x1 = &m.X[(int64_t)(p[1] - 1) * 2];
    double a = x1[0] - x0[0], b = x1[1] - x0[1];
    return std::sqrt(a * a + b * b);
  }
  const double* x1 = &m.X[(int64_t)(p[1] - 1) * 3];
  const double* x2 = &m.X[(int64_t)(p[2] - 1) * 3];
  double a[3] = {x1[0] - x0[0], x1[1] - x0[1], x1[2] - x0[2]};
  double b[3] = {x2[0] - x0[0], x2[1] - x0[1], x2[2] - x0[2]};
  double n1 = a[1] * b[2] - a[2] * b[1], n2 = a[2] * b[0] - a[0] * b[2], n3 = a[0] * b[1] - a[1] * b[0];
  return std::sqrt(n1 * n1 + n2 * n2 + n3 * n3);
}

// integral of 1 over a sub-simplex with the degree-2 rule: sum_q w_q * meas(J)
double integrate_one(const Mesh& m, int64_t cell0) {
  Rule q = simplex_rule_deg2(m.ds);
  double dV = simplex_meas(m, cell0);
  double s = 0;
  for (int i = 0; i < q.n; ++i) s += q.w[i] * dV;
  return s;
}

// Shape functions of the bg cell (Q1 on n-cubes, P1 on simplices) at reference point eta:
// gradients in REFERENCE coordinates, vertex order A1.
void bg_ref_grads(int D, bool simplex, const double* eta, double g[8][3]) {
  if (simplex) {
    for (int d = 0; d < D; ++d) g[0][d] = -1.0;
    for (int i = 0; i < D; ++i)
      for (int d = 0; d < D; ++d) g[i + 1][d] = (i == d) ? 1.0 : 0.0;
    return;
  }
  int nv = 1 << D;
  for (int a = 0; a < nv; ++a) {
    for (int d = 0; d < D; ++d) {
      double v = 1.0;
      for (int e = 0; e < D; ++e) {
        int bit = (a >> e) & 1;
        if (e == d)
          v *= bit ? 1.0 : -1.0;
        else
          v *= bit ? eta[e] : (1.0 - eta[e]);
      }
      g[a][d] = v;
    }
  }
}

// Inverse-transpose Jacobian of the bg cell map at eta (constant for Cartesian n-cubes and for simplices)
void bg_inv_jac_t(const BgGrid& g, int64_t cell0, double invJt[3][3]) {
  const int D = g.D;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) invJt[a][b] = 0.0;
  if (!g.simplexified) {
    for (int d = 0; d < D; ++d) invJt[d][d] = 1.0 / g.dx[d];
    return;
  }
  int64_t nodes[8];
  g.cell_nodes(cell0, nodes);
  double x[4][3];
  for (int i = 0; i <= D; ++i) g.node_coords(nodes[i], x[i]);
  // Jt[i][d] = d x_d / d xi_i = x_{i+1}[d]-x_1[d]
  double Jt[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < D; ++i)
    for (int d = 0; d < D; ++d) Jt[i][d] = x[i + 1][d] - x[0][d];
  // grad_x phi = inv(Jt) * grad_xi phi ; store inv(Jt) in invJt
  if (D == 2) {
    double det = det2(Jt);
    invJt[0][0] = Jt[1][1] / det;
    invJt[0][1] = -Jt[0][1] / det;
    invJt[1][0] = -Jt[1][0] / det;
    invJt[1][1] = Jt[0][0] / det;
  } else {
    double det = det3(Jt);
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

// K_ab += sum_q w_q |J_s| grad phi_a(eta_q) . grad phi_b(eta_q) for one subcell
void laplacian_subcell(const Result& r, int64_t sc0, double* K /* nv*nv, += */) {
  const BgGrid& g = r.grid;
  const Mesh& m = r.subcells;
  const int D = g.D, nv = g.nv;
  Rule q = simplex_rule_deg2(D);
  double dV = simplex_meas(m, sc0);
  const int32_t* p = &m.c2p[sc0 * (D + 1)];
  double invJt[3][3];
  bg_inv_jac_t(g, m.c2bg[sc0] - 1, invJt);
  for (int iq = 0; iq < q.n; ++iq) {
    // P1 shape functions of the sub-simplex at xi_q: N_1 = 1 - sum xi, N_{i+1} = xi_i
    double N[4];
    double s = 0;
    for (int d = 0; d < D; ++d) s += q.xi[iq][d];
    N[0] = 1.0 - s;
    for (int d = 0; d < D; ++d) N[d + 1] = q.xi[iq][d];
    double eta[3] = {0, 0, 0};
    for (int i = 0; i <= D; ++i)
      for (int d = 0; d < D; ++d) eta[d] += N[i] * m.R[(int64_t)(p[i] - 1) * D + d];
    double gr[8][3], gx[8][3];
    bg_ref_grads(D, g.simplexified, eta, gr);
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

// Uncut bg cell: tensor Gauss (n-cube) or simplex rule (simplexified bg)
double bgcell_volume(const BgGrid& g, int64_t cell0) {
  const int D = g.D;
  if (!g.simplexified) {
    Rule q = ncube_rule_deg2(D);
    double detJ = g.dx[0] * g.dx[1] * (D == 3 ? g.dx[2] : 1.0);
    double s = 0;
    for (int i = 0; i < q.n; ++i) s += q.w[i] * std::fabs(detJ);
    return s;
  }
  int64_t nodes[8];
  g.cell_nodes(cell0, nodes);
  double x[4][3];
  for (int i = 0; i <= D; ++i) g.node_coords(nodes[i], x[i]);
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < D; ++i)
    for (int d = 0; d < D; ++d) J[i][d] = x[i + 1][d] - x[0][d];
  double dV = std::fabs(D == 2 ? det2(J) : det3(J));
  Rule q = simplex_rule_deg2(D);
  double s = 0;
  for (int i = 0; i < q.n; ++i) s += q.w[i] * dV;
  return s;
}

void bgcell_laplacian(const BgGrid& g, int64_t cell0, double* K /* nv*nv, = */) {
  const int D = g.D, nv = g.nv;
  for (int i = 0; i < nv * nv; ++i) K[i] = 0.0;
  double invJt[3][3];
  bg_inv_jac_t(g, cell0, invJt);
  Rule q = g.simplexified ? simplex_rule_deg2(D) : ncube_rule_deg2(D);
  double dV;
  if (!g.simplexified) {
    dV = std::fabs(g.dx[0] * g.dx[1] * (D == 3 ? g.dx[2] : 1.0));
  } else {
    int64_t nodes[8];
    g.cell_nodes(cell0, nodes);
    double x[4][3];
    for (int i = 0; i <= D; ++i) g.node_coords(nodes[i], x[i]);
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < D; ++i)
      for (int d = 0; d < D; ++d) J[i][d] = x[i + 1][d] - x[0][d];
    dV = std::fabs(D == 2 ? det2(J) : det3(J));
  }
  for (int iq = 0; iq < q.n; ++iq) {
    double gr[8][3], gx[8][3];
    bg_ref_grads(D, g.simplexified, q.xi[iq], gr);
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

}  // namespace

// =============================================================================
// C interface for the ctypes test harness (tests/oracle_binding.py)
// =============================================================================
extern "C" {

void* gecut_oracle_cut(const OGrid* grid, const OLeaf* leaves, int L, const double* const* nodal_values) {
  return do_cut(*grid, leaves, L, nodal_values);
}

void gecut_oracle_free(void* h) { delete (Result*)h; }

// sizes: [Nc, Nn, Ncut, Nsc, Nscp, Nsf, Nsfp, L, D, nv]
void gecut_oracle_sizes(void* h, int64_t* out) {
  Result* r = (Result*)h;
  out[0] = r->grid.num_cells;
  out[1] = r->grid.num_nodes;
  out[2] = (int64_t)r->cutcell_to_cell.size();
  out[3] = r->subcells.ncells();
  out[4] = r->subcells.npoints();
  out[5] = r->subfacets.ncells();
  out[6] = r->subfacets.npoints();
  out[7] = r->L;
  out[8] = r->grid.D;
  out[9] = r->grid.nv;
}

// Returns a pointer to an internal array (valid until gecut_oracle_free). `ls` selects the row for L x N arrays.
const void* gecut_oracle_array(void* h, const char* name, int ls, int64_t* n) {
  Result* r = (Result*)h;
  std::string s(name);
#define RET(vec)                 \
  {                              \
    *n = (int64_t)(vec).size();  \
    return (const void*)(vec).data(); \
  }
  if (s == "ls_to_bgcell_to_inoutcut") RET(r->ls_to_bgcell_to_inoutcut[ls]);
  if (s == "cutcell_to_cell") RET(r->cutcell_to_cell);
  if (s == "subcells.cell_to_points") RET(r->subcells.c2p);
  if (s == "subcells.cell_to_bgcell") RET(r->subcells.c2bg);
  if (s == "subcells.point_to_coords") RET(r->subcells.X);
  if (s == "subcells.point_to_rcoords") RET(r->subcells.R);
  if (s == "ls_to_subcell_to_inout") RET(r->ls_to_subcell_to_inout[ls]);
  if (s == "subfacets.facet_to_points") RET(r->subfacets.c2p);
  if (s == "subfacets.facet_to_bgcell") RET(r->subfacets.c2bg);
  if (s == "subfacets.facet_to_normal") RET(r->subfacets.c2n);
  if (s == "subfacets.point_to_coords") RET(r->subfacets.X);
  if (s == "subfacets.point_to_rcoords") RET(r->subfacets.R);
  if (s == "ls_to_subfacet_to_inout") RET(r->ls_to_subfacet_to_inout[ls]);
#undef RET
  *n = -1;
  return nullptr;
}

// compute_bgcell_to_inoutcut(cut,geo) (EmbeddedDiscretizations.jl:91-105)
void gecut_oracle_bgcell_to_inoutcut(void* h, const int32_t* prog, int len, int8_t* out) {
  Result* r = (Result*)h;
  for (int64_t c = 0; c < r->grid.num_cells; ++c) out[c] = eval_inoutcut(prog, len, r->ls_to_bgcell_to_inoutcut, c);
}

// compute_subcell_to_inout(cut,geo) (EmbeddedDiscretizations.jl:325-339)
void gecut_oracle_subcell_to_inout(void* h, const int32_t* prog, int len, int8_t* out) {
  Result* r = (Result*)h;
  for (int64_t c = 0; c < r->subcells.ncells(); ++c) out[c] = eval_inoutcut(prog, len, r->ls_to_subcell_to_inout, c);
}

// Triangulation(cut,CutInOrOut(side),geo) (EmbeddedDiscretizations.jl:303-311): returns the number of selected
// subcells, ids (1-based) in `ids` (capacity Nsc)
int64_t gecut_oracle_select_subcells(void* h, const int32_t* prog, int len, int side, int32_t* ids) {
  Result* r = (Result*)h;
  int64_t n = 0;
  for (int64_t c = 0; c < r->subcells.ncells(); ++c) {
    int8_t bg = eval_inoutcut(prog, len, r->ls_to_bgcell_to_inoutcut, r->subcells.c2bg[c] - 1);
    int8_t sc = eval_inoutcut(prog, len, r->ls_to_subcell_to_inout, c);
    if (bg == CUT && sc == side) ids[n++] = (int32_t)(c + 1);
  }
  return n;
}

// Triangulation(cut,in_or_out::Integer,geo) (:313-317) / ActiveInOrOut (:319-323): bg cell ids (1-based)
// mode 0: state == side; mode 1: state == CUT || state == side
int64_t gecut_oracle_select_bgcells(void* h, const int32_t* prog, int len, int side, int mode, int32_t* ids) {
  Result* r = (Result*)h;
  int64_t n = 0;
  for (int64_t c = 0; c < r->grid.num_cells; ++c) {
    int8_t bg = eval_inoutcut(prog, len, r->ls_to_bgcell_to_inoutcut, c);
    bool keep = mode == 0 ? (bg == side) : (bg == CUT || bg == side);
    if (keep) ids[n++] = (int32_t)(c + 1);
  }
  return n;
}

// EmbeddedBoundary(cut,geo) (:417-434) when prog2 == NULL, EmbeddedBoundary(cut,geo1,geo2) (:442-463) otherwise.
// ids 1-based, orientation +-1
int64_t gecut_oracle_select_boundary(void* h, const int32_t* prog1, int len1, const int32_t* prog2, int len2,
                                     int32_t* ids, int8_t* orientation) {
  Result* r = (Result*)h;
  int64_t n = 0;
  for (int64_t f = 0; f < r->subfacets.ncells(); ++f) {
    int8_t s1, o1;
    eval_inoutboundary(prog1, len1, r->ls_to_subfacet_to_inout, f, &s1, &o1);
    bool keep = s1 == INTERFACE;
    if (keep && prog2 != nullptr) keep = eval_inoutcut(prog2, len2, r->ls_to_subfacet_to_inout, f) == INTERFACE;
    if (keep) {
      ids[n] = (int32_t)(f + 1);
      orientation[n] = o1;
      n++;
    }
  }
  return n;
}

// sum(integrate(1, CellQuadrature(SubCellTriangulation, 2))) per selected subcell
void gecut_oracle_subcell_measures(void* h, const int32_t* ids, int64_t n, double* out) {
  Result* r = (Result*)h;
  for (int64_t i = 0; i < n; ++i) out[i] = integrate_one(r->subcells, ids[i] - 1);
}
void gecut_oracle_subfacet_measures(void* h, const int32_t* ids, int64_t n, double* out) {
  Result* r = (Result*)h;
  for (int64_t i = 0; i < n; ++i) out[i] = integrate_one(r->subfacets, ids[i] - 1);
}
double gecut_oracle_bgcell_measure(void* h, int32_t cell1) { return bgcell_volume(((Result*)h)->grid, cell1 - 1); }

// Laplacian element matrices of the cut cells, subcell contributions summed per bg cell in subcell order
// (move_contributions, SubCellTriangulations.jl:63-72).  K: Ncut x nv x nv (slot = position in cutcell_to_cell), zeroed here.
void gecut_oracle_cutcell_laplacian(void* h, const int32_t* ids, int64_t n, double* K) {
  Result* r = (Result*)h;
  const int nv = r->grid.nv;
  int64_t ncut = (int64_t)r->cutcell_to_cell.size();
  for (int64_t i = 0; i < ncut * nv * nv; ++i) K[i] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t sc0 = ids[i] - 1;
    int32_t bg = r->subcells.c2bg[sc0];
    int64_t slot = std::lower_bound(r->cutcell_to_cell.begin(), r->cutcell_to_cell.end(), bg) - r->cutcell_to_cell.begin();
    laplacian_subcell(*r, sc0, K + slot * nv * nv);
  }
}
void gecut_oracle_bgcell_laplacian(void* h, int32_t cell1, double* K) { bgcell_laplacian(((Result*)h)->grid, cell1 - 1, K); }

// leaf evaluation at arbitrary points (for the K1 parity test)
// ---- cut_facets ---------------------------------------------------------------------------------------------
void* gecut_oracle_cut_facets(const OGrid* grid, const OLeaf* leaves, int L, const double* const* nodal_values) {
  return do_cut_facets(*grid, leaves, L, nodal_values);
}
void gecut_oracle_facets_free(void* h) { delete (FacetResult*)h; }
// sizes: [Nf, Ncutf, Nsf, Nsfp, L, D, nvf, Nboundary]
void gecut_oracle_facets_sizes(void* h, int64_t* out) {
  FacetResult* r = (FacetResult*)h;
  out[0] = (int64_t)r->facet_is_boundary.size();
  out[1] = (int64_t)r->cutfacet_to_facet.size();
  out[2] = r->subfacets.ncells();
  out[3] = r->subfacets.c2p.empty() ? 0 : r->subfacets.npoints();
  out[4] = r->L;
  out[5] = r->grid.D;
  out[6] = r->nvf;
  int64_t nb = 0;
  for (int8_t b : r->facet_is_boundary) nb += b;
  out[7] = nb;
}
const void* gecut_oracle_facets_array(void* h, const char* name, int ls, int64_t* n) {
  FacetResult* r = (FacetResult*)h;
  std::string s(name);
#define RET(vec)                 \
  {                              \
    *n = (int64_t)(vec).size();  \
    return (const void*)(vec).data(); \
  }
  if (s == "ls_to_facet_to_inoutcut") RET(r->ls_to_facet_to_inoutcut[ls]);
  if (s == "cutfacet_to_facet") RET(r->cutfacet_to_facet);
  if (s == "facet_is_boundary") RET(r->facet_is_boundary);
  if (s == "facet_nodes") RET(r->facet_nodes);
  if (s == "subfacets.cell_to_points") RET(r->subfacets.c2p);
  if (s == "subfacets.cell_to_bgcell") RET(r->subfacets.c2bg);
  if (s == "subfacets.point_to_coords") RET(r->subfacets.X);
  if (s == "subfacets.point_to_rcoords") RET(r->subfacets.R);
  if (s == "ls_to_subfacet_to_inout") RET(r->ls_to_subfacet_to_inout[ls]);
#undef RET
  *n = -1;
  return nullptr;
}
// compute_bgfacet_to_inoutcut(cut,geo) / compute_subfacet_to_inout(cut,geo) (EmbeddedFacetDiscretizations.jl:216-248)
void gecut_oracle_bgfacet_to_inoutcut(void* h, const int32_t* prog, int len, int8_t* out) {
  FacetResult* r = (FacetResult*)h;
  for (int64_t f = 0; f < (int64_t)r->facet_is_boundary.size(); ++f)
    out[f] = eval_inoutcut(prog, len, r->ls_to_facet_to_inoutcut, f);
}
void gecut_oracle_facets_subfacet_to_inout(void* h, const int32_t* prog, int len, int8_t* out) {
  FacetResult* r = (FacetResult*)h;
  for (int64_t f = 0; f < r->subfacets.ncells(); ++f) out[f] = eval_inoutcut(prog, len, r->ls_to_subfacet_to_inout, f);
}
// BoundaryTriangulation(facets,cut,in_or_out,geo) (EmbeddedFacetDiscretizations.jl:147-201) on the boundary facets
// (which = 0) or on one side of SkeletonTriangulation (which = 1: interior facets).  mode 0: CutInOrOut(side) ->
// subfacets; mode 1: Integer side -> bg facets with state == side; mode 2: ActiveInOrOut(side) -> bg facets CUT or side.
int64_t gecut_oracle_select_facets(void* h, const int32_t* prog, int len, int which, int mode, int side, int32_t* ids) {
  FacetResult* r = (FacetResult*)h;
  int64_t n = 0;
  auto in_trian = [&](int64_t f0) { return which == 0 ? r->facet_is_boundary[f0] == 1 : r->facet_is_boundary[f0] == 0; };
  if (mode == 0) {
    for (int64_t sf = 0; sf < r->subfacets.ncells(); ++sf) {
      const int64_t f0 = r->subfacets.c2bg[sf] - 1;
      const int a = eval_inoutcut(prog, len, r->ls_to_facet_to_inoutcut, f0);
      const int b = eval_inoutcut(prog, len, r->ls_to_subfacet_to_inout, sf);
      if (in_trian(f0) && a == CUT && b == side) {
        if (ids) ids[n] = (int32_t)(sf + 1);
        n++;
      }
    }
  } else {
    for (int64_t f0 = 0; f0 < (int64_t)r->facet_is_boundary.size(); ++f0) {
      const int a = eval_inoutcut(prog, len, r->ls_to_facet_to_inoutcut, f0);
      const bool ok = mode == 1 ? a == side : (a == CUT || a == side);
      if (in_trian(f0) && ok) {
        if (ids) ids[n] = (int32_t)(f0 + 1);
        n++;
      }
    }
  }
  return n;
}
void gecut_oracle_facets_subfacet_measures(void* h, const int32_t* ids, int64_t n, double* out) {
  FacetResult* r = (FacetResult*)h;
  for (int64_t i = 0; i < n; ++i) out[i] = integrate_one(r->subfacets, ids[i] - 1);
}
// integral of 1 over a whole bg facet (degree-2 rule of the facet polytope: tensor Gauss on the (D-1)-cube, the simplex
// rule on the TRI facets of a simplexified 3-D model; the map is affine)
double gecut_oracle_bgfacet_measure(void* h, int32_t facet1) {
  FacetResult* r = (FacetResult*)h;
  const BgGrid& g = r->grid;
  const int nvf = r->nvf, D = g.D;
  double x[4][3];
  for (int i = 0; i < nvf; ++i) g.node_coords(r->facet_nodes[(int64_t)(facet1 - 1) * nvf + i], x[i]);
  double meas;
  if (D == 2) {
    double a = x[1][0] - x[0][0], b = x[1][1] - x[0][1];
    meas = std::sqrt(a * a + b * b);
  } else {
    double a[3] = {x[1][0] - x[0][0], x[1][1] - x[0][1], x[1][2] - x[0][2]};
    double b[3] = {x[2][0] - x[0][0], x[2][1] - x[0][1], x[2][2] - x[0][2]};
    double n1 = a[1] * b[2] - a[2] * b[1], n2 = a[2] * b[0] - a[0] * b[2], n3 = a[0] * b[1] - a[1] * b[0];
    meas = std::sqrt(n1 * n1 + n2 * n2 + n3 * n3);
  }
  Rule q = g.simplexified ? simplex_rule_deg2(D - 1) : ncube_rule_deg2(D - 1);
  double s = 0;
  for (int i = 0; i < q.n; ++i) s += q.w[i] * meas;
  return s;
}

// ---- AgFEM aggregation (src/AgFEM/CellAggregation.jl:104-242) --------------------------------------------------
// aggregate(AggregateCutCellsByThreshold(threshold), cut, cut_facets, geo, loc): `cell_to_cellin` (Int32, 1-based root
// cell of every active cell, 0 elsewhere).  Topology from the facet grid of the FacetResult (Gridap: cell_to_faces in
// local facet order, face_to_cells ascending).  Returns 0, or 1 when the 20 iterations did not aggregate every cut cell
// (the reference's @assert all_aggregated).
int gecut_oracle_aggregate(void* hc, void* hf, const int32_t* prog, int len, int loc, double threshold, int32_t* out) {
  Result* r = (Result*)hc;
  FacetResult* fr = (FacetResult*)hf;
  const BgGrid& g = r->grid;
  const int D = g.D, nvf = fr->nvf, nv = g.nv;
  const int64_t n_cells = g.num_cells, n_facets = (int64_t)fr->facet_is_boundary.size();
  // get_faces(topo,D,D-1) / get_faces(topo,D-1,D)
  const std::vector<std::vector<int>>& lfacets =
      g.simplexified ? (D == 2 ? LFACETS_TRI : LFACETS_TET) : (D == 2 ? LFACETS_QUAD : LFACETS_HEX);
  const int nlf = (int)lfacets.size();
  std::map<std::array<int64_t, 4>, int32_t> facet_of;
  for (int64_t f = 0; f < n_facets; ++f) {
    std::array<int64_t, 4> key = {-1, -1, -1, -1};
    for (int i = 0; i < nvf; ++i) key[i] = fr->facet_nodes[f * nvf + i];
    std::sort(key.begin(), key.end());
    facet_of[key] = (int32_t)f;
  }
  std::vector<int32_t> cell_to_faces(n_cells * nlf);
  std::vector<std::vector<int32_t>> face_to_cells(n_facets);
  for (int64_t cell = 0; cell < n_cells; ++cell) {
    int64_t nodes[8];
    g.cell_nodes(cell, nodes);
    for (int lf = 0; lf < nlf; ++lf) {
      std::array<int64_t, 4> key = {-1, -1, -1, -1};
      for (int i = 0; i < nvf; ++i) key[i] = nodes[lfacets[lf][i] - 1];
      std::sort(key.begin(), key.end());
      const int32_t f = facet_of.at(key);
      cell_to_faces[cell * nlf + lf] = f;
      face_to_cells[f].push_back((int32_t)cell);
    }
  }
  // _aggregate_by_threshold (:104-131): unit cut measures, states
  std::vector<double> unit(n_cells, 0.0);
  std::vector<int8_t> cell_to_inoutcut(n_cells), facet_to_inoutcut(n_facets);
  for (int64_t c = 0; c < n_cells; ++c) {
    cell_to_inoutcut[c] = eval_inoutcut(prog, len, r->ls_to_bgcell_to_inoutcut, c);
    if (cell_to_inoutcut[c] == loc) unit[c] = bgcell_volume(g, c);
  }
  for (int64_t sc = 0; sc < r->subcells.ncells(); ++sc) {
    const int64_t c = r->subcells.c2bg[sc] - 1;
    if (cell_to_inoutcut[c] == CUT && eval_inoutcut(prog, len, r->ls_to_subcell_to_inout, sc) == loc)
      unit[c] += integrate_one(r->subcells, sc);
  }
  for (int64_t c = 0; c < n_cells; ++c) unit[c] = unit[c] / bgcell_volume(g, c);
  for (int64_t f = 0; f < n_facets; ++f) facet_to_inoutcut[f] = eval_inoutcut(prog, len, fr->ls_to_facet_to_inoutcut, f);
  // _aggregate_by_threshold_barrier (:133-189)
  std::vector<int32_t> cell_to_cellin(n_cells, 0);
  std::vector<char> touched(n_cells, 0);
  for (int64_t c = 0; c < n_cells; ++c)
    if (unit[c] >= threshold) {
      cell_to_cellin[c] = (int32_t)(c + 1);
      touched[c] = 1;
    }
  auto coords = [&](int64_t cell, double x[8][3]) {
    int64_t nodes[8];
    g.cell_nodes(cell, nodes);
    for (int i = 0; i < nv; ++i) {
      x[i][0] = x[i][1] = x[i][2] = 0.0;
      g.node_coords(nodes[i], x[i]);
    }
  };
  bool all_aggregated = false;
  for (int iter = 0; iter < 20; ++iter) {
    all_aggregated = true;
    for (int64_t cell = 0; cell < n_cells; ++cell) {
      if (touched[cell] || cell_to_inoutcut[cell] != CUT) continue;
      // _find_best_neighbor (:191-233)
      double dmin = std::numeric_limits<double>::infinity();
      int64_t best = -1;
      double xi[8][3], xj[8][3];
      coords(cell, xi);
      for (int lf = 0; lf < nlf; ++lf) {
        const int32_t face = cell_to_faces[cell * nlf + lf];
        const int io = facet_to_inoutcut[face];
        if (io != CUT && io != loc) continue;
        for (int32_t neigh : face_to_cells[face]) {
          if (neigh == cell || !touched[neigh]) continue;
          const int64_t cellin = cell_to_cellin[neigh] - 1;
          coords(cellin, xj);
          double d = 0.0;
          for (int p = 0; p < nv; ++p)
            for (int q = 0; q < nv; ++q) {
              double s2 = 0.0;
              for (int c = 0; c < D; ++c) {
                const double w = xi[p][c] - xj[q][c];
                s2 += w * w;
              }
              d = std::max(d, std::sqrt(s2));
            }
          if ((1.0 + 1.0e-9) * d < dmin) {
            dmin = d;
            best = neigh;
          }
        }
      }
      if (best >= 0)
        cell_to_cellin[cell] = cell_to_cellin[best];
      else
        all_aggregated = false;
    }
    if (all_aggregated) break;
    for (int64_t c = 0; c < n_cells; ++c)  // _touch_aggregated_cells! (:235-241)
      if (cell_to_cellin[c] > 0) touched[c] = 1;
  }
  std::memcpy(out, cell_to_cellin.data(), sizeof(int32_t) * n_cells);
  return all_aggregated ? 0 : 1;
}

// GhostSkeleton(cut,in_or_out,geo) (EmbeddedDiscretizations.jl:504-543): facet_to_mask via _fill_ghost_skeleton_mask!
// with face_to_cells of the facet grid of `hf`
void gecut_oracle_ghost_skeleton(void* hc, void* hf, const int32_t* prog, int len, int in_or_out, int8_t* mask) {
  Result* r = (Result*)hc;
  FacetResult* fr = (FacetResult*)hf;
  const BgGrid& g = r->grid;
  const int D = g.D, nvf = fr->nvf;
  const int64_t n_cells = g.num_cells, n_facets = (int64_t)fr->facet_is_boundary.size();
  const std::vector<std::vector<int>>& lfacets =
      g.simplexified ? (D == 2 ? LFACETS_TRI : LFACETS_TET) : (D == 2 ? LFACETS_QUAD : LFACETS_HEX);
  std::map<std::array<int64_t, 4>, int32_t> facet_of;
  for (int64_t f = 0; f < n_facets; ++f) {
    std::array<int64_t, 4> key = {-1, -1, -1, -1};
    for (int i = 0; i < nvf; ++i) key[i] = fr->facet_nodes[f * nvf + i];
    std::sort(key.begin(), key.end());
    facet_of[key] = (int32_t)f;
  }
  std::vector<std::vector<int32_t>> facet_to_cells(n_facets);
  for (int64_t cell = 0; cell < n_cells; ++cell) {
    int64_t nodes[8];
    g.cell_nodes(cell, nodes);
    for (const std::vector<int>& lf : lfacets) {
      std::array<int64_t, 4> key = {-1, -1, -1, -1};
      for (int i = 0; i < nvf; ++i) key[i] = nodes[lf[i] - 1];
      std::sort(key.begin(), key.end());
      facet_to_cells[facet_of.at(key)].push_back((int32_t)cell);
    }
  }
  for (int64_t facet = 0; facet < n_facets; ++facet) {
    int ncells_around_cut = 0, ncells_around_active = 0;
    for (int32_t cell : facet_to_cells[facet]) {
      const int inoutcut = eval_inoutcut(prog, len, r->ls_to_bgcell_to_inoutcut, cell);
      if (inoutcut == CUT) ncells_around_cut += 1;
      if (inoutcut == CUT || inoutcut == in_or_out) ncells_around_active += 1;
    }
    mask[facet] = (ncells_around_cut > 0 && ncells_around_active == 2) ? 1 : 0;
  }
}

void gecut_oracle_eval_leaf(const OLeaf* leaf, int D, const double* x, int64_t n, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = eval_leaf(*leaf, x + i * D, D);
}
void gecut_oracle_node_values(const OGrid* grid, const OLeaf* leaf, double* out) {
  BgGrid g(*grid);
  for (int64_t node = 0; node < g.num_nodes; ++node) {
    double x[3];
    g.node_coords(node, x);
    out[node] = eval_leaf(*leaf, x, g.D);
  }
}
void gecut_oracle_node_coords(const OGrid* grid, double* out) {
  BgGrid g(*grid);
  for (int64_t node = 0; node < g.num_nodes; ++node) g.node_coords(node, out + node * g.D);
}

}  // extern "C"
