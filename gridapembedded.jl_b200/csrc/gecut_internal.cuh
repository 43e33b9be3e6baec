// gecut_internal.cuh -- shared declarations of the B200 cutting library (not part of the ABI).
//
// Data layout in HBM (all arrays device resident):
//   ls_to_bgcell_to_inoutcut  L rows of Nc Int8, row pitch padded to 256 B
//   cutmask                   1 bit per bg cell ("CUT by any level set"), grouped as [row][x-word][cell type]
//   cutcell_to_cell           Ncut Int32 (1-based, ascending)
//   SubCellData / SubFacetData arrays exactly as the reference structs (Int32 1-based ids, Float64 AoS points)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "gecut.h"
#include "gecut_tables.h"

#define GC_LMAX GECUT_MAX_LEVELSETS
#define GC_IN (-1)
#define GC_OUT 1
#define GC_CUT 0

// ------------------------------------------------------------------------------------------------
// Grid descriptor handed to kernels (by value).
// ------------------------------------------------------------------------------------------------
struct GcGrid {
  int D;
  int simplexified;
  double pmin[3], dx[3];
  int n[3];    // n-cubes per direction (whole model)
  int nn[3];   // nodes per direction (whole model)
  int k0, k1;  // slab: n-cube layers [k0,k1) of the last direction
  int nv;      // vertices per bg cell
  int nt;      // simplices per n-cube (2 / 6)
  int cpc;     // bg cells per n-cube (1, or nt when simplexified)
  int nt0;     // stage-0 simplices per bg cell (nt for n-cubes, 1 for simplices)
  int64_t layer_cubes;  // n-cubes per layer of the last direction
  int64_t layer_nodes;  // nodes per plane of the last direction
  int64_t num_cubes, num_cells, num_nodes;  // of the slab
};

struct GcLeaves {
  int L;
  gecut_leaf leaf[GC_LMAX];
  const double* nodal[GC_LMAX];  // device pointers for NODAL leaves
  int64_t nodal_base;            // global id of the first node the NODAL vectors hold (0: whole model; slab node offset)
  // classification only: sign words of NODAL leaves written by k_nodal_signs (bit b of word (row, wx) = value of node
  // 32*wx+b of node row `row` is > 0; rows = node planes k0..k1 of the slab x node rows of a plane), else nullptr
  const uint32_t* signs[GC_LMAX];
  int signs_wpr;                 // words per node row
};

// Postfix boolean program over the level-set rows.  Programs with more than one operator also carry a truth table:
// `lut` (device memory, owned by the context cache) has 3^nused entries indexed by the base-3 digits state+1 of the
// level sets the program uses (used[0] = least significant), so that evaluating the CSG tree is one gather.
constexpr int GC_LUT_MAXUSED = 10;  // 3^10 = 59049 bytes
struct GcProgram {
  int len;
  int8_t tok[GECUT_MAX_PROGRAM];
  const int8_t* lut;
  const int8_t* blut;  // boundary programs: (state + 1) | (orientation + 1) << 2 of compute_inoutboundary, same index
  int nused;
  int8_t used[GC_LUT_MAXUSED];
  // memoised results of this program over the subcell / bg-cell state rows of one result (device, owned by the result):
  // when `rows` of an evaluation is sc_base (bg_base) the answer is sc_row[e] (bg_row[e])
  const int8_t *sc_base, *sc_row, *bg_base, *bg_row;
};

// Device pointers of the result layout.
struct GcLayout {
  int8_t* bg_state;
  int64_t bg_pitch;
  int32_t* cutcells;
  int32_t* sc_c2p;
  int32_t* sc_c2bg;
  double* sc_X;
  double* sc_R;
  int8_t* sc_state;
  int64_t sc_pitch;
  int32_t* sf_c2p;
  int32_t* sf_c2bg;
  double* sf_N;
  double* sf_X;
  double* sf_R;
  int8_t* sf_state;
  int64_t sf_pitch;
  // not part of the reference layout: integral of 1 over every subcell / subfacet with the degree-2 rule
  // (sum_q w_q*meas(J)), written by the subtriangulation kernels while the vertices are on chip
  double* sc_meas;
  double* sf_meas;
};

// ------------------------------------------------------------------------------------------------
// Context / result / selection objects behind the opaque ABI handles.
// ------------------------------------------------------------------------------------------------
// per-kernel CUDA-event timing (bench.py roofline leg): off by default, zero cost when off
struct GcProfRec {
  const char* name;
  cudaEvent_t a, b;
};

struct gecut_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // side streams + events: the tree-walk buckets of a multi-level-set cut are independent launches whose run time is set by
  // their slowest thread, so they run concurrently (fork from / join into `stream`)
  cudaStream_t side[8] = {nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[8] = {nullptr};
  std::string err;
  int64_t launches = 0;
  // caching device allocator: blocks freed by results / selections are kept and handed out again to requests of the
  // same size.  A cut step asks for the same sizes every time, so after the first step no driver allocation (and none
  // of its multi-millisecond stalls) happens inside the hot path.  All work of a context is ordered on one stream,
  // which makes immediate reuse safe.
  std::unordered_map<void*, size_t> live;
  std::multimap<size_t, void*> cache;
  size_t cached_bytes = 0;
  bool prof_on = false;
  std::vector<GcProfRec> prof;
  std::string prof_report;
  GcSimplexTable* d_tables = nullptr;  // [3] on the device
  uint8_t* d_simplexify_raw = nullptr;  // [2][6][4]
  uint8_t* d_simplexify_pos = nullptr;  // [2][6][4]
  // truth tables of boolean programs, keyed by the token string (device copy + the host copy the upload reads from)
  std::map<std::string, std::pair<int8_t*, std::vector<int8_t>>> lut_cache;
  void* h_pinned = nullptr;             // small pinned staging buffer for scalar read-backs
  size_t h_pinned_bytes = 0;
  // multi-GPU (gecut_comm.cu): ncclComm_t of this rank, flag + scratch plane of the consistency guard
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  int* comm_flag = nullptr;
  double* comm_plane = nullptr;
  size_t comm_plane_bytes = 0;
};

struct gecut_result {
  gecut_ctx* ctx = nullptr;
  // selections made from this result keep it alive: gecut_result_free with live selections only marks the result, the
  // last gecut_selection_free releases it (host finalizers -- Julia's GC, Python's __del__ -- run in no particular order)
  mutable int live_selections = 0;
  bool free_requested = false;
  GcGrid grid{};
  int L = 0;
  GcLeaves leaves{};
  gecut_sizes sizes{};
  GcLayout lay{};
  uint32_t* cutmask = nullptr;
  uint32_t* cutcount = nullptr;   // cut cells per 32-cube group (k_classify -> compaction, then freed)
  // one level set on simplexified 3-D models: k_classify also counts, per group, the cut TETs whose vertices split 2-2
  // (second channel of `cutcount`) and marks them; the three sizes of the layout are linear in (cut cells, 2-2 cells) and
  // the compaction hands the fill kernel the number of 2-2 cells before every tile, so the L = 1 path needs no count pass,
  // no second scan and no second read-back
  bool want_n22 = false;
  int64_t n22 = 0;
  uint32_t* rowcount = nullptr;   // [2][rows]: cut cells / 2-2 TETs per x-row of cubes (k_classify -> scan -> k_cutmask_emit)
  uint32_t* cut22mask = nullptr;  // which cut TETs split 2-2 (layout of cutmask; k_classify -> k_cutmask_emit, then freed)
  uint32_t* tile_n22 = nullptr;   // 2-2 TETs before every tile of 32 cut cells (k_cutmask_emit -> k_cut1_fill, then freed)
  // memoised program rows (key = program tokens): results of multi-operator boolean programs over the subcell and the
  // bg-cell state rows, evaluated once and shared by every selection / integration that uses the same geometry
  std::map<std::string, std::pair<int8_t*, int8_t*>> prog_rows;
  uint32_t* cut_sc_ptr = nullptr;  // 2*(Ncut+1): (first subcell, first subcell point) (0-based) of every cut bg cell
  unsigned long long* bg_counts_dev = nullptr;  // [L][2][6] IN / OUT bg cells per level set and cell type
  int64_t bg_counts[GC_LMAX * 12] = {0};        // host copy (filled with the cut-cell count read-back)
  // one level set: sums written by the fill kernel ([0] measure of the IN subcells, [1] of the OUT subcells, [2] of the
  // subfacets, [3] number of IN subcells); selections over the single leaf need no pass over the subcells
  double2* cell_vio = nullptr;  // one level set, simplex bg cells: (IN, OUT) subcell measure of every cut cell (fill kernel)
  bool has_l1_stats = false;
  double l1_stats[4] = {0, 0, 0, 0};
  double* l1_stats_pinned() { return l1_stats; }
  bool sf_points_alias = false;    // L == 1: the subfacet point arrays ARE the subcell point arrays on the device
  bool owns_nodal[GC_LMAX] = {false};
  double* nodal_dev[GC_LMAX] = {nullptr};
};

struct gecut_selection {
  const gecut_result* res = nullptr;
  int kind = 0;
  int side = 0;
  GcProgram prog{};
  GcProgram prog2{};
  bool has_prog2 = false;
  int64_t n_sub = 0;  // selected subcells / subfacets
  int64_t n_bg = 0;   // selected bg cells
  int32_t* sub_ids = nullptr;
  int8_t* orientation = nullptr;
  // one level set, single-leaf program: the selected subcells are those with state == lazy_side; count and measure come
  // from the result's tile statistics, the id list is written only when somebody asks for it
  bool sub_lazy = false;
  int lazy_side = 0;
  bool sub_identity = false;  // the selection is 1..n_sub with orientation +1; ids are materialised only on demand
  int32_t* bg_ids = nullptr;  // materialised lazily
  int64_t bg_type_count[8] = {0};
  double* K_cut = nullptr;  // Ncut x nv x nv of the last laplacian integration
};

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
#define GC_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                           \
      return (_e == cudaErrorMemoryAllocation) ? GECUT_ERR_OOM : GECUT_ERR_CUDA;               \
    }                                                                                          \
  } while (0)

#define GC_CHECK(st)          \
  do {                        \
    int _s = (st);            \
    if (_s != GECUT_OK) return _s; \
  } while (0)

static inline int64_t gc_div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

void gc_prof_before(gecut_ctx* ctx, const char* name);
void gc_prof_after(gecut_ctx* ctx);
// every kernel launch of the library goes through this macro (launch counter + optional event timing)
#define GC_LAUNCH(ctx, name, ...)            \
  do {                                       \
    if ((ctx)->prof_on) gc_prof_before(ctx, name); \
    __VA_ARGS__;                             \
    if ((ctx)->prof_on) gc_prof_after(ctx);  \
    (ctx)->launches++;                       \
  } while (0)

int gc_malloc(gecut_ctx* ctx, void** p, size_t bytes);
void gc_free(gecut_ctx* ctx, void* p);
void gc_trim(gecut_ctx* ctx);  // give every cached block back to the driver
int gc_launch_check(gecut_ctx* ctx, const char* what);

// Scratch blocks of one call: handed back to the context's cache on EVERY exit path (GC_CHECK / GC_CUDA return early), so a
// failed call -- typically an allocation failure of the layout at large sizes -- pins nothing and the documented recovery
// (gecut_trim, then a smaller slab) has the memory to work with.  `keep` releases a block from the guard (it became part
// of a result).
struct GcScratch {
  gecut_ctx* ctx;
  std::vector<void*> blocks;
  explicit GcScratch(gecut_ctx* c) : ctx(c) {}
  GcScratch(const GcScratch&) = delete;
  GcScratch& operator=(const GcScratch&) = delete;
  ~GcScratch() {
    for (void* p : blocks) gc_free(ctx, p);
  }
  template <class T>
  int alloc(T** p, size_t bytes) {
    const int st = gc_malloc(ctx, reinterpret_cast<void**>(p), bytes);
    if (st == GECUT_OK) blocks.push_back(*p);
    return st;
  }
  void keep(void* p) {
    for (auto& b : blocks)
      if (b == p) b = nullptr;
  }
};
// per-entity measures of results that did not write them at cut time (simplex bg cells, one level set)
int gc_ensure_measures(gecut_ctx* ctx, const struct gecut_result* res);

// ---- host helpers of gecut_api.cu shared with gecut_facets.cu ----
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};
int gc_make_grid(gecut_ctx* ctx, const gecut_grid* in, GcGrid* g);
int gc_make_program(gecut_ctx* ctx, const int32_t* program, int len, int L, GcProgram* p);

// ---- scan (gecut_classify.cu) ----
// exclusive prefix sum of n uint32 (in place allowed); *total_dev (uint64, device) receives the grand total
int gc_exclusive_scan_u32(gecut_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, uint64_t* total_dev);
// the same for `nch` arrays of n elements stored back to back (one launch set for all channels)
int gc_exclusive_scan_u32_multi(gecut_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int nch, uint64_t* totals_dev);

// ---- classification (gecut_classify.cu) ----
int gc_classify(gecut_ctx* ctx, gecut_result* res);        // states + cutmask
int gc_compact_cutcells(gecut_ctx* ctx, gecut_result* res);  // cutmask -> cutcell_to_cell, sizes.num_cutcells
int gc_eval_leaf_nodes(gecut_ctx* ctx, const GcGrid& g, const gecut_leaf& leaf, double* out_dev, bool slab_only = false);

// ---- subtriangulation (gecut_cutcells.cu) ----
int gc_cut_cells(gecut_ctx* ctx, gecut_result* res);

// ---- selectors (gecut_select.cu) ----
int gc_eval_rows(gecut_ctx* ctx, const int8_t* rows, int64_t pitch, int64_t n, const GcProgram& prog, int8_t* out_dev,
                 bool out_padded = false);  // out_padded: out_dev holds a multiple of 16 bytes
// flags (uint32, n) -> ids (1-based, ascending); the flags buffer is consumed (scanned in place)
int gc_compact_flags(gecut_ctx* ctx, uint32_t* flags, int64_t n, int32_t** ids_out, int64_t* count,
                     const int8_t* orient_all, int8_t** orient_out);
int gc_select_cells(gecut_ctx* ctx, gecut_selection* sel);
int gc_select_boundary(gecut_ctx* ctx, gecut_selection* sel);
int gc_materialize_bg_ids(gecut_ctx* ctx, gecut_selection* sel);
int gc_materialize_sub_ids(gecut_ctx* ctx, gecut_selection* sel);
int gc_selection_gather(gecut_ctx* ctx, const gecut_selection* sel, int32_t* to_points_dev, int32_t* to_bgcell_dev,
                        double* normals_dev);

// ---- integration (gecut_integrate.cu) ----
int gc_integrate_measure(gecut_ctx* ctx, const gecut_selection* sel, double* values_dev, double* total_host);
int gc_integrate_laplacian(gecut_ctx* ctx, gecut_selection* sel, double* K_bg_host);
int gc_cell_measure(gecut_ctx* ctx, const gecut_selection* sel, double* values_dev);

// ------------------------------------------------------------------------------------------------
// device helpers shared by the kernels
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// ---- measures of simplices (Gridap `meas`: |det J| resp. sqrt(det JtJ)); plain operators, the library is compiled
//      with --fmad=false so that host (oracle-like) and device evaluation orders coincide
__host__ __device__ inline double det2x2(const double J[3][3]) { return J[0][0] * J[1][1] - J[0][1] * J[1][0]; }
__host__ __device__ inline double det3x3(const double J[3][3]) {
  return J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
         (J[0][0] * J[1][2] * J[2][1] + J[0][1] * J[1][0] * J[2][2] + J[0][2] * J[1][1] * J[2][0]);
}

// |det J| of a D-simplex / sqrt(det JtJ) of a (D-1)-simplex whose vertices are p[0..ds] (D doubles each)
template <int D, int DS>
__device__ __forceinline__ double simplex_meas(const double (*p)[3]) {
  if (DS == D) {
    double J[3][3];
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int d = 0; d < D; ++d) J[i][d] = p[i + 1][d] - p[0][d];
    return fabs(D == 2 ? det2x2(J) : det3x3(J));
  }
  if (D == 2) {
    double a = p[1][0] - p[0][0], b = p[1][1] - p[0][1];
    return sqrt(a * a + b * b);
  }
  double a[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
  double b[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
  double n1 = a[1] * b[2] - a[2] * b[1], n2 = a[2] * b[0] - a[0] * b[2], n3 = a[0] * b[1] - a[1] * b[0];
  return sqrt(n1 * n1 + n2 * n2 + n3 * n3);
}

// sum_q w_q * meas with the degree-2 rule of the DS-simplex (weights 1/2 x2, 1/6 x3, 1/24 x4)
template <int DS>
__device__ __forceinline__ double integrate_one(double meas) {
  const double w = DS == 1 ? 0.5 : (DS == 2 ? 1.0 / 6 : 1.0 / 24);
  const int nq = DS + 1;
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < nq; ++q) s += w * meas;
  return s;
}


// Leaf formulas (AnalyticalGeometries.jl:75-78,146-198) with explicit round-to-nearest intrinsics: Julia never
// contracts a*b+c into an FMA, so neither may we (IN/OUT of nodes lying on a surface flips on one ulp).
__device__ __forceinline__ double gc_dot3(const double* a, const double* b, int D) {
  double s = __dmul_rn(a[0], b[0]);
  s = __dadd_rn(s, __dmul_rn(a[1], b[1]));
  if (D == 3) s = __dadd_rn(s, __dmul_rn(a[2], b[2]));
  return s;
}

__device__ __forceinline__ double gc_eval_leaf(const gecut_leaf& lf, const double* x, int D) {
  double w[3];
  w[0] = __dsub_rn(x[0], lf.x0[0]);
  w[1] = __dsub_rn(x[1], lf.x0[1]);
  w[2] = (D == 3) ? __dsub_rn(x[2], lf.x0[2]) : 0.0;
  switch (lf.kind) {
    case GECUT_LEAF_SPHERE:
      return __dsub_rn(gc_dot3(w, w, D), __dmul_rn(lf.R, lf.R));
    case GECUT_LEAF_CYLINDER: {
      double A = gc_dot3(w, lf.v, D);
      double H2 = gc_dot3(w, w, D);
      double B = __dsub_rn(H2, __dmul_rn(A, A));
      return __dsub_rn(B, __dmul_rn(lf.R, lf.R));
    }
    case GECUT_LEAF_PLANE:
      return gc_dot3(w, lf.v, D);
    case GECUT_LEAF_DOUGHNUT: {
      double t = __dsub_rn(lf.R, __dsqrt_rn(__dadd_rn(__dmul_rn(w[0], w[0]), __dmul_rn(w[1], w[1]))));
      return __dsub_rn(__dadd_rn(__dmul_rn(t, t), __dmul_rn(w[2], w[2])), __dmul_rn(lf.r, lf.r));
    }
    default:
      return 0.0;
  }
}

// CartesianCoordinates getindex: x0[d] + (I[d]-1)*dx[d] (Gridap, assumption A4)
__device__ __forceinline__ double gc_coord(const GcGrid& g, int d, int i) {
  return __dadd_rn(g.pmin[d], __dmul_rn((double)i, g.dx[d]));
}

// a + c*(b-a)  (set_point_data!, CutTriangulations.jl:123-135)
__device__ __forceinline__ double gc_lerp(double a, double b, double c) {
  return __dadd_rn(a, __dmul_rn(c, __dsub_rn(b, a)));
}

// 3-valued boolean tables (EmbeddedDiscretizations.jl:136-177)
__host__ __device__ __forceinline__ int gc_io_union(int a, int b) {
  if (a == GC_OUT && b == GC_OUT) return GC_OUT;
  if ((a == GC_CUT && b == GC_CUT) || (a == GC_CUT && b == GC_OUT) || (a == GC_OUT && b == GC_CUT)) return GC_CUT;
  return GC_IN;
}
__host__ __device__ __forceinline__ int gc_io_intersection(int a, int b) {
  if (a == GC_IN && b == GC_IN) return GC_IN;
  if ((a == GC_CUT && b == GC_CUT) || (a == GC_CUT && b == GC_IN) || (a == GC_IN && b == GC_CUT)) return GC_CUT;
  return GC_OUT;
}
__host__ __device__ __forceinline__ int gc_io_setdiff(int a, int b) {
  if (a == GC_IN && b == GC_OUT) return GC_IN;
  if ((a == GC_CUT && b == GC_CUT) || (a == GC_CUT && b == GC_OUT) || (a == GC_IN && b == GC_CUT)) return GC_CUT;
  return GC_OUT;
}
__host__ __device__ __forceinline__ int gc_io_complementary(int a) {
  if (a == GC_OUT) return GC_IN;
  if (a == GC_IN) return GC_OUT;
  return GC_CUT;
}

// compute_inoutcut(tree) (EmbeddedDiscretizations.jl:107-134) on element e of `rows`
__device__ __forceinline__ int gc_eval_inoutcut(const GcProgram& p, const int8_t* rows, int64_t pitch, int64_t e) {
  if (p.len == 1) return rows[(int64_t)p.tok[0] * pitch + e];  // single leaf: no stack machine
  if (p.len == 2 && p.tok[1] == GECUT_OP_COMPLEMENT) return gc_io_complementary(rows[(int64_t)p.tok[0] * pitch + e]);
  if (p.sc_row && rows == p.sc_base) return p.sc_row[e];  // memoised program row
  if (p.bg_row && rows == p.bg_base) return p.bg_row[e];
  if (p.lut) {  // truth table: base-3 index from the states of the used level sets
    int idx = 0;
    for (int k = p.nused - 1; k >= 0; --k) idx = idx * 3 + (rows[(int64_t)p.used[k] * pitch + e] + 1);
    return __ldg(p.lut + idx);
  }
  int8_t st[GECUT_MAX_PROGRAM / 2 + 2];
  int sp = 0;
  for (int t = 0; t < p.len; ++t) {
    int tok = p.tok[t];
    if (tok >= 0) {
      st[sp++] = rows[(int64_t)tok * pitch + e];
    } else if (tok == GECUT_OP_COMPLEMENT) {
      st[sp - 1] = (int8_t)gc_io_complementary(st[sp - 1]);
    } else {
      int b = st[--sp], a = st[--sp];
      st[sp++] = (int8_t)(tok == GECUT_OP_UNION ? gc_io_union(a, b)
                                                : (tok == GECUT_OP_INTERSECT ? gc_io_intersection(a, b) : gc_io_setdiff(a, b)));
    }
  }
  return st[0];
}

// same program on explicit per-level-set values
__host__ __device__ __forceinline__ int gc_eval_inoutcut_vals(const GcProgram& p, const int8_t* vals) {
#ifdef __CUDA_ARCH__
  if (p.lut) {
    int idx = 0;
    for (int k = p.nused - 1; k >= 0; --k) idx = idx * 3 + (vals[p.used[k]] + 1);
    return __ldg(p.lut + idx);
  }
#endif
  int8_t st[GECUT_MAX_PROGRAM / 2 + 2];
  int sp = 0;
  for (int t = 0; t < p.len; ++t) {
    int tok = p.tok[t];
    if (tok >= 0) {
      st[sp++] = vals[tok];
    } else if (tok == GECUT_OP_COMPLEMENT) {
      st[sp - 1] = (int8_t)gc_io_complementary(st[sp - 1]);
    } else {
      int b = st[--sp], a = st[--sp];
      st[sp++] = (int8_t)(tok == GECUT_OP_UNION ? gc_io_union(a, b)
                                                : (tok == GECUT_OP_INTERSECT ? gc_io_intersection(a, b) : gc_io_setdiff(a, b)));
    }
  }
  return st[0];
}

// compute_inoutboundary(tree) (EmbeddedDiscretizations.jl:179-245): state and orientation
__device__ __forceinline__ void gc_eval_inoutboundary(const GcProgram& p, const int8_t* rows, int64_t pitch, int64_t e,
                                                      int* state, int* orientation) {
  if (p.len == 1) {  // single leaf: orientation +1
    *state = rows[(int64_t)p.tok[0] * pitch + e];
    *orientation = 1;
    return;
  }
  int8_t st[GECUT_MAX_PROGRAM / 2 + 2], orr[GECUT_MAX_PROGRAM / 2 + 2];
  int sp = 0;
  for (int t = 0; t < p.len; ++t) {
    int tok = p.tok[t];
    if (tok >= 0) {
      st[sp] = rows[(int64_t)tok * pitch + e];
      orr[sp] = 1;
      sp++;
    } else if (tok == GECUT_OP_COMPLEMENT) {
      if (st[sp - 1] == GECUT_INTERFACE) orr[sp - 1] = (int8_t)-orr[sp - 1];
      st[sp - 1] = (int8_t)gc_io_complementary(st[sp - 1]);
    } else {
      int b = st[sp - 1], a = st[sp - 2];
      int d2 = orr[sp - 1], d1 = orr[sp - 2];
      sp -= 2;
      int rs, ro;
      if (tok == GECUT_OP_SETDIFF) {
        rs = gc_io_setdiff(a, b);
        ro = (a == GECUT_INTERFACE) ? d1 : ((b == GECUT_INTERFACE) ? -d2 : d1);
      } else {
        rs = tok == GECUT_OP_UNION ? gc_io_union(a, b) : gc_io_intersection(a, b);
        ro = (a == GECUT_INTERFACE) ? d1 : ((b == GECUT_INTERFACE) ? d2 : d1);
      }
      st[sp] = (int8_t)rs;
      orr[sp] = (int8_t)ro;
      sp++;
    }
  }
  *state = st[0];
  *orientation = orr[0];
}

// the same on explicit per-level-set values (host: builds the boundary truth table)
__host__ __device__ __forceinline__ void gc_eval_inoutboundary_vals(const GcProgram& p, const int8_t* vals, int* state,
                                                                    int* orientation) {
  int8_t st[GECUT_MAX_PROGRAM / 2 + 2], orr[GECUT_MAX_PROGRAM / 2 + 2];
  int sp = 0;
  for (int t = 0; t < p.len; ++t) {
    int tok = p.tok[t];
    if (tok >= 0) {
      st[sp] = vals[tok];
      orr[sp] = 1;
      sp++;
    } else if (tok == GECUT_OP_COMPLEMENT) {
      if (st[sp - 1] == GECUT_INTERFACE) orr[sp - 1] = (int8_t)-orr[sp - 1];
      st[sp - 1] = (int8_t)gc_io_complementary(st[sp - 1]);
    } else {
      int b = st[sp - 1], a = st[sp - 2];
      int d2 = orr[sp - 1], d1 = orr[sp - 2];
      sp -= 2;
      int rs, ro;
      if (tok == GECUT_OP_SETDIFF) {
        rs = gc_io_setdiff(a, b);
        ro = (a == GECUT_INTERFACE) ? d1 : ((b == GECUT_INTERFACE) ? -d2 : d1);
      } else {
        rs = tok == GECUT_OP_UNION ? gc_io_union(a, b) : gc_io_intersection(a, b);
        ro = (a == GECUT_INTERFACE) ? d1 : ((b == GECUT_INTERFACE) ? d2 : d1);
      }
      st[sp] = (int8_t)rs;
      orr[sp] = (int8_t)ro;
      sp++;
    }
  }
  *state = st[0];
  *orientation = orr[0];
}

#endif  // __CUDACC__
