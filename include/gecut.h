/* gecut.h -- C ABI of the B200-native cutting-and-integration library (libgecut.so).
 *
 * Drop-in boundary for the hot path of gridap/GridapEmbedded.jl v0.9.11
 *     cut(LevelSetCutter(), bgmodel, geo)  +  Triangulation(cutgeo,PHYSICAL) / EmbeddedBoundary(cutgeo) / Measure
 * on Cartesian (optionally simplexified) background models.  A Julia `Cutter` subtype
 * (`src/Interfaces/Cutters.jl:22-60`) reaches the GPU only through `ccall`s of the entry points below
 * (see INTEGRATION.md for the shim); the Python mirror in gridapembedded.jl_b200/ binds the same symbols with ctypes.
 *
 * Conventions
 *   - every function returns an int status (GECUT_OK == 0); no exception, callback or C++ type crosses the ABI;
 *     `gecut_last_error(ctx)` gives the message of the last failure on that context;
 *   - calls are blocking: outputs are complete when the call returns (the work itself runs on the context's stream);
 *   - one context per host thread / per GPU; results, selections belong to the context that made them;
 *   - all ids in the returned layout are 1-based Int32 and states are Int8 exactly as in the reference structs
 *     (`SubCellData` src/Interfaces/SubCellTriangulations.jl:2-7, `SubFacetData` src/Interfaces/SubFacetTriangulations.jl:2-8,
 *     `EmbeddedDiscretization` src/Interfaces/EmbeddedDiscretizations.jl:36-45); coordinates are Float64, AoS (D per point);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with GECUT_ERR_CUDA.
 */
#ifndef GECUT_H
#define GECUT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes -------------------------------------------------------------------------------------------- */
#define GECUT_OK 0
#define GECUT_ERR_INVALID 1        /* bad argument (NULL pointer, bad dimension, bad program ...)                    */
#define GECUT_ERR_NOTIMPLEMENTED 2 /* mirrors @notimplementedif of the reference (CutTriangulations.jl:495-503,538) */
#define GECUT_ERR_CUDA 3           /* CUDA runtime failure (message in gecut_last_error)                             */
#define GECUT_ERR_OOM 4            /* device allocation failed                                                        */
#define GECUT_ERR_OVERFLOW 5       /* an Int32 id of the reference layout would overflow                             */
#define GECUT_ERR_COMM 6           /* NCCL failure (message in gecut_last_error)                                     */

/* ---- states (src/Interfaces/Interfaces.jl:70-73) ----------------------------------------------------------------- */
#define GECUT_IN (-1)
#define GECUT_OUT 1
#define GECUT_CUT 0
#define GECUT_INTERFACE 0

/* ---- leaf kinds = the "bytecode" of AnalyticalGeometry leaves (src/LevelSetCutters/AnalyticalGeometries.jl) -------- */
#define GECUT_LEAF_SPHERE 0   /* _sphere   :146-150  (w.w - R^2), also `disk` in 2-D                                 */
#define GECUT_LEAF_CYLINDER 1 /* _cylinder :176-182  (v already normalised by the constructor, :165)                  */
#define GECUT_LEAF_PLANE 2    /* _plane    :194-198  ((x-x0).v, v raw)                                              */
#define GECUT_LEAF_DOUGHNUT 3 /* _doughnut_fun :75-78                                                               */
#define GECUT_LEAF_NODAL 4    /* values given per background node: DiscreteGeometry (DiscreteGeometries.jl:81-98),     */
                              /* arbitrary closures and popcorn                                                        */
#define GECUT_MAX_LEVELSETS 16

/* ---- postfix boolean program over level-set rows (replaces the recursive compute_inoutcut / compute_inoutboundary,
 *      src/Interfaces/EmbeddedDiscretizations.jl:107-245): token >= 0 pushes row `token`, negative tokens are operators */
#define GECUT_OP_UNION (-1)
#define GECUT_OP_INTERSECT (-2)
#define GECUT_OP_SETDIFF (-3)
#define GECUT_OP_COMPLEMENT (-4)
#define GECUT_MAX_PROGRAM 64

typedef struct gecut_ctx gecut_ctx;             /* device, stream, scratch                                            */
typedef struct gecut_result gecut_result;       /* device-resident EmbeddedDiscretization                             */
typedef struct gecut_selection gecut_selection; /* device-resident Triangulation(cut,...) / EmbeddedBoundary(cut,...)   */

/* CartesianDiscreteModel(pmin,pmax,partition) [simplexify(...)]; `slab_begin/slab_end` restrict the cut to the n-cubes
 * whose index in the LAST direction lies in [slab_begin, slab_end) (0,0 = whole model) -- the z-slab partition that
 * mirrors src/Distributed/DistributedDiscretizations.jl:29-40.  Node coordinates always use the GLOBAL indices so a
 * slab is bit-identical to the same cells of the whole model; cell ids in the result are local to the slab
 * (global id = local id + gecut_sizes.cell_offset). */
typedef struct {
  int32_t D;            /* 2 or 3                                                                                 */
  int32_t simplexified; /* 0: QUAD/HEX cells, 1: TRI/TET cells of simplexify(model)                                 */
  double pmin[3];
  double pmax[3];
  int32_t partition[3];
  int32_t slab_begin;
  int32_t slab_end;
  int32_t nodal_slab_local; /* 0: NODAL vectors hold every node of the model; 1: only the nodes of the slab
                               (planes slab_begin..slab_end of the last direction, num_bgnodes values)            */
} gecut_grid;

/* One unique leaf of the CSG tree, in `_find_unique_leaves` order (DiscreteGeometries.jl:65-74). */
typedef struct {
  int32_t kind;
  int32_t pad;
  double x0[3];
  double v[3];
  double R;
  double r;
} gecut_leaf;

typedef struct {
  int64_t num_bgcells;          /* Nc  (cells of the slab)                                                         */
  int64_t num_bgnodes;          /* Nn  (nodes of the slab incl. its upper plane)                                    */
  int64_t num_cutcells;         /* Ncut: bg cells CUT by at least one level set                                     */
  int64_t num_subcells;         /* Nsc                                                                             */
  int64_t num_subcell_points;   /* Nscp                                                                            */
  int64_t num_subfacets;        /* Nsf                                                                             */
  int64_t num_subfacet_points;  /* Nsfp                                                                            */
  int64_t cell_offset;          /* global bg cell id = local id + cell_offset                                      */
  int64_t node_offset;          /* global node id = local id + node_offset                                          */
  int32_t num_levelsets;        /* L                                                                               */
  int32_t D;
  int32_t num_vertices_per_bgcell; /* nv                                                                           */
  int32_t subfacet_points_alias;   /* 1: one level set -- the subfacet point arrays ARE the subcell point arrays    */
                                   /* (CutTriangulations.jl:365-378); a binding may share one host array for both   */
                                   /* and pass NULL for sf_point_to_coords / sf_point_to_rcoords to gecut_result_copy */
} gecut_sizes;

/* Pointers to the arrays of the layout.  Used both to hand out device pointers (gecut_result_device_ptrs) and to name
 * caller-allocated HOST destinations (gecut_result_copy; NULL entries are skipped).  Row r of an L x N Int8 array
 * starts at base + r*pitch with pitch = N for host copies and gecut_buffers.state_pitch_* for device pointers. */
typedef struct {
  int8_t* ls_to_bgcell_to_inoutcut;   /* L x Nc                                                                    */
  int32_t* cutcell_to_cell;           /* Ncut, ascending, 1-based                                                   */
  int32_t* sc_cell_to_points;         /* Nsc*(D+1)   Table.data (ptrs are 1:(D+1):...)                              */
  int32_t* sc_cell_to_bgcell;         /* Nsc                                                                       */
  double* sc_point_to_coords;         /* Nscp*D                                                                    */
  double* sc_point_to_rcoords;        /* Nscp*D                                                                    */
  int8_t* ls_to_subcell_to_inout;     /* L x Nsc                                                                   */
  int32_t* sf_facet_to_points;        /* Nsf*D                                                                     */
  int32_t* sf_facet_to_bgcell;        /* Nsf                                                                       */
  double* sf_facet_to_normal;         /* Nsf*D                                                                     */
  double* sf_point_to_coords;         /* Nsfp*D                                                                    */
  double* sf_point_to_rcoords;        /* Nsfp*D                                                                    */
  int8_t* ls_to_subfacet_to_inout;    /* L x Nsf                                                                   */
  int64_t state_pitch_bgcell;         /* device row pitches (bytes), filled by gecut_result_device_ptrs            */
  int64_t state_pitch_subcell;
  int64_t state_pitch_subfacet;
} gecut_buffers;

/* ---- context ------------------------------------------------------------------------------------------------- */
/* `cuda_stream` is a cudaStream_t (NULL = the legacy default stream).  Replaces nothing in the reference (it has no
 * device); needed by every other call. */
int gecut_create(int device, void* cuda_stream, gecut_ctx** ctx);
int gecut_destroy(gecut_ctx* ctx);
int gecut_set_stream(gecut_ctx* ctx, void* cuda_stream);
const char* gecut_last_error(const gecut_ctx* ctx);
/* Device memory freed by gecut_result_free / gecut_selection_free is cached by the context and reused by later calls
 * (a cut step asks for the same sizes every time); gecut_trim returns the cached blocks to the driver.  Results and
 * selections must be freed before their context is destroyed. */
int gecut_trim(gecut_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's `gpu_launches`). */
int64_t gecut_launch_count(const gecut_ctx* ctx);

/* Per-kernel CUDA-event timing for the benchmark's roofline leg: gecut_profile(ctx,1) starts recording every kernel
 * launch of this context, gecut_profile_report returns "kernel_name launches total_ms\n" lines (valid until the next
 * call), gecut_profile(ctx,0) stops and clears. */
int gecut_profile(gecut_ctx* ctx, int enable);
const char* gecut_profile_report(gecut_ctx* ctx);

/* ---- cut ------------------------------------------------------------------------------------------------------ */
/* cut(::LevelSetCutter, bgmodel, geo)  (src/LevelSetCutters/LevelSetCutters.jl:79-105 -> CutTriangulations.jl:443-460,
 * 309-343): evaluates the L unique leaves at the nodes, classifies every bg cell per level set, extracts + simplexifies
 * the cut cells and cuts them sequentially by every level set (ls = L..1), producing subcells, boundary subfacets with
 * normals, and the IN/OUT rows.  `nodal_values[ls]` is read only for NODAL leaves: one double per node (node ids
 * lexicographic x fastest) of the whole model, or of the slab when grid->nodal_slab_local is set; resident on the device
 * if `nodal_on_device` else on the host. */
int gecut_cut(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaves, int num_leaves,
              const double* const* nodal_values, int nodal_on_device, gecut_result** out);
/* compute_bgcell_to_inoutcut(::LevelSetCutter, bgmodel, geo) (LevelSetCutters.jl:154-157,175-205): classification only
 * (result has no subcells / subfacets). */
int gecut_classify(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaves, int num_leaves,
                   const double* const* nodal_values, int nodal_on_device, gecut_result** out);
/* discretize(geo, model) (src/LevelSetCutters/DiscreteGeometries.jl:48-63): leaf `leaf` at every node of the model into
 * `values` (Nn doubles, device pointer if on_device else host).  With a slab and grid->nodal_slab_local set: only the node
 * planes slab_begin..slab_end (the slab-local vector gecut_cut accepts, (slab_end-slab_begin+1) planes). */
int gecut_eval_leaf_at_nodes(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaf, double* values,
                             int on_device);
/* Selections made from a result keep it alive: freeing a result that still has selections only marks it, the last
 * gecut_selection_free releases it (finalizers of a garbage-collected host run in no particular order); no new selection
 * can be made from a result after gecut_result_free.  The same holds for facet results and facet selections. */
int gecut_result_free(gecut_result* res);
int gecut_result_sizes(const gecut_result* res, gecut_sizes* sizes);
int gecut_result_device_ptrs(const gecut_result* res, gecut_buffers* dev);
int gecut_result_copy(const gecut_result* res, const gecut_buffers* host);

/* Page-locked host memory for the destinations of gecut_result_copy / gecut_selection_copy / the element matrices: a
 * binding wraps the pointer in its own array type without a copy (Julia: unsafe_wrap(Vector{Point{D,Float64}}, ptr, n) feeds
 * SubCellData, src/Interfaces/SubCellTriangulations.jl:2-7, directly) and the device -> host copy runs at the PCIe rate in one
 * pass (pageable destinations are staged by the driver at a fraction of it).  Portable across contexts; no context needed.
 * gecut_host_register pins memory the host language allocated itself (e.g. a Julia Vector the GC owns). */
int gecut_host_alloc(size_t bytes, void** ptr);
int gecut_host_free(void* ptr);
int gecut_host_register(void* ptr, size_t bytes);
int gecut_host_unregister(void* ptr);

/* ---- selectors -------------------------------------------------------------------------------------------------- */
/* compute_bgcell_to_inoutcut(cut, geo) (src/Interfaces/EmbeddedDiscretizations.jl:91-105) -> Nc Int8 on the host */
int gecut_bgcell_to_inoutcut(const gecut_result* res, const int32_t* program, int len, int8_t* host_out);
/* compute_subcell_to_inout(cut, geo) (EmbeddedDiscretizations.jl:325-339) -> Nsc Int8 on the host */
int gecut_subcell_to_inout(const gecut_result* res, const int32_t* program, int len, int8_t* host_out);

#define GECUT_SEL_PHYSICAL 0 /* Triangulation(cut,(CutInOrOut(side),side),geo)  (:297-317)                           */
#define GECUT_SEL_ACTIVE 1   /* Triangulation(cut,ActiveInOrOut(side),geo)      (:319-323)                           */
#define GECUT_SEL_BOUNDARY 2 /* EmbeddedBoundary(cut,geo1[,geo2])               (:417-463)                           */
#define GECUT_SEL_CUTONLY 3  /* Triangulation(cut,CutInOrOut(side),geo)         (:303-311) subcells only             */
#define GECUT_SEL_BGONLY 4   /* Triangulation(cut,side::Integer,geo)            (:313-317) bg cells only             */
/* side = GECUT_IN or GECUT_OUT.  PHYSICAL: sub-entities = subcells with bg state CUT and subcell state == side,
 * bg entities = bg cells with state == side.  ACTIVE: bg cells with state CUT or side. */
int gecut_select_cells(const gecut_result* res, int kind, const int32_t* program, int len, int side,
                       gecut_selection** out);
/* program2 == NULL: EmbeddedBoundary(cut,geo1); else the two-geometry form. */
int gecut_select_boundary(const gecut_result* res, const int32_t* program1, int len1, const int32_t* program2,
                          int len2, gecut_selection** out);
int gecut_selection_free(gecut_selection* sel);
/* n_sub: selected subcells / subfacets, n_bg: selected bg cells */
int gecut_selection_sizes(const gecut_selection* sel, int64_t* n_sub, int64_t* n_bg);
/* host copies (NULL skipped): 1-based ids of the selected sub-entities (SubCellData(st,ids) / SubFacetData(st,ids,o),
 * SubCellTriangulations.jl:9-17, SubFacetTriangulations.jl:10-21), boundary orientation (+-1), selected bg cell ids */
int gecut_selection_copy(const gecut_selection* sel, int32_t* sub_ids, int8_t* orientation, int32_t* bg_ids);
/* re-indexed tables of the selection on the host (NULL skipped): connectivity rows, bgcell, oriented normals */
int gecut_selection_gather(const gecut_selection* sel, int32_t* to_points, int32_t* to_bgcell, double* normals);

/* ---- integration (Gridap Measure(trian,2) semantics) ---------------------------------------- */
/* sum(∫(1)dΩ) resp. sum(∫(1)dΓ): `values` (host, NULL skipped) gets one value per selected sub-entity; *total also
 * includes the selected uncut bg cells (tensor Gauss / simplex rule per cell). */
int gecut_integrate_measure(const gecut_selection* sel, double* values, double* total);
/* ∫(∇v·∇u)dΩ element matrices of the bg shape functions (Q1 on QUAD/HEX, P1 on TRI/TET), contributions of the selected
 * subcells accumulated per cut bg cell in subcell order (move_contributions, SubCellTriangulations.jl:63-72):
 * `K_cut` (host, NULL skipped) is Ncut x nv x nv, row k belongs to cutcell_to_cell[k] (zero if the cell has no selected
 * subcell); `K_bg` (host, NULL skipped) is ntypes x nv x nv for the uncut cells (ntypes = 1 for n-cubes, 2/6 for
 * simplexified models, type = (cell-1) mod ntypes).  With K_cut == NULL the matrices stay on the device
 * (gecut_selection_device_laplacian) and the call is STREAM-ORDERED: it returns once the kernel is queued on the
 * context's stream; later calls on the context and work queued on that stream see the finished matrices. */
int gecut_integrate_laplacian(const gecut_selection* sel, double* K_cut, double* K_bg);
/* get_cell_measure(trian, bgtrian) (the aggregation input of AgFEM, src/AgFEM/CellAggregation.jl:111-113): measure of the
 * selected triangulation per bg cell -- the whole cell for selected uncut cells, the sum over the selected subcells (in
 * subcell order) for cut cells, 0 elsewhere.  `values` has num_bgcells entries; host memory, or device memory when
 * values_on_device != 0.  PHYSICAL, CUT-only and bg-only selections. */
int gecut_cell_measure(const gecut_selection* sel, double* values, int values_on_device);
/* device pointer to the Ncut x nv x nv matrices of the last gecut_integrate_laplacian on this selection */
int gecut_selection_device_laplacian(const gecut_selection* sel, double** K_cut_dev);

/* ---- cut_facets ----------------------------------------------------------------------------- */
/* cut_facets(::LevelSetCutter, bgmodel, geo) (src/LevelSetCutters/LevelSetCutters.jl:107-142): the marching-simplex
 * machinery on Grid(ReferenceFE{D-1}, bgmodel) without boundary -> EmbeddedFacetDiscretization
 * (src/Interfaces/EmbeddedFacetDiscretizations.jl:28-35).  Cartesian models, n-cube or simplexified (the facet grid of the
 * TRI / TET mesh: SEGMENT / TRI bg facets, cut without a stage-0 simplexification).  With grid->slab_begin/end the result is
 * the facet discretization of the rank's LOCAL model -- layers [slab_begin, slab_end) as a Cartesian model of its own (local
 * facet numbering; its lower and upper planes are boundary facets of the local model), node coordinates global, so a slab's
 * facets are bit-identical to the same facets of the whole model (the reference's ranks cut their local models the same way,
 * src/Distributed/DistributedDiscretizations.jl:42-53).
 * Facet numbering (Gridap `generate_cell_to_faces`, assumption A8 in DESIGN.md): first encounter over (cell ascending,
 * local facet ascending), local facets QUAD [1,2],[3,4],[1,3],[2,4] / HEX [1,2,3,4],[5,6,7,8],[1,2,5,6],[3,4,7,8],
 * [1,3,5,7],[2,4,6,8], TRI [1,2],[1,3],[2,3] / TET [1,2,3],[1,2,4],[1,3,4],[2,3,4]; a facet keeps the node order of the local
 * facet of the cell that meets it first.  Both numberings are closed-form on a Cartesian model (no topology tables). */
typedef struct gecut_facet_result gecut_facet_result;       /* device-resident EmbeddedFacetDiscretization          */
typedef struct gecut_facet_selection gecut_facet_selection; /* BoundaryTriangulation / SkeletonTriangulation(cut_facets,...) */

typedef struct {
  int64_t num_bgfacets;         /* Nf                                                                              */
  int64_t num_boundary_facets;  /* facets with one cell around                                                     */
  int64_t num_cutfacets;        /* facets CUT by at least one level set                                            */
  int64_t num_subfacets;        /* Nsf: (D-1)-simplices of the cut facets                                          */
  int64_t num_subfacet_points;  /* Nsfp                                                                            */
  int32_t num_levelsets;
  int32_t D;
  int32_t num_vertices_per_bgfacet; /* 2 (SEGMENT), 4 (QUAD) or 3 (TRI, simplexified 3-D models)                    */
  int32_t pad;
} gecut_facet_sizes;

/* The closed-form facet numbering on the HOST (index arithmetic only, no device, no context): 1-based nodes
 * (`*nodes_per_facet` per facet, facet node order) and boundary flags of the bg facets [first, first + count) (0-based ids) of
 * the model -- of the slab's local model when grid->slab_begin/end are set.  NULL outputs are skipped; count = 0 just returns
 * the sizes.  The same functions the kernels decode facet ids with (csrc/gecut_facets.cuh); tests/test_facet_numbering.py pins
 * them to the oracle's first-encounter numbering (Gridap `generate_cell_to_faces`) without a GPU. */
int gecut_facet_topology_host(const gecut_grid* grid, int64_t first, int64_t count, int32_t* facet_to_nodes,
                              int8_t* facet_is_boundary, int64_t* num_facets, int32_t* nodes_per_facet);

/* host destinations (NULL entries skipped) of gecut_facet_result_copy */
typedef struct {
  int8_t* ls_to_facet_to_inoutcut;  /* L x Nf                                                                      */
  int32_t* cutfacet_to_facet;       /* Ncutf, ascending, 1-based                                                    */
  int32_t* facet_to_nodes;          /* Nf x nvf, 1-based node ids (Grid(ReferenceFE{D-1},model) connectivity)       */
  int8_t* facet_is_boundary;        /* Nf                                                                          */
  int32_t* sf_cell_to_points;       /* Nsf*D   SubCellData{D-1,D}.cell_to_points data                              */
  int32_t* sf_cell_to_bgcell;       /* Nsf     bg FACET ids                                                        */
  double* sf_point_to_coords;       /* Nsfp*D                                                                      */
  double* sf_point_to_rcoords;      /* Nsfp*(D-1)  reference coordinates in the bg facet                           */
  int8_t* ls_to_subfacet_to_inout;  /* L x Nsf                                                                     */
} gecut_facet_buffers;

/* classify_only != 0: compute_bgfacet_to_inoutcut(::LevelSetCutter, bgmodel, geo) (LevelSetCutters.jl:169-173) -- states
 * and the cut facet list, no sub-facets. */
int gecut_cut_facets(gecut_ctx* ctx, const gecut_grid* grid, const gecut_leaf* leaves, int num_leaves,
                     const double* const* nodal_values, int nodal_on_device, int classify_only,
                     gecut_facet_result** out);
int gecut_facet_result_free(gecut_facet_result* res);
int gecut_facet_result_sizes(const gecut_facet_result* res, gecut_facet_sizes* sizes);
int gecut_facet_result_copy(const gecut_facet_result* res, const gecut_facet_buffers* host);
/* compute_bgfacet_to_inoutcut(cut, geo) / compute_subfacet_to_inout(cut, geo) (EmbeddedFacetDiscretizations.jl:216-248) */
int gecut_bgfacet_to_inoutcut(const gecut_facet_result* res, const int32_t* program, int len, int8_t* host_out);
int gecut_subfacet_to_inout(const gecut_facet_result* res, const int32_t* program, int len, int8_t* host_out);

#define GECUT_FACETS_BOUNDARY 0 /* BoundaryTriangulation(cut_facets, in_or_out, geo)   (:112-201)                      */
#define GECUT_FACETS_SKELETON 1 /* one side of SkeletonTriangulation(cut_facets, in_or_out, geo) (:54-103)             */
/* kind = GECUT_SEL_PHYSICAL (sub-facets of CUT facets with state == side + facets with state == side), GECUT_SEL_CUTONLY,
 * GECUT_SEL_BGONLY or GECUT_SEL_ACTIVE (facets CUT or side), restricted to the boundary / interior facets. */
int gecut_select_facets(const gecut_facet_result* res, int which, int kind, const int32_t* program, int len, int side,
                        gecut_facet_selection** out);
int gecut_facet_selection_free(gecut_facet_selection* sel);
int gecut_facet_selection_sizes(const gecut_facet_selection* sel, int64_t* n_sub, int64_t* n_bg);
/* host copies (NULL skipped): 1-based ids of the selected sub-facets / bg facets, ascending */
int gecut_facet_selection_copy(const gecut_facet_selection* sel, int32_t* sub_ids, int32_t* bg_ids);
/* sum(∫(1)dΓ) over the selection with Measure(trian,2): `sub_values` (host, NULL skipped) one value per selected
 * sub-facet; *total also includes the selected whole facets */
int gecut_facet_selection_measure(const gecut_facet_selection* sel, double* sub_values, double* total);

/* ---- AgFEM aggregation ------------------------------------------------------------------------ */
/* aggregate(AggregateCutCellsByThreshold(threshold), cut, cut_facets, geo, in_or_out) (src/AgFEM/CellAggregation.jl:72-189):
 * `cell_to_cellin` (num_bgcells Int32; host memory, or device memory when on_device != 0) receives for every cell the
 * 1-based root cell it is aggregated to -- itself for cells whose unit cut measure is >= threshold, the root of the best
 * touched face neighbour (closest root, ties by Gridap's local facet order) for the other CUT cells, 0 elsewhere.
 * `facets` may be a classification-only facet result (compute_bgfacet_to_inoutcut).  `facet_program` is the same geometry
 * compiled over the FACET discretization's own level-set numbering (its oid_to_ls, CellAggregation.jl:90); NULL = the
 * two discretizations number their level sets alike and `program` serves both.  threshold = 1 is
 * AggregateAllCutCells().  Fails with GECUT_ERR_INVALID when 20 sweeps do not aggregate every cut cell (the reference's
 * @assert all_aggregated).  Whole Cartesian models, n-cube or simplexified (cells = the simplices, neighbours through the
 * facet grid of the simplices; `cut` and `facets` must come from the same model). */
int gecut_aggregate(const gecut_result* cut, const gecut_facet_result* facets, const int32_t* program, int len,
                    const int32_t* facet_program, int facet_len, int side, double threshold, int32_t* cell_to_cellin,
                    int on_device);
/* The same over z-slabs (src/Distributed/DistributedAgFEM.jl:52-144, `_distributed_aggregate_by_threshold_barrier`): every slab
 * sweeps its own cut cells, the cell layers two slabs share as ghosts (touched flags and roots) are refreshed after every sweep,
 * the loop ends when no slab has an unaggregated cut cell.  A sweep only reads state frozen during the sweep, so the roots equal
 * the whole-model aggregation's.  `cell_to_cellin[r]` (num_bgcells of slab r) receives GLOBAL 1-based root cell ids, as the
 * reference stores them (:101-105).  `cuts[r]` / `facets[r]`: results of gecut_cut / gecut_cut_facets with the same slab, all in
 * one context, in ascending layer order, together covering the model (a model cut slab by slab on one GPU). */
int gecut_aggregate_slabs(const gecut_result* const* cuts, const gecut_facet_result* const* facets, int nslabs,
                          const int32_t* program, int len, const int32_t* facet_program, int facet_len, int side,
                          double threshold, int32_t* const* cell_to_cellin, int on_device);

/* GhostSkeleton(cut, in_or_out, geo) (src/Interfaces/EmbeddedDiscretizations.jl:504-543): the interior facets with at
 * least one CUT cell around and both cells active (CUT or `side`; side = GECUT_IN / GECUT_OUT for ACTIVE_IN / ACTIVE_OUT,
 * GECUT_CUT for the cut cells alone).  *n_ghost = number of such facets; `facet_to_mask` (host, num_bgfacets Int8, NULL
 * skipped) is the mask handed to SkeletonTriangulation(model, facet_to_mask); `facet_ids` (host, *n_ghost Int32 from a
 * previous call, NULL skipped) the same facets as an ascending 1-based list.  Facet numbering as in gecut_cut_facets.
 * Whole Cartesian models, n-cube or simplexified. */
int gecut_ghost_skeleton(const gecut_result* cut, const int32_t* program, int len, int side, int64_t* n_ghost,
                         int8_t* facet_to_mask, int32_t* facet_ids);
/* The same over z-slabs of one context (in ascending layer order, covering the model): `facet_to_mask[r]` (host, NULL skipped)
 * is the mask of slab r's LOCAL model -- the facet numbering of gecut_cut_facets with the same slab -- and n_ghost[r] the
 * number of its marked facets.  A facet on the plane two slabs share is an interior facet of the model: the cell on its other
 * side is read from the adjacent slab, and the facet is marked (and counted) in both local models. */
int gecut_ghost_skeleton_slabs(const gecut_result* const* cuts, int nslabs, const int32_t* program, int len, int side,
                               int64_t* n_ghost, int8_t* const* facet_to_mask);

/* ---- multi-GPU: z-slab partition, one context per GPU / process ------------------------------ */
/* cut(cutter, bgmodel::DistributedDiscreteModel, geo) (src/Distributed/DistributedDiscretizations.jl:33-40): every rank
 * cuts its own slab (gecut_grid.slab_begin/end) -- no data-path collective.  What does cross ranks is enqueued on the
 * context's stream by the calls below (NCCL over NVLink); a host language needs nothing but these entry points:
 *   rank 0: gecut_comm_unique_id(id); the host program distributes the GECUT_COMM_ID_BYTES bytes with what it already has
 *   (MPI.Bcast in GridapDistributed, a file, torch.distributed ...); every rank: gecut_comm_init(ctx, id, rank, world). */
#define GECUT_COMM_ID_BYTES 128
#define GECUT_COMM_MAX_VALUES 16
int gecut_comm_unique_id(void* id);
int gecut_comm_init(gecut_ctx* ctx, const void* id, int rank, int world);
int gecut_comm_free(gecut_ctx* ctx);
int gecut_comm_rank(const gecut_ctx* ctx, int* rank, int* world);
/* Ghost update of level sets that exist only as slab-local nodal vectors (`consistent!`, :164-170): `slab_values[i]`
 * (device) holds num_planes = slab_end - slab_begin + 1 node planes of layer_nodes doubles; the first plane goes to rank-1,
 * the last plane is overwritten with the first plane of rank+1 (the last rank owns its top plane).  NULL entries are
 * skipped.  Stream-ordered: returns once the grouped send/recv is queued; a following gecut_cut on this context sees it. */
int gecut_comm_exchange_ghost_planes(gecut_ctx* ctx, double* const* slab_values, int num_vectors, int64_t layer_nodes,
                                     int64_t num_planes);
/* isconsistent_bgcell_to_inoutcut (:136-174): the reference compares every rank's ghost-cell states with the owner's and
 * reduces with &.  In the slab partition the only data two ranks both classify with is their shared node plane, so the
 * guard compares the SIGNS (what compute_case reads) of this rank's copy (last plane) with the owner's (first plane of
 * rank+1) for every vector, then all-reduces the flag with min.  Closed-form leaves are recomputed from global indices
 * and cannot disagree: pass num_vectors = 0 (the flag is still reduced, so a rank that failed elsewhere may lower it).
 * Blocking; *consistent = 1 or 0 on every rank. */
int gecut_comm_check_consistency(gecut_ctx* ctx, const double* const* slab_values, int num_vectors, int64_t layer_nodes,
                                 int64_t num_planes, int* consistent);
/* sum over the ranks of n <= GECUT_COMM_MAX_VALUES doubles (global integrals and counts, the `sum(∫...)` of a distributed
 * triangulation): `host_values` (NULL: result_dev already holds the local numbers) are placed in `result_dev` (device) as
 * kernel arguments and all-reduced in place.  Stream-ordered, no host synchronisation; read with gecut_comm_fetch. */
int gecut_comm_allreduce_sum(gecut_ctx* ctx, const double* host_values, int n, double* result_dev);
int gecut_comm_fetch(gecut_ctx* ctx, const double* values_dev, int n, double* host_out);
/* sizes of every rank's result (all_sizes[world]): offsets of the rank-major global layout = exclusive sums.  Blocking. */
int gecut_comm_allgather_sizes(gecut_ctx* ctx, const gecut_result* res, gecut_sizes* all_sizes);
/* gecut_aggregate_slabs with one slab per rank (ranks own the z-slabs in ascending order): ghost cell layers by grouped
 * ncclSend / ncclRecv with rank-1 / rank+1 after every sweep (`consistent!`, DistributedAgFEM.jl:124-128), termination by
 * ncclAllReduce (`reduction!(&, all_aggregated)`, :130).  Collective: every rank calls it.  Global root ids. */
/* gecut_ghost_skeleton_slabs with one slab per rank: the state rows of the boundary cell layers travel by grouped
 * ncclSend / ncclRecv.  Collective. */
int gecut_comm_ghost_skeleton(const gecut_result* cut, const int32_t* program, int len, int side, int64_t* n_ghost,
                              int8_t* facet_to_mask);
int gecut_comm_aggregate(const gecut_result* cut, const gecut_facet_result* facets, const int32_t* program, int len,
                         const int32_t* facet_program, int facet_len, int side, double threshold, int32_t* cell_to_cellin,
                         int on_device);

#ifdef __cplusplus
}
#endif
#endif /* GECUT_H */
