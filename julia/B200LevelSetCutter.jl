# B200LevelSetCutter.jl -- reference-side binding of libgecut.so (SOURCE ONLY: Julia is not available in the build
# image, so this file has never been executed; it documents the drop-in a GridapEmbedded maintainer would add).
#
# It implements the `Cutter` plugin API of the reference (src/Interfaces/Cutters.jl:22-60) for Cartesian background
# models: `cut(cutter, bgmodel, geo)`, `compute_bgcell_to_inoutcut(cutter, bgmodel, geo)`, `cut_facets(cutter, bgmodel, geo)`
# and `compute_bgfacet_to_inoutcut(cutter, bgmodel, geo)`, returning the unchanged
# `EmbeddedDiscretization` layout (src/Interfaces/EmbeddedDiscretizations.jl:36-45) so that Triangulation(cutgeo,...),
# EmbeddedBoundary, AgFEM etc. run on it as they do on the result of the in-tree LevelSetCutter.
module B200LevelSetCutters

using Gridap, Gridap.Geometry, Gridap.Arrays, Gridap.ReferenceFEs
using GridapEmbedded.Interfaces, GridapEmbedded.CSG, GridapEmbedded.LevelSetCutters
using GridapEmbedded.Interfaces: SubCellData, SubFacetData, EmbeddedDiscretization, EmbeddedFacetDiscretization
using GridapEmbedded.LevelSetCutters: _find_unique_leaves, discretize
import GridapEmbedded.Interfaces: cut, compute_bgcell_to_inoutcut, cut_facets, compute_bgfacet_to_inoutcut

const libgecut = get(ENV, "GECUT_LIB", "libgecut.so")

# ---- mirrors of the C structs of include/gecut.h -----------------------------------------------------------------
struct GecutGrid
  D::Int32; simplexified::Int32
  pmin::NTuple{3,Float64}; pmax::NTuple{3,Float64}
  partition::NTuple{3,Int32}; slab_begin::Int32; slab_end::Int32; nodal_slab_local::Int32
end
struct GecutLeaf
  kind::Int32; pad::Int32
  x0::NTuple{3,Float64}; v::NTuple{3,Float64}; R::Float64; r::Float64
end
struct GecutSizes
  num_bgcells::Int64; num_bgnodes::Int64; num_cutcells::Int64; num_subcells::Int64; num_subcell_points::Int64
  num_subfacets::Int64; num_subfacet_points::Int64; cell_offset::Int64; node_offset::Int64
  num_levelsets::Int32; D::Int32; num_vertices_per_bgcell::Int32; subfacet_points_alias::Int32
end
mutable struct GecutBuffers
  ls_to_bgcell_to_inoutcut::Ptr{Int8}; cutcell_to_cell::Ptr{Int32}
  sc_cell_to_points::Ptr{Int32}; sc_cell_to_bgcell::Ptr{Int32}; sc_point_to_coords::Ptr{Float64}; sc_point_to_rcoords::Ptr{Float64}
  ls_to_subcell_to_inout::Ptr{Int8}
  sf_facet_to_points::Ptr{Int32}; sf_facet_to_bgcell::Ptr{Int32}; sf_facet_to_normal::Ptr{Float64}
  sf_point_to_coords::Ptr{Float64}; sf_point_to_rcoords::Ptr{Float64}; ls_to_subfacet_to_inout::Ptr{Int8}
  state_pitch_bgcell::Int64; state_pitch_subcell::Int64; state_pitch_subfacet::Int64
end

const LEAF_NODAL = Int32(4)

"""
    struct B200LevelSetCutter <: Cutter

`device`: CUDA device ordinal.  The context is created lazily and kept for the lifetime of the cutter.
"""
mutable struct B200LevelSetCutter <: Cutter
  device::Int
  ctx::Ptr{Cvoid}
  B200LevelSetCutter(device::Integer=0) = new(device, C_NULL)
end

function _ctx(c::B200LevelSetCutter)
  if c.ctx == C_NULL
    r = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:gecut_create, libgecut), Cint, (Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), c.device, C_NULL, r)
    st == 0 || error("gecut_create failed with status $st (no CUDA device: there is no CPU fallback)")
    c.ctx = r[]
    finalizer(x -> ccall((:gecut_destroy, libgecut), Cint, (Ptr{Cvoid},), x.ctx), c)
  end
  c.ctx
end

_check(c, st, what) = st == 0 || error("$what failed ($st): " *
  unsafe_string(ccall((:gecut_last_error, libgecut), Cstring, (Ptr{Cvoid},), c.ctx)))

_pad3(x, z) = ntuple(i -> i <= length(x) ? x[i] : z, 3)

# Cartesian descriptor of the background model (only CartesianDiscreteModel and simplexify(CartesianDiscreteModel)
# are accelerated; everything else mirrors @notimplementedif of CutTriangulations.jl:495-503)
function _grid(model::CartesianDiscreteModel)
  desc = get_cartesian_descriptor(model)
  D = length(desc.partition)
  pmin = Tuple(desc.origin); pmax = Tuple(desc.origin .+ desc.sizes .* desc.partition)
  GecutGrid(D, 0, _pad3(Float64.(pmin), 0.0), _pad3(Float64.(pmax), 0.0), _pad3(Int32.(desc.partition), Int32(1)), 0, 0, 0)
end

# Leaf records: closed-form leaves are recognised by the closure's captured fields (spherefun, cylinderfun, planefun,
# doughnutfun of src/LevelSetCutters/AnalyticalGeometries.jl:57-59,130-132,167-169,186-188); anything else -- arbitrary
# closures, popcorn, DiscreteGeometry payloads -- is passed as NODAL values evaluated on the host
# (j_to_ls = fun.(point_to_coords), DiscreteGeometries.jl:51).
function _leaf_record(f, point_to_coords)
  n = nameof(typeof(f))
  s = string(n)
  if occursin("spherefun", s) || occursin("diskfun", s)
    return GecutLeaf(0, 0, _pad3(Float64.(Tuple(f.x0)), 0.0), (0.0, 0.0, 0.0), Float64(f.R), 0.0), nothing
  elseif occursin("cylinderfun", s)
    return GecutLeaf(1, 0, _pad3(Float64.(Tuple(f.x0)), 0.0), _pad3(Float64.(Tuple(f.d)), 0.0), Float64(f.R), 0.0), nothing
  elseif occursin("planefun", s)
    return GecutLeaf(2, 0, _pad3(Float64.(Tuple(f.x0)), 0.0), _pad3(Float64.(Tuple(f.v)), 0.0), 0.0, 0.0), nothing
  elseif occursin("doughnutfun", s)
    return GecutLeaf(3, 0, _pad3(Float64.(Tuple(f.x0)), 0.0), (0.0, 0.0, 0.0), Float64(f.R), Float64(f.r)), nothing
  elseif f isa AbstractVector
    return GecutLeaf(LEAF_NODAL, 0, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), 0.0, 0.0), collect(Float64, f)
  else
    return GecutLeaf(LEAF_NODAL, 0, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), 0.0, 0.0), Float64[f(x) for x in point_to_coords]
  end
end

function _run(cutter::B200LevelSetCutter, fname::Symbol, bgmodel, geom, simplexified::Bool, cartesian)
  ctx = _ctx(cutter)
  j_to_fun, oid_to_ls = _find_unique_leaves(get_tree(geom))
  coords = collect1d(get_node_coordinates(get_grid(bgmodel)))
  recs = GecutLeaf[]; nodal = Vector{Union{Nothing,Vector{Float64}}}()
  for f in j_to_fun
    r, v = _leaf_record(f, coords); push!(recs, r); push!(nodal, v)
  end
  g = _grid(cartesian)
  g = GecutGrid(g.D, simplexified ? 1 : 0, g.pmin, g.pmax, g.partition, 0, 0, 0)
  ptrs = Ptr{Float64}[v === nothing ? C_NULL : pointer(v) for v in nodal]
  res = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve nodal begin
    st = ccall((fname, libgecut), Cint,
               (Ptr{Cvoid}, Ref{GecutGrid}, Ptr{GecutLeaf}, Cint, Ptr{Ptr{Float64}}, Cint, Ptr{Ptr{Cvoid}}),
               ctx, g, recs, length(recs), ptrs, 0, res)
  end
  _check(cutter, st, string(fname))
  res[], oid_to_ls
end

# ---- zero-copy host arrays (SURVEY 8f.3) -----------------------------------------------------------------------------------
# The layout arrays are allocated page-locked by the library (gecut_host_alloc) and wrapped in place: gecut_result_copy's
# cudaMemcpyAsync lands directly in the memory that becomes `SubCellData.point_to_coords::Vector{Point{D,Float64}}`
# (`Point{D,Float64}` is isbits, so the wrap is a reinterpretation of the D*n doubles, src/Interfaces/SubCellTriangulations.jl:2-7).
# One pass over the data, at the PCIe rate; no `collect`, no pageable staging.  The memory returns to the library when the
# wrapping Array is garbage collected.
function _pinned(::Type{T}, n::Integer) where {T}
  n == 0 && return Vector{T}(undef, 0)
  r = Ref{Ptr{Cvoid}}(C_NULL)
  st = ccall((:gecut_host_alloc, libgecut), Cint, (Csize_t, Ptr{Ptr{Cvoid}}), n * sizeof(T), r)
  st == 0 || error("gecut_host_alloc($(n * sizeof(T)) bytes) failed with status $st")
  p = r[]
  v = unsafe_wrap(Vector{T}, Ptr{T}(p), n; own = false)
  finalizer(_ -> ccall((:gecut_host_free, libgecut), Cint, (Ptr{Cvoid},), p), v)
  v
end

# L rows of n Int8 in ONE pinned block (gecut_result_copy writes rows at pitch n); every row is its own Vector{Int8}, as the
# reference's Vector{Vector{Int8}} wants, and the block is released when the last row has been collected
function _pinned_rows(n::Integer, L::Integer)
  (n == 0 || L == 0) && return [Vector{Int8}(undef, n) for _ in 1:L], Ptr{Int8}(C_NULL)
  r = Ref{Ptr{Cvoid}}(C_NULL)
  st = ccall((:gecut_host_alloc, libgecut), Cint, (Csize_t, Ptr{Ptr{Cvoid}}), n * L, r)
  st == 0 || error("gecut_host_alloc($(n * L) bytes) failed with status $st")
  p = r[]
  alive = Threads.Atomic{Int}(L)
  rows = [unsafe_wrap(Vector{Int8}, Ptr{Int8}(p) + (l - 1) * n, n; own = false) for l in 1:L]
  for row in rows
    finalizer(row) do _
      Threads.atomic_sub!(alive, 1) == 1 && ccall((:gecut_host_free, libgecut), Cint, (Ptr{Cvoid},), p)
    end
  end
  rows, Ptr{Int8}(p)
end

function _copy_out(cutter, res)
  sz = Ref{GecutSizes}()
  _check(cutter, ccall((:gecut_result_sizes, libgecut), Cint, (Ptr{Cvoid}, Ref{GecutSizes}), res, sz), "gecut_result_sizes")
  s = sz[]; D = Int(s.D); L = Int(s.num_levelsets)
  P = Point{D,Float64}
  bg, bgp   = _pinned_rows(s.num_bgcells, L)
  sci, scip = _pinned_rows(s.num_subcells, L)
  sfi, sfip = _pinned_rows(s.num_subfacets, L)
  c2p = _pinned(Int32, s.num_subcells*(D+1)); c2b = _pinned(Int32, s.num_subcells)
  X = _pinned(P, s.num_subcell_points); R = _pinned(P, s.num_subcell_points)
  f2p = _pinned(Int32, s.num_subfacets*D); f2b = _pinned(Int32, s.num_subfacets)
  N = _pinned(P, s.num_subfacets)
  # one level set: the subfacet points ARE the subcell points (CutTriangulations.jl:365-378): the two structs share the arrays
  alias = s.subfacet_points_alias != 0
  fX = alias ? X : _pinned(P, s.num_subfacet_points); fR = alias ? R : _pinned(P, s.num_subfacet_points)
  dp(v) = Ptr{Float64}(pointer(v))
  b = GecutBuffers(bgp, C_NULL, pointer(c2p), pointer(c2b), dp(X), dp(R), scip,
                   pointer(f2p), pointer(f2b), dp(N), alias ? Ptr{Float64}(C_NULL) : dp(fX),
                   alias ? Ptr{Float64}(C_NULL) : dp(fR), sfip, 0, 0, 0)
  GC.@preserve bg c2p c2b X R sci f2p f2b N fX fR sfi begin
    _check(cutter, ccall((:gecut_result_copy, libgecut), Cint, (Ptr{Cvoid}, Ref{GecutBuffers}), res, b), "gecut_result_copy")
  end
  # Table ptrs are arithmetic (every row has D+1 resp. D entries): 4 bytes per row, filled on the host
  table(data, n, stride) = Table(data, collect(Int32, 1:stride:(n*stride+1)))
  subcells  = SubCellData(table(c2p, s.num_subcells, D+1), c2b, X, R)
  subfacets = SubFacetData(table(f2p, s.num_subfacets, D), N, f2b, fX, fR)
  bg, subcells, sci, subfacets, sfi
end

_cartesian(m::CartesianDiscreteModel) = (m, false)
# simplexify(CartesianDiscreteModel) keeps no descriptor: the caller passes it, see INTEGRATION.md
struct SimplexifiedCartesian{M}; model::M; cartesian::CartesianDiscreteModel; end

function cut(cutter::B200LevelSetCutter, bgmodel::CartesianDiscreteModel, geom)
  res, oid_to_ls = _run(cutter, :gecut_cut, bgmodel, geom, false, bgmodel)
  data = _copy_out(cutter, res)
  ccall((:gecut_result_free, libgecut), Cint, (Ptr{Cvoid},), res)
  EmbeddedDiscretization(bgmodel, data..., oid_to_ls, geom)
end

function cut(cutter::B200LevelSetCutter, bg::SimplexifiedCartesian, geom)
  res, oid_to_ls = _run(cutter, :gecut_cut, bg.model, geom, true, bg.cartesian)
  data = _copy_out(cutter, res)
  ccall((:gecut_result_free, libgecut), Cint, (Ptr{Cvoid},), res)
  EmbeddedDiscretization(bg.model, data..., oid_to_ls, geom)
end

function compute_bgcell_to_inoutcut(cutter::B200LevelSetCutter, bgmodel::CartesianDiscreteModel, geom)
  res, oid_to_ls = _run(cutter, :gecut_classify, bgmodel, geom, false, bgmodel)
  # boolean tree on the device: postfix program over the level-set rows (compute_inoutcut, EmbeddedDiscretizations.jl:107-177)
  prog = Int32[]
  function rec(n)
    if n.leftchild === nothing
      push!(prog, Int32(oid_to_ls[objectid(first(n.data))] - 1))
    else
      rec(n.leftchild); n.rightchild === nothing || rec(n.rightchild)
      push!(prog, Dict(:∪ => Int32(-1), :∩ => Int32(-2), :- => Int32(-3), :! => Int32(-4))[first(n.data)])
    end
  end
  rec(get_tree(geom))
  out = Vector{Int8}(undef, num_cells(bgmodel))
  _check(cutter, ccall((:gecut_bgcell_to_inoutcut, libgecut), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Int8}),
                       res, prog, length(prog), out), "gecut_bgcell_to_inoutcut")
  ccall((:gecut_result_free, libgecut), Cint, (Ptr{Cvoid},), res)
  out
end

# ---- cut_facets (LevelSetCutters.jl:107-110 -> EmbeddedFacetDiscretization, EmbeddedFacetDiscretizations.jl:28-35) -------
struct GecutFacetSizes
  num_bgfacets::Int64; num_boundary_facets::Int64; num_cutfacets::Int64; num_subfacets::Int64
  num_subfacet_points::Int64; num_levelsets::Int32; D::Int32; num_vertices_per_bgfacet::Int32; pad::Int32
end
struct GecutFacetBuffers
  ls_to_facet_to_inoutcut::Ptr{Int8}; cutfacet_to_facet::Ptr{Int32}; facet_to_nodes::Ptr{Int32}
  facet_is_boundary::Ptr{Int8}; sf_cell_to_points::Ptr{Int32}; sf_cell_to_bgcell::Ptr{Int32}
  sf_point_to_coords::Ptr{Float64}; sf_point_to_rcoords::Ptr{Float64}; ls_to_subfacet_to_inout::Ptr{Int8}
end

function _run_facets(cutter::B200LevelSetCutter, bgmodel, geom, classify_only::Bool;
                     simplexified::Bool=false, cartesian=bgmodel)
  ctx = _ctx(cutter)
  j_to_fun, oid_to_ls = _find_unique_leaves(get_tree(geom))
  coords = collect1d(get_node_coordinates(get_grid(bgmodel)))
  recs = GecutLeaf[]; nodal = Vector{Union{Nothing,Vector{Float64}}}()
  for f in j_to_fun
    r, v = _leaf_record(f, coords); push!(recs, r); push!(nodal, v)
  end
  g = _grid(cartesian)
  g = GecutGrid(g.D, simplexified ? 1 : 0, g.pmin, g.pmax, g.partition, 0, 0, 0)
  ptrs = Ptr{Float64}[v === nothing ? C_NULL : pointer(v) for v in nodal]
  res = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve nodal begin
    st = ccall((:gecut_cut_facets, libgecut), Cint,
               (Ptr{Cvoid}, Ref{GecutGrid}, Ptr{GecutLeaf}, Cint, Ptr{Ptr{Float64}}, Cint, Cint, Ptr{Ptr{Cvoid}}),
               ctx, g, recs, length(recs), ptrs, 0, classify_only ? 1 : 0, res)
  end
  _check(cutter, st, "gecut_cut_facets")
  res[], oid_to_ls
end

# The facet ids of the library follow Gridap's own numbering of Grid(ReferenceFE{D-1}, model) (first encounter over
# (cell, local facet), DESIGN.md assumption A8), so `cell_to_bgcell` indexes get_face_nodes(model, D-1) directly.
cut_facets(cutter::B200LevelSetCutter, bgmodel::CartesianDiscreteModel, geom) = _cut_facets(cutter, bgmodel, geom, false, bgmodel)
# simplexified model: the facet grid of the TRI / TET mesh (SEGMENT / TRI bg facets), same first-encounter numbering
cut_facets(cutter::B200LevelSetCutter, bg::SimplexifiedCartesian, geom) = _cut_facets(cutter, bg.model, geom, true, bg.cartesian)

function _cut_facets(cutter::B200LevelSetCutter, bgmodel, geom, simplexified::Bool, cartesian)
  res, oid_to_ls = _run_facets(cutter, bgmodel, geom, false; simplexified=simplexified, cartesian=cartesian)
  sz = Ref{GecutFacetSizes}()
  _check(cutter, ccall((:gecut_facet_result_sizes, libgecut), Cint, (Ptr{Cvoid}, Ref{GecutFacetSizes}), res, sz),
         "gecut_facet_result_sizes")
  s = sz[]; D = Int(s.D); L = Int(s.num_levelsets)
  bg, bgp   = _pinned_rows(s.num_bgfacets, L)
  sfi, sfip = _pinned_rows(s.num_subfacets, L)
  c2p = _pinned(Int32, s.num_subfacets*D); c2b = _pinned(Int32, s.num_subfacets)
  X = _pinned(Point{D,Float64}, s.num_subfacet_points); R = _pinned(Point{D-1,Float64}, s.num_subfacet_points)
  b = GecutFacetBuffers(bgp, C_NULL, C_NULL, C_NULL, pointer(c2p), pointer(c2b), Ptr{Float64}(pointer(X)),
                        Ptr{Float64}(pointer(R)), sfip)
  GC.@preserve bg c2p c2b X R sfi begin
    _check(cutter, ccall((:gecut_facet_result_copy, libgecut), Cint, (Ptr{Cvoid}, Ref{GecutFacetBuffers}), res, b),
           "gecut_facet_result_copy")
  end
  ccall((:gecut_facet_result_free, libgecut), Cint, (Ptr{Cvoid},), res)
  subfacets = SubCellData(Table(c2p, collect(Int32, 1:D:(s.num_subfacets*D+1))), c2b, X, R)
  EmbeddedFacetDiscretization(bgmodel, bg, subfacets, sfi, oid_to_ls, geom)
end

compute_bgfacet_to_inoutcut(cutter::B200LevelSetCutter, bgmodel::CartesianDiscreteModel, geom) =
  _bgfacet_to_inoutcut(cutter, bgmodel, geom, false, bgmodel)
compute_bgfacet_to_inoutcut(cutter::B200LevelSetCutter, bg::SimplexifiedCartesian, geom) =
  _bgfacet_to_inoutcut(cutter, bg.model, geom, true, bg.cartesian)

function _bgfacet_to_inoutcut(cutter::B200LevelSetCutter, bgmodel, geom, simplexified::Bool, cartesian)
  res, oid_to_ls = _run_facets(cutter, bgmodel, geom, true; simplexified=simplexified, cartesian=cartesian)
  prog = _postfix(get_tree(geom), oid_to_ls)
  out = Vector{Int8}(undef, num_facets(bgmodel))
  _check(cutter, ccall((:gecut_bgfacet_to_inoutcut, libgecut), Cint, (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Int8}),
                       res, prog, length(prog), out), "gecut_bgfacet_to_inoutcut")
  ccall((:gecut_facet_result_free, libgecut), Cint, (Ptr{Cvoid},), res)
  out
end

# postfix program of a CSG tree over the 0-based level-set rows (compute_inoutcut, EmbeddedDiscretizations.jl:107-177)
function _postfix(tree, oid_to_ls)
  prog = Int32[]
  function rec(n)
    if n.leftchild === nothing
      push!(prog, Int32(oid_to_ls[objectid(first(n.data))] - 1))
    else
      rec(n.leftchild); n.rightchild === nothing || rec(n.rightchild)
      push!(prog, Dict(:∪ => Int32(-1), :∩ => Int32(-2), :- => Int32(-3), :! => Int32(-4))[first(n.data)])
    end
  end
  rec(tree)
  prog
end

# ---- device-resident quadrature path ---------------------------------------------------------------------------------
# For consumers that do not need the EmbeddedDiscretization on the host at all (volume / surface functionals, cut-cell
# element matrices, aggregation): the result stays on the GPU, selections are boolean programs, only the integrals and
# the matrices come back.  Replaces, for this path, Triangulation(cutgeo,PHYSICAL,geo) + Measure + ∫ / get_cell_measure
# (EmbeddedDiscretizations.jl:297-323, CellAggregation.jl:111-113) and aggregate (CellAggregation.jl:72-189).
mutable struct B200DeviceCut
  cutter::B200LevelSetCutter
  res::Ptr{Cvoid}
  oid_to_ls::Dict{UInt,Int}
  geom
  bgmodel
  simplexified::Bool
  cartesian
end

function device_cut(cutter::B200LevelSetCutter, bgmodel::CartesianDiscreteModel, geom)
  res, oid_to_ls = _run(cutter, :gecut_cut, bgmodel, geom, false, bgmodel)
  d = B200DeviceCut(cutter, res, oid_to_ls, geom, bgmodel, false, bgmodel)
  finalizer(x -> ccall((:gecut_result_free, libgecut), Cint, (Ptr{Cvoid},), x.res), d)
  d
end

function device_cut(cutter::B200LevelSetCutter, bg::SimplexifiedCartesian, geom)
  res, oid_to_ls = _run(cutter, :gecut_cut, bg.model, geom, true, bg.cartesian)
  d = B200DeviceCut(cutter, res, oid_to_ls, geom, bg.model, true, bg.cartesian)
  finalizer(x -> ccall((:gecut_result_free, libgecut), Cint, (Ptr{Cvoid},), x.res), d)
  d
end

_program(d::B200DeviceCut, geo) = _postfix(get_tree(geo), d.oid_to_ls)
_program(d::B200DeviceCut, name::String) = _program(d, get_geometry(d.geom, name))

const SEL_PHYSICAL = Cint(0)

function _select_cells(d::B200DeviceCut, geo, side::Integer)
  prog = _program(d, geo)
  sel = Ref{Ptr{Cvoid}}(C_NULL)
  _check(d.cutter, ccall((:gecut_select_cells, libgecut), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}, Cint, Cint, Ptr{Ptr{Cvoid}}),
                         d.res, SEL_PHYSICAL, prog, length(prog), side, sel), "gecut_select_cells")
  sel[]
end

_free_selection(sel) = ccall((:gecut_selection_free, libgecut), Cint, (Ptr{Cvoid},), sel)

"sum(∫(1)dΩ) over Triangulation(cutgeo,PHYSICAL_IN/OUT,geo), Measure(Ω,2)"
function physical_measure(d::B200DeviceCut, geo=d.geom; side::Integer=IN)
  sel = _select_cells(d, geo, side)
  total = Ref{Float64}(0.0)
  _check(d.cutter, ccall((:gecut_integrate_measure, libgecut), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                         sel, C_NULL, total), "gecut_integrate_measure")
  _free_selection(sel)
  total[]
end

"sum(∫(1)dΓ) over EmbeddedBoundary(cutgeo,geo1[,geo2]), Measure(Γ,2)"
function boundary_measure(d::B200DeviceCut, geo1=d.geom, geo2=nothing)
  p1 = _program(d, geo1)
  p2 = geo2 === nothing ? Int32[] : _program(d, geo2)
  sel = Ref{Ptr{Cvoid}}(C_NULL)
  _check(d.cutter, ccall((:gecut_select_boundary, libgecut), Cint,
                         (Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Int32}, Cint, Ptr{Ptr{Cvoid}}),
                         d.res, p1, length(p1), geo2 === nothing ? C_NULL : pointer(p2), length(p2), sel),
         "gecut_select_boundary")
  total = Ref{Float64}(0.0)
  _check(d.cutter, ccall((:gecut_integrate_measure, libgecut), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                         sel[], C_NULL, total), "gecut_integrate_measure")
  _free_selection(sel[])
  total[]
end

"get_cell_measure(Triangulation(cutgeo,PHYSICAL,geo), Ω_bg): one value per background cell"
function cell_measure(d::B200DeviceCut, geo=d.geom; side::Integer=IN)
  sel = _select_cells(d, geo, side)
  out = Vector{Float64}(undef, num_cells(d.bgmodel))
  _check(d.cutter, ccall((:gecut_cell_measure, libgecut), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), sel, out, 0),
         "gecut_cell_measure")
  _free_selection(sel)
  out
end

"""
Element matrices of ∫(∇(v)⋅∇(u))dΩ for the background shape functions: `K_cut[:,:,k]` is the contribution of the selected
subcells of cut cell `cutcell_to_cell[k]` (move_contributions), `K_bg` the matrix of an uncut cell.
"""
function cutcell_laplacian(d::B200DeviceCut, geo=d.geom; side::Integer=IN)
  sz = Ref(GecutSizes(ntuple(_ -> 0, fieldcount(GecutSizes))...))
  ccall((:gecut_result_sizes, libgecut), Cint, (Ptr{Cvoid}, Ptr{GecutSizes}), d.res, sz)
  nv = 2^num_cell_dims(d.bgmodel)
  ncut = Int(sz[].num_cutcells)
  K_cut = Array{Float64}(undef, nv, nv, ncut)   # column-major: K_cut[b,a,k] = row-major (k,a,b) of the library; symmetric
  K_bg = Array{Float64}(undef, nv, nv)
  sel = _select_cells(d, geo, side)
  _check(d.cutter, ccall((:gecut_integrate_laplacian, libgecut), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                         sel, K_cut, K_bg), "gecut_integrate_laplacian")
  _free_selection(sel)
  K_cut, K_bg
end

"aggregate(AggregateCutCellsByThreshold(threshold), cutgeo, cutgeo_facets, geo, in_or_out) on the device"
function device_aggregate(d::B200DeviceCut, geo=d.geom; side::Integer=IN, threshold::Real=1.0)
  # classification of the facets is all the aggregation needs (simplexified models: the facet grid of the TRI / TET mesh)
  fres, _ = _run_facets(d.cutter, d.bgmodel, d.geom, true; simplexified=d.simplexified, cartesian=d.cartesian)
  prog = _program(d, geo)
  out = Vector{Int32}(undef, num_cells(d.bgmodel))
  # the facet result was cut from the same geometry object, so both discretizations share oid_to_ls: no second program
  st = ccall((:gecut_aggregate, libgecut), Cint,
             (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Cint, Ptr{Int32}, Cint, Cint, Cdouble, Ptr{Int32}, Cint),
             d.res, fres, prog, length(prog), C_NULL, 0, side, threshold, out, 0)
  ccall((:gecut_facet_result_free, libgecut), Cint, (Ptr{Cvoid},), fres)
  _check(d.cutter, st, "gecut_aggregate")
  out
end

"facet_to_mask of GhostSkeleton(cutgeo, in_or_out, geo)"
function ghost_skeleton_mask(d::B200DeviceCut, geo=d.geom; side::Integer=IN)
  prog = _program(d, geo)
  n = Ref{Int64}(0)
  mask = Vector{Int8}(undef, num_faces(d.bgmodel, num_cell_dims(d.bgmodel) - 1))
  _check(d.cutter, ccall((:gecut_ghost_skeleton, libgecut), Cint,
                         (Ptr{Cvoid}, Ptr{Int32}, Cint, Cint, Ptr{Int64}, Ptr{Int8}, Ptr{Int32}),
                         d.res, prog, length(prog), side, n, mask, C_NULL), "gecut_ghost_skeleton")
  Vector{Bool}(mask .!= 0)
end

export B200LevelSetCutter, SimplexifiedCartesian
export B200DeviceCut, device_cut, physical_measure, boundary_measure, cell_measure, cutcell_laplacian
export device_aggregate, ghost_skeleton_mask

end # module
