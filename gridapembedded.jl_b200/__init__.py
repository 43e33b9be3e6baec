"""gecut -- B200-native cutting-and-integration hot path of GridapEmbedded.jl (host-side mirror).

The names below mirror the reference's user API (`src/Exports.jl:8-66`): geometry factories and CSG
booleans, `cut`, `LevelSetCutter`, `Triangulation(cutgeo, PHYSICAL, name)`, `EmbeddedBoundary`,
`Measure`.  All compute goes through the C ABI of `libgecut.so` (include/gecut.h); there is no CPU
fallback -- importing this package works without a GPU, calling `cut` does not.
"""
from . import csg, geometries, cartesian, cutters, discretizations, measures, distributed, facets, agfem  # noqa: F401
from .csg import (Node, Leaf, Geometry, union, intersect, intersection, setdiff, complement, get_geometry,
                  get_geometry_names, get_tree, get_name, get_metadata, replace_metadata, replace_data)
from .geometries import (AnalyticalGeometry, DiscreteGeometry, BoundingBox, discretize, doughnut, popcorn, sphere,
                         disk, cylinder, plane, square, quadrilateral, cube, tube, olympic_rings)
from .cartesian import CartesianDiscreteModel, simplexify
from .cutters import Cutter, LevelSetCutter, cut, compute_bgcell_to_inoutcut
from .discretizations import (IN, OUT, CUT, INTERFACE, CUT_IN, CUT_OUT, PHYSICAL, PHYSICAL_IN, PHYSICAL_OUT, ACTIVE,
                              ACTIVE_IN, ACTIVE_OUT, EmbeddedDiscretization, SubCellData, SubFacetData,
                              Triangulation, EmbeddedBoundary)
from .measures import Measure, integrate_measure, integrate_laplacian, get_cell_measure
from .distributed import DistributedLevelSetCutter, slab_range
from .facets import (EmbeddedFacetDiscretization, cut_facets, compute_bgfacet_to_inoutcut, compute_subfacet_to_inout,
                     BoundaryTriangulation, SkeletonTriangulation, facet_topology)
from .agfem import AggregateCutCellsByThreshold, AggregateAllCutCells, aggregate, aggregate_slabs, GhostSkeleton, GhostSkeleton_slabs
