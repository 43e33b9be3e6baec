"""The closed-form facet numbering of the library (csrc/gecut_facets.cuh: the index arithmetic the facet kernels decode
ids with, reached here through the host-only entry point `gecut_facet_topology_host`) against the oracle's first-encounter
numbering (Gridap `generate_cell_to_faces`, DESIGN.md assumption A8): facet -> ordered node tuple and the one-cell-around
flag, n-cube and simplexified models, whole models and z-slabs.  No GPU involved."""
import numpy as np
import pytest

import gecut
from gecut import CartesianDiscreteModel, simplexify, disk, sphere, facet_topology
import oracle_binding as ob

PARTITIONS = [(4, 3), (1, 1), (1, 5), (7, 1), (9, 6), (3, 4, 2), (1, 1, 1), (1, 3, 1), (2, 1, 4), (5, 3, 3)]


def _model(part, sx):
    D = len(part)
    m = CartesianDiscreteModel((0, 1) * D, part)
    return simplexify(m) if sx else m


@pytest.mark.parametrize("sx", [False, True])
@pytest.mark.parametrize("part", PARTITIONS)
def test_facet_numbering_matches_first_encounter(part, sx):
    model = _model(part, sx)
    D = model.D
    geo = disk(0.3, x0=(0.5, 0.5)) if D == 2 else sphere(0.3, x0=(0.5, 0.5, 0.5))
    oc = ob.oracle_cut_facets(model, geo)
    nodes, isb = facet_topology(model)
    assert nodes.shape == (oc.Nf, oc.nvf)
    assert np.array_equal(nodes, oc.array("facet_nodes").reshape(-1, oc.nvf) + 1)
    assert np.array_equal(isb, oc.array("facet_is_boundary"))
    assert int(isb.sum()) == oc.Nboundary


@pytest.mark.parametrize("sx", [False, True])
def test_slab_facet_numbering_is_the_local_model(sx):
    """A z-slab's facet grid is the local model's: layers [k0, k1) as a Cartesian model of its own."""
    for part, (k0, k1) in (((5, 4), (1, 3)), ((3, 2, 6), (2, 5))):
        model = _model(part, sx)
        D = model.D
        local = _model(part[:-1] + (k1 - k0,), sx)
        n1, b1 = facet_topology(model, slab=(k0, k1))
        n2, b2 = facet_topology(local)
        assert np.array_equal(n1, n2) and np.array_equal(b1, b2)


def test_facet_topology_host_argument_checks():
    import ctypes as C
    from gecut import _lib
    from gecut.cutters import make_grid
    L = _lib.lib()
    g = make_grid(CartesianDiscreteModel((0, 1, 0, 1), (4, 4)))
    nf, nvf = C.c_int64(0), C.c_int32(0)
    assert L.gecut_facet_topology_host(C.byref(g), 0, 0, None, None, C.byref(nf), C.byref(nvf)) == _lib.GECUT_OK
    assert (nf.value, nvf.value) == (40, 2)
    assert L.gecut_facet_topology_host(C.byref(g), 39, 2, None, None, None, None) == _lib.GECUT_ERR_INVALID
    assert L.gecut_facet_topology_host(None, 0, 0, None, None, None, None) == _lib.GECUT_ERR_INVALID
    g.slab_begin, g.slab_end = 3, 9
    assert L.gecut_facet_topology_host(C.byref(g), 0, 0, None, None, None, None) == _lib.GECUT_ERR_INVALID
