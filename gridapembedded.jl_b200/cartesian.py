"""Background models of the host-side mirror: the Cartesian descriptor that Gridap's
`CartesianDiscreteModel(pmin,pmax,partition)` carries, and `simplexify(model)`.

Only what the cutting path reads is kept (assumptions A1, A2, A4):
  * nodes lexicographic, x fastest, `x_d = pmin_d + i_d*((pmax_d-pmin_d)/n_d)`;
  * n-cube cells lexicographic, vertices x fastest;
  * `simplexify` emits the 2 (QUAD) / 6 (HEX) simplices of each n-cube consecutively.
The device kernels rebuild node coordinates from this descriptor; nothing of size O(cells)
is stored on the host.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

SIMPLEXIFY_QUAD = ((1, 2, 3), (2, 3, 4))
SIMPLEXIFY_HEX = ((1, 2, 3, 7), (1, 2, 5, 7), (2, 3, 4, 7), (2, 4, 7, 8), (2, 5, 6, 7), (2, 6, 7, 8))


class CartesianDiscreteModel:
    """`CartesianDiscreteModel(pmin,pmax,partition)` or `CartesianDiscreteModel(domain,partition)`
    with `domain = (x0,x1,y0,y1[,z0,z1])`."""

    def __init__(self, *args, simplexified: bool = False):
        if len(args) == 3:
            pmin, pmax, partition = args
        elif len(args) == 2:
            domain, partition = args
            pmin = tuple(domain[0::2])
            pmax = tuple(domain[1::2])
        else:
            raise TypeError("CartesianDiscreteModel(pmin,pmax,partition) or (domain,partition)")
        self.D = len(partition)
        if self.D not in (2, 3):
            raise NotImplementedError("only 2-D and 3-D background models are supported")
        if len(pmin) != self.D or len(pmax) != self.D:
            raise ValueError("pmin/pmax/partition dimension mismatch")
        self.pmin = tuple(float(x) for x in pmin)
        self.pmax = tuple(float(x) for x in pmax)
        self.partition = tuple(int(n) for n in partition)
        if any(n <= 0 for n in self.partition):
            raise ValueError("partition entries must be positive")
        if any(b <= a for a, b in zip(self.pmin, self.pmax)):
            raise ValueError("pmax must be > pmin in every direction")
        self.simplexified = bool(simplexified)

    # ---- sizes -------------------------------------------------------------------------
    @property
    def num_cubes(self) -> int:
        return int(np.prod(self.partition, dtype=np.int64))

    @property
    def simplices_per_cube(self) -> int:
        return 2 if self.D == 2 else 6

    @property
    def num_cells(self) -> int:
        return self.num_cubes * (self.simplices_per_cube if self.simplexified else 1)

    @property
    def num_nodes(self) -> int:
        return int(np.prod([n + 1 for n in self.partition], dtype=np.int64))

    @property
    def num_vertices_per_cell(self) -> int:
        return self.D + 1 if self.simplexified else 2 ** self.D

    @property
    def sizes(self):
        """`CartesianDescriptor.sizes = (pmax-pmin)/partition`."""
        return tuple((b - a) / n for a, b, n in zip(self.pmin, self.pmax, self.partition))

    def same_nodes(self, other: "CartesianDiscreteModel") -> bool:
        return (self.pmin, self.pmax, self.partition) == (other.pmin, other.pmax, other.partition)

    # ---- small-model helpers (host, tests and host-evaluated leaves) ----------------------
    def node_coordinates(self, planes=None) -> np.ndarray:
        """(num_nodes, D) float64, `x0[d] + (I[d]-1)*dx[d]`.  `planes=(k0,k1)`: only the node planes k0..k1 (inclusive) of
        the last direction -- the nodes a z-slab [k0,k1) of cells touches, in the same (global) order."""
        axes = [self.pmin[d] + np.arange(self.partition[d] + 1, dtype=np.float64) * self.sizes[d]
                for d in range(self.D)]
        if planes is not None:
            k0, k1 = int(planes[0]), int(planes[1])
            if not (0 <= k0 <= k1 <= self.partition[-1]):
                raise ValueError("node plane range out of bounds")
            axes[-1] = axes[-1][k0:k1 + 1]
        grids = np.meshgrid(*axes, indexing="ij")
        # x fastest -> Fortran order of the (i,j,k) index space
        return np.stack([g.reshape(-1, order="F") for g in grids], axis=1)

    def cell_node_ids(self) -> np.ndarray:
        """(num_cells, nv) int64, 1-based node ids."""
        D = self.D
        nn = [n + 1 for n in self.partition]
        idx = np.meshgrid(*[np.arange(n) for n in self.partition], indexing="ij")
        ijk = [g.reshape(-1, order="F") for g in idx]
        cube = []
        for lv in range(2 ** D):
            node = np.zeros_like(ijk[0])
            stride = 1
            for d in range(D):
                node = node + (ijk[d] + ((lv >> d) & 1)) * stride
                stride *= nn[d]
            cube.append(node)
        cube = np.stack(cube, axis=1) + 1
        if not self.simplexified:
            return cube
        lt = np.asarray(SIMPLEXIFY_QUAD if D == 2 else SIMPLEXIFY_HEX) - 1
        return cube[:, lt].reshape(-1, D + 1)

    def __repr__(self):
        return (f"CartesianDiscreteModel(pmin={self.pmin}, pmax={self.pmax}, partition={self.partition}, "
                f"simplexified={self.simplexified})")


def simplexify(model: CartesianDiscreteModel) -> CartesianDiscreteModel:
    """`simplexify(model)` (Gridap, assumption A2)."""
    if model.simplexified:
        return model
    return CartesianDiscreteModel(model.pmin, model.pmax, model.partition, simplexified=True)
