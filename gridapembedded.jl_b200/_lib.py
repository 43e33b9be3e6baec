"""ctypes binding of the C ABI (include/gecut.h) -- the only way the host mirror reaches the GPU.

The shared library is built in-tree by `build_library()` (nvcc, sm_100a) into
`gridapembedded.jl_b200/lib/libgecut.so`.  If it is missing or cannot be loaded every compute call raises:
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libgecut.so")
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(HERE, "..", "include")

GECUT_OK = 0
GECUT_ERR_INVALID, GECUT_ERR_NOTIMPLEMENTED, GECUT_ERR_CUDA, GECUT_ERR_OOM, GECUT_ERR_OVERFLOW = 1, 2, 3, 4, 5
SEL_PHYSICAL, SEL_ACTIVE, SEL_BOUNDARY, SEL_CUTONLY, SEL_BGONLY = 0, 1, 2, 3, 4
MAX_LEVELSETS = 16
MAX_PROGRAM = 64


class GecutError(RuntimeError):
    pass


class Grid(C.Structure):
    _fields_ = [("D", C.c_int32), ("simplexified", C.c_int32), ("pmin", C.c_double * 3), ("pmax", C.c_double * 3),
                ("partition", C.c_int32 * 3), ("slab_begin", C.c_int32), ("slab_end", C.c_int32),
                ("nodal_slab_local", C.c_int32)]


class LeafRec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("x0", C.c_double * 3), ("v", C.c_double * 3),
                ("R", C.c_double), ("r", C.c_double)]


class Sizes(C.Structure):
    _fields_ = [("num_bgcells", C.c_int64), ("num_bgnodes", C.c_int64), ("num_cutcells", C.c_int64),
                ("num_subcells", C.c_int64), ("num_subcell_points", C.c_int64), ("num_subfacets", C.c_int64),
                ("num_subfacet_points", C.c_int64), ("cell_offset", C.c_int64), ("node_offset", C.c_int64),
                ("num_levelsets", C.c_int32), ("D", C.c_int32), ("num_vertices_per_bgcell", C.c_int32),
                ("subfacet_points_alias", C.c_int32)]


BUFFER_FIELDS = ["ls_to_bgcell_to_inoutcut", "cutcell_to_cell", "sc_cell_to_points", "sc_cell_to_bgcell",
                 "sc_point_to_coords", "sc_point_to_rcoords", "ls_to_subcell_to_inout", "sf_facet_to_points",
                 "sf_facet_to_bgcell", "sf_facet_to_normal", "sf_point_to_coords", "sf_point_to_rcoords",
                 "ls_to_subfacet_to_inout"]


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BUFFER_FIELDS] + [("state_pitch_bgcell", C.c_int64),
                                                           ("state_pitch_subcell", C.c_int64),
                                                           ("state_pitch_subfacet", C.c_int64)]


class FacetSizes(C.Structure):
    _fields_ = [("num_bgfacets", C.c_int64), ("num_boundary_facets", C.c_int64), ("num_cutfacets", C.c_int64),
                ("num_subfacets", C.c_int64), ("num_subfacet_points", C.c_int64), ("num_levelsets", C.c_int32),
                ("D", C.c_int32), ("num_vertices_per_bgfacet", C.c_int32), ("pad", C.c_int32)]


FACET_BUFFER_FIELDS = ["ls_to_facet_to_inoutcut", "cutfacet_to_facet", "facet_to_nodes", "facet_is_boundary",
                       "sf_cell_to_points", "sf_cell_to_bgcell", "sf_point_to_coords", "sf_point_to_rcoords",
                       "ls_to_subfacet_to_inout"]


class FacetBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in FACET_BUFFER_FIELDS]


FACETS_BOUNDARY, FACETS_SKELETON = 0, 1

SOURCES = ["gecut_api.cu", "gecut_classify.cu", "gecut_cutcells.cu", "gecut_select.cu", "gecut_integrate.cu",
           "gecut_facets.cu", "gecut_aggregate.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-ffp-contract=off"]


def _source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(INCLUDE, "gecut.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


HASH_PATH = os.path.join(LIB_DIR, "BUILD_HASH")


def _needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    return open(HASH_PATH).read().strip() != _source_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (`--fmad=false`: the reference's Julia arithmetic never fuses
    a*b+c and IN/OUT decisions of nodes lying on a surface depend on it) and link libgecut.so in-tree."""
    if not force and not _needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    objdir = os.path.join(LIB_DIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            failed.append(src)
    if failed:
        raise GecutError(f"nvcc failed for {failed}")
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    with open(HASH_PATH, "w") as fh:
        fh.write(_source_hash())
    return LIB_PATH


_lib = None


def lib():
    """Load libgecut.so (building it first when the sources are newer).  Raises GecutError when it cannot be
    loaded -- the product path never degrades to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        build_library()
    except FileNotFoundError as e:  # no nvcc on this machine: use the prebuilt library if there is one
        if not os.path.exists(LIB_PATH):
            raise GecutError(f"libgecut.so is missing and nvcc is unavailable: {e}")
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:
        raise GecutError(f"cannot load {LIB_PATH}: {e}")
    vp, i32, i64, cint = C.c_void_p, C.c_int32, C.c_int64, C.c_int
    P = C.POINTER
    L.gecut_create.argtypes = [cint, vp, P(vp)]
    L.gecut_destroy.argtypes = [vp]
    L.gecut_set_stream.argtypes = [vp, vp]
    L.gecut_last_error.argtypes = [vp]
    L.gecut_last_error.restype = C.c_char_p
    L.gecut_launch_count.argtypes = [vp]
    L.gecut_launch_count.restype = i64
    L.gecut_trim.argtypes = [vp]
    L.gecut_trim.restype = cint
    L.gecut_profile.argtypes = [vp, cint]
    L.gecut_profile.restype = cint
    L.gecut_profile_report.argtypes = [vp]
    L.gecut_profile_report.restype = C.c_char_p
    L.gecut_cut.argtypes = [vp, P(Grid), P(LeafRec), cint, P(vp), cint, P(vp)]
    L.gecut_classify.argtypes = [vp, P(Grid), P(LeafRec), cint, P(vp), cint, P(vp)]
    L.gecut_eval_leaf_at_nodes.argtypes = [vp, P(Grid), P(LeafRec), vp, cint]
    L.gecut_result_free.argtypes = [vp]
    L.gecut_result_sizes.argtypes = [vp, P(Sizes)]
    L.gecut_result_device_ptrs.argtypes = [vp, P(Buffers)]
    L.gecut_result_copy.argtypes = [vp, P(Buffers)]
    L.gecut_bgcell_to_inoutcut.argtypes = [vp, vp, cint, vp]
    L.gecut_subcell_to_inout.argtypes = [vp, vp, cint, vp]
    L.gecut_select_cells.argtypes = [vp, cint, vp, cint, cint, P(vp)]
    L.gecut_select_boundary.argtypes = [vp, vp, cint, vp, cint, P(vp)]
    L.gecut_selection_free.argtypes = [vp]
    L.gecut_selection_sizes.argtypes = [vp, P(i64), P(i64)]
    L.gecut_selection_copy.argtypes = [vp, vp, vp, vp]
    L.gecut_selection_gather.argtypes = [vp, vp, vp, vp]
    L.gecut_integrate_measure.argtypes = [vp, vp, P(C.c_double)]
    L.gecut_integrate_laplacian.argtypes = [vp, vp, vp]
    L.gecut_selection_device_laplacian.argtypes = [vp, P(vp)]
    L.gecut_cell_measure.argtypes = [vp, vp, C.c_int]
    L.gecut_cut_facets.argtypes = [vp, P(Grid), P(LeafRec), cint, P(vp), cint, cint, P(vp)]
    L.gecut_facet_result_free.argtypes = [vp]
    L.gecut_facet_result_sizes.argtypes = [vp, P(FacetSizes)]
    L.gecut_facet_result_copy.argtypes = [vp, P(FacetBuffers)]
    L.gecut_bgfacet_to_inoutcut.argtypes = [vp, vp, cint, vp]
    L.gecut_subfacet_to_inout.argtypes = [vp, vp, cint, vp]
    L.gecut_select_facets.argtypes = [vp, cint, cint, vp, cint, cint, P(vp)]
    L.gecut_facet_selection_free.argtypes = [vp]
    L.gecut_facet_selection_sizes.argtypes = [vp, P(i64), P(i64)]
    L.gecut_facet_selection_copy.argtypes = [vp, vp, vp]
    L.gecut_facet_selection_measure.argtypes = [vp, vp, P(C.c_double)]
    L.gecut_aggregate.argtypes = [vp, vp, vp, cint, cint, C.c_double, vp, cint]
    L.gecut_ghost_skeleton.argtypes = [vp, vp, cint, cint, P(i64), vp, vp]
    for name in FACET_SYMBOLS:
        getattr(L, name).restype = cint
    for name in ("gecut_create", "gecut_destroy", "gecut_set_stream", "gecut_cut", "gecut_classify",
                 "gecut_eval_leaf_at_nodes", "gecut_result_free", "gecut_result_sizes", "gecut_result_device_ptrs",
                 "gecut_result_copy", "gecut_bgcell_to_inoutcut", "gecut_subcell_to_inout", "gecut_select_cells",
                 "gecut_select_boundary", "gecut_selection_free", "gecut_selection_sizes", "gecut_selection_copy",
                 "gecut_selection_gather", "gecut_integrate_measure", "gecut_integrate_laplacian",
                 "gecut_selection_device_laplacian", "gecut_cell_measure"):
        getattr(L, name).restype = cint
    _lib = L
    return L


FACET_SYMBOLS = ["gecut_cut_facets", "gecut_facet_result_free", "gecut_facet_result_sizes", "gecut_facet_result_copy",
                 "gecut_bgfacet_to_inoutcut", "gecut_subfacet_to_inout", "gecut_select_facets",
                 "gecut_facet_selection_free", "gecut_facet_selection_sizes", "gecut_facet_selection_copy",
                 "gecut_facet_selection_measure", "gecut_aggregate", "gecut_ghost_skeleton"]

EXPORTED_SYMBOLS = FACET_SYMBOLS + ["gecut_create", "gecut_destroy", "gecut_set_stream", "gecut_last_error", "gecut_launch_count",
                    "gecut_profile", "gecut_profile_report", "gecut_trim",
                    "gecut_cut", "gecut_classify", "gecut_eval_leaf_at_nodes", "gecut_result_free",
                    "gecut_result_sizes", "gecut_result_device_ptrs", "gecut_result_copy", "gecut_bgcell_to_inoutcut",
                    "gecut_subcell_to_inout", "gecut_select_cells", "gecut_select_boundary", "gecut_selection_free",
                    "gecut_selection_sizes", "gecut_selection_copy", "gecut_selection_gather",
                    "gecut_integrate_measure", "gecut_integrate_laplacian", "gecut_selection_device_laplacian",
                    "gecut_cell_measure"]


class Context:
    """`gecut_ctx`: one per GPU / host thread."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        self._lib = lib()
        rc = self._lib.gecut_create(int(device), C.c_void_p(stream or 0), C.byref(self._h))
        if rc != GECUT_OK:
            raise GecutError(f"gecut_create failed with status {rc} (no CUDA device? there is no CPU fallback)")
        self.device = device

    def check(self, rc: int, what: str = ""):
        if rc != GECUT_OK:
            msg = self._lib.gecut_last_error(self._h)
            raise GecutError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")

    def last_error(self) -> str:
        msg = self._lib.gecut_last_error(self._h)
        return msg.decode() if msg else ""

    def set_stream(self, stream: int):
        self.check(self._lib.gecut_set_stream(self._h, C.c_void_p(stream or 0)), "gecut_set_stream")

    @property
    def launch_count(self) -> int:
        return int(self._lib.gecut_launch_count(self._h))

    @property
    def handle(self):
        return self._h

    def trim(self):
        """Return the context's cached device blocks to the driver."""
        self.check(self._lib.gecut_trim(self._h), "gecut_trim")

    def profile(self, enable: bool):
        self.check(self._lib.gecut_profile(self._h, 1 if enable else 0), "gecut_profile")

    def profile_report(self) -> dict:
        """{kernel name: (launches, total ms)} since profile(True)."""
        txt = self._lib.gecut_profile_report(self._h).decode()
        out = {}
        for line in txt.splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    def close(self):
        if self._h:
            self._lib.gecut_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device: int | None = None) -> Context:
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
