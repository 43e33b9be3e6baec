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
GECUT_ERR_INVALID, GECUT_ERR_NOTIMPLEMENTED, GECUT_ERR_CUDA, GECUT_ERR_OOM, GECUT_ERR_OVERFLOW, GECUT_ERR_COMM = 1, 2, 3, 4, 5, 6
COMM_ID_BYTES, COMM_MAX_VALUES = 128, 16
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
           "gecut_facets.cu", "gecut_aggregate.cu", "gecut_comm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-ffp-contract=off"]


def _nvcc() -> str:
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


_nvcc_version_cache = None


def _nvcc_version() -> str:
    global _nvcc_version_cache
    if _nvcc_version_cache is None:
        try:
            out = subprocess.run([_nvcc(), "--version"], capture_output=True, text=True, timeout=30).stdout
            _nvcc_version_cache = out.strip().splitlines()[-1] if out.strip() else "unknown"
        except Exception:
            _nvcc_version_cache = "absent"
    return _nvcc_version_cache


def _source_hash(with_compiler: bool = True) -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(INCLUDE, "gecut.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    if with_compiler:
        h.update(_nvcc_version().encode())
    return h.hexdigest()


HASH_PATH = os.path.join(LIB_DIR, "BUILD_HASH")


def _needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    return open(HASH_PATH).read().strip() != _source_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (`--fmad=false`: the reference's Julia arithmetic never fuses
    a*b+c and IN/OUT decisions of nodes lying on a surface depend on it) and link libgecut.so in-tree.

    Safe under torchrun: the ranks serialise on a file lock, the first one builds into a private directory and moves the
    finished library into place atomically, the others find an up-to-date library when they get the lock."""
    if not force and not _needs_build():
        return LIB_PATH
    import fcntl
    import shutil
    import tempfile
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _needs_build():  # another rank built it while this one waited
                return LIB_PATH
            nvcc = _nvcc()
            tmp = tempfile.mkdtemp(prefix=".build-", dir=LIB_DIR)
            try:
                procs = []
                objs = []
                for src in SOURCES:
                    obj = os.path.join(tmp, src.replace(".cu", ".o"))
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
                so = os.path.join(tmp, "libgecut.so")
                subprocess.check_call([nvcc, "-shared", "-o", so] + objs + ["-lcudart", "-lnccl"])
                os.replace(so, LIB_PATH)
                with open(os.path.join(tmp, "BUILD_HASH"), "w") as fh:
                    fh.write(_source_hash())
                os.replace(os.path.join(tmp, "BUILD_HASH"), HASH_PATH)
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
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
    # libgecut.so needs libnccl.so.2.  PyTorch ships its own copy under the same SONAME: load torch first so that one
    # process never maps two NCCL versions (whichever is mapped first serves both).
    try:
        import torch  # noqa: F401
    except Exception:
        pass
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
    L.gecut_host_alloc.argtypes = [C.c_size_t, P(vp)]
    L.gecut_host_free.argtypes = [vp]
    L.gecut_host_register.argtypes = [vp, C.c_size_t]
    L.gecut_host_unregister.argtypes = [vp]
    for name in ("gecut_host_alloc", "gecut_host_free", "gecut_host_register", "gecut_host_unregister"):
        getattr(L, name).restype = cint
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
    L.gecut_facet_topology_host.argtypes = [P(Grid), i64, i64, vp, vp, P(i64), P(C.c_int32)]
    L.gecut_bgfacet_to_inoutcut.argtypes = [vp, vp, cint, vp]
    L.gecut_subfacet_to_inout.argtypes = [vp, vp, cint, vp]
    L.gecut_select_facets.argtypes = [vp, cint, cint, vp, cint, cint, P(vp)]
    L.gecut_facet_selection_free.argtypes = [vp]
    L.gecut_facet_selection_sizes.argtypes = [vp, P(i64), P(i64)]
    L.gecut_facet_selection_copy.argtypes = [vp, vp, vp]
    L.gecut_facet_selection_measure.argtypes = [vp, vp, P(C.c_double)]
    L.gecut_aggregate.argtypes = [vp, vp, vp, cint, vp, cint, cint, C.c_double, vp, cint]
    L.gecut_aggregate_slabs.argtypes = [P(vp), P(vp), cint, vp, cint, vp, cint, cint, C.c_double, P(vp), cint]
    L.gecut_comm_aggregate.argtypes = [vp, vp, vp, cint, vp, cint, cint, C.c_double, vp, cint]
    L.gecut_ghost_skeleton.argtypes = [vp, vp, cint, cint, P(i64), vp, vp]
    L.gecut_ghost_skeleton_slabs.argtypes = [P(vp), cint, vp, cint, cint, P(i64), P(vp)]
    L.gecut_comm_ghost_skeleton.argtypes = [vp, vp, cint, cint, P(i64), vp]
    L.gecut_comm_unique_id.argtypes = [vp]
    L.gecut_comm_init.argtypes = [vp, vp, cint, cint]
    L.gecut_comm_free.argtypes = [vp]
    L.gecut_comm_rank.argtypes = [vp, P(cint), P(cint)]
    L.gecut_comm_exchange_ghost_planes.argtypes = [vp, P(vp), cint, i64, i64]
    L.gecut_comm_check_consistency.argtypes = [vp, P(vp), cint, i64, i64, P(cint)]
    L.gecut_comm_allreduce_sum.argtypes = [vp, P(C.c_double), cint, vp]
    L.gecut_comm_fetch.argtypes = [vp, vp, cint, P(C.c_double)]
    L.gecut_comm_allgather_sizes.argtypes = [vp, vp, P(Sizes)]
    for name in FACET_SYMBOLS + COMM_SYMBOLS:
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
                 "gecut_facet_topology_host",
                 "gecut_bgfacet_to_inoutcut", "gecut_subfacet_to_inout", "gecut_select_facets",
                 "gecut_facet_selection_free", "gecut_facet_selection_sizes", "gecut_facet_selection_copy",
                 "gecut_facet_selection_measure", "gecut_aggregate", "gecut_aggregate_slabs", "gecut_ghost_skeleton",
                 "gecut_ghost_skeleton_slabs"]

COMM_SYMBOLS = ["gecut_comm_unique_id", "gecut_comm_init", "gecut_comm_free", "gecut_comm_rank",
                "gecut_comm_exchange_ghost_planes", "gecut_comm_check_consistency", "gecut_comm_allreduce_sum",
                "gecut_comm_fetch", "gecut_comm_allgather_sizes", "gecut_comm_aggregate",
                "gecut_comm_ghost_skeleton"]

EXPORTED_SYMBOLS = FACET_SYMBOLS + COMM_SYMBOLS + ["gecut_create", "gecut_destroy", "gecut_set_stream", "gecut_last_error", "gecut_launch_count",
                    "gecut_profile", "gecut_profile_report", "gecut_trim",
                    "gecut_host_alloc", "gecut_host_free", "gecut_host_register", "gecut_host_unregister",
                    "gecut_cut", "gecut_classify", "gecut_eval_leaf_at_nodes", "gecut_result_free",
                    "gecut_result_sizes", "gecut_result_device_ptrs", "gecut_result_copy", "gecut_bgcell_to_inoutcut",
                    "gecut_subcell_to_inout", "gecut_select_cells", "gecut_select_boundary", "gecut_selection_free",
                    "gecut_selection_sizes", "gecut_selection_copy", "gecut_selection_gather",
                    "gecut_integrate_measure", "gecut_integrate_laplacian", "gecut_selection_device_laplacian",
                    "gecut_cell_measure"]


def pinned_empty(shape, dtype):
    """numpy array over page-locked memory from gecut_host_alloc (the zero-copy destination of gecut_result_copy: what the
    Julia shim wraps with unsafe_wrap); freed with the array."""
    import weakref
    import numpy as np
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if len(tuple(shape)) else 1
    ptr = C.c_void_p()
    L = lib()
    rc = L.gecut_host_alloc(C.c_size_t(max(n, 1) * dt.itemsize), C.byref(ptr))
    if rc != GECUT_OK:
        raise GecutError(f"gecut_host_alloc failed with status {rc}")
    buf = (C.c_char * (max(n, 1) * dt.itemsize)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    weakref.finalize(buf, L.gecut_host_free, ptr)
    return arr


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

    # ---- multi-GPU (include/gecut.h "multi-GPU"): NCCL collectives enqueued on this context's stream -------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128 bytes to hand to every rank's `comm_init` (rank 0 creates them; the host program distributes them)."""
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = lib().gecut_comm_unique_id(buf)
        if rc != GECUT_OK:
            raise GecutError(f"gecut_comm_unique_id failed with status {rc}")
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, world: int):
        if len(uid) != COMM_ID_BYTES:
            raise ValueError("the communicator id has GECUT_COMM_ID_BYTES bytes")
        self.check(self._lib.gecut_comm_init(self._h, C.c_char_p(uid), int(rank), int(world)), "gecut_comm_init")

    def comm_free(self):
        self.check(self._lib.gecut_comm_free(self._h), "gecut_comm_free")

    @property
    def comm_rank_world(self):
        r, w = C.c_int(), C.c_int()
        self.check(self._lib.gecut_comm_rank(self._h, C.byref(r), C.byref(w)), "gecut_comm_rank")
        return int(r.value), int(w.value)

    def comm_exchange_ghost_planes(self, slab_values, layer_nodes: int):
        """In-place ghost update of slab-local nodal vectors (CUDA float64 torch tensors of num_planes * layer_nodes values)."""
        vecs = [v for v in slab_values]
        if not vecs:
            return
        n = int(vecs[0].numel())
        if any(int(v.numel()) != n or n % layer_nodes for v in vecs):
            raise ValueError("every vector holds the same whole number of node planes")
        ptrs = (C.c_void_p * len(vecs))(*[v.data_ptr() for v in vecs])
        self.check(self._lib.gecut_comm_exchange_ghost_planes(self._h, ptrs, len(vecs), int(layer_nodes), n // layer_nodes),
                   "gecut_comm_exchange_ghost_planes")

    def comm_check_consistency(self, slab_values=(), layer_nodes: int = 0) -> bool:
        """`isconsistent_bgcell_to_inoutcut` (DistributedDiscretizations.jl:136-174) for the slab partition."""
        vecs = [v for v in slab_values]
        ok = C.c_int(0)
        if vecs:
            n = int(vecs[0].numel())
            ptrs = (C.c_void_p * len(vecs))(*[v.data_ptr() for v in vecs])
            rc = self._lib.gecut_comm_check_consistency(self._h, ptrs, len(vecs), int(layer_nodes), n // layer_nodes,
                                                        C.byref(ok))
        else:
            rc = self._lib.gecut_comm_check_consistency(self._h, None, 0, 0, 0, C.byref(ok))
        self.check(rc, "gecut_comm_check_consistency")
        return bool(ok.value)

    def comm_allreduce_sum(self, values, result_dev_ptr: int):
        """Global sums of a few doubles into device memory at `result_dev_ptr`, stream-ordered (no host wait)."""
        n = len(values)
        arr = (C.c_double * n)(*[float(v) for v in values])
        self.check(self._lib.gecut_comm_allreduce_sum(self._h, arr, n, C.c_void_p(result_dev_ptr)),
                   "gecut_comm_allreduce_sum")

    def comm_fetch(self, dev_ptr: int, n: int):
        out = (C.c_double * n)()
        self.check(self._lib.gecut_comm_fetch(self._h, C.c_void_p(dev_ptr), n, out), "gecut_comm_fetch")
        return [float(x) for x in out]

    def comm_allgather_sizes(self, result_handle):
        _, w = self.comm_rank_world
        arr = (Sizes * w)()
        self.check(self._lib.gecut_comm_allgather_sizes(self._h, result_handle, arr), "gecut_comm_allgather_sizes")
        return list(arr)

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
