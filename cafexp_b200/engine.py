"""ctypes binding of the C ABI in include/cafe_b200.h (cafexp_b200/libcafe_b200.so).

This is the thinnest possible layer: numpy arrays in, numpy arrays out, every call goes to the CUDA
library.  There is no CPU fallback — if the library has not been built (``__graft_entry__.build()``)
or no CUDA device is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcafe_b200.so")

BASE_LOGMAX = 0
GAMMA_LINSUM = 1
OPT_RESCALE = 1
OPT_MAX_SLOTS = 2

ERR_NAMES = {-1: "ERR_ARG", -2: "ERR_CUDA", -3: "ERR_COUNT_RANGE", -4: "ERR_LIMIT"}

#: every symbol include/cafe_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "cafe_b200_abi_version", "cafe_b200_get_limits", "cafe_b200_device_count", "cafe_b200_create", "cafe_b200_destroy",
    "cafe_b200_last_error", "cafe_b200_set_families", "cafe_b200_set_error_model", "cafe_b200_set_option", "cafe_b200_set_stream",
    "cafe_b200_eval", "cafe_b200_eval_device", "cafe_b200_reconstruct", "cafe_b200_build_matrices", "cafe_b200_matrix_size",
    "cafe_b200_prune_roots", "cafe_b200_launch_count", "cafe_b200_last_timings", "cafe_b200_root_max", "cafe_b200_pvalues",
    "cafe_b200_branch_probabilities", "cafe_b200_create_multi", "cafe_b200_n_devices", "cafe_b200_alloc_pinned", "cafe_b200_free_pinned",
    "cafe_b200_set_families_ex", "cafe_b200_plan_program", "cafe_b200_describe", "cafe_b200_fetch_category_likelihoods", "cafe_b200_timing_history", "cafe_b200_host_seconds", "cafe_b200_set_build_partition", "cafe_b200_read_family_table",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class CafeB200Error(RuntimeError):
    pass


#: int gather(void* user, void* matrices, size_t slab_bytes, int n_parts, void* cuda_stream)
GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)


class _Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("parent", _ip), ("child_offset", _ip), ("child_list", _ip),
                ("leaf_col", _ip), ("branch", _dp), ("lambda_index", _ip)]


class _Limits(C.Structure):
    _fields_ = [("max_matrix_size", C.c_int), ("max_categories", C.c_int), ("max_nodes", C.c_int), ("families_per_tile", C.c_int)]


_lib = None


def load_library():
    """Load libcafe_b200.so; raises if it was not built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CafeB200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). cafexp_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.cafe_b200_abi_version.restype = C.c_int
    L.cafe_b200_get_limits.argtypes = [C.POINTER(_Limits)]
    L.cafe_b200_device_count.restype = C.c_int
    L.cafe_b200_create.restype = C.c_int
    L.cafe_b200_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Tree), _i32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
    L.cafe_b200_create_multi.restype = C.c_int
    L.cafe_b200_create_multi.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Tree), C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                         _ip, C.c_int]
    L.cafe_b200_n_devices.restype = C.c_int
    L.cafe_b200_n_devices.argtypes = [C.c_void_p]
    L.cafe_b200_alloc_pinned.restype = C.c_void_p
    L.cafe_b200_alloc_pinned.argtypes = [C.c_size_t]
    L.cafe_b200_free_pinned.argtypes = [C.c_void_p]
    L.cafe_b200_set_families_ex.restype = C.c_int
    L.cafe_b200_set_families_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64]
    L.cafe_b200_plan_program.restype = C.c_int
    L.cafe_b200_plan_program.argtypes = [C.POINTER(_Tree), _ip, C.c_int, _ip, _ip, C.c_int, _ip, _ip]
    L.cafe_b200_fetch_category_likelihoods.restype = C.c_int
    L.cafe_b200_fetch_category_likelihoods.argtypes = [C.c_void_p, C.c_int, _dp]
    L.cafe_b200_timing_history.restype = C.c_int
    L.cafe_b200_timing_history.argtypes = [C.c_void_p, C.c_int, _dp]
    L.cafe_b200_host_seconds.restype = C.c_int
    L.cafe_b200_host_seconds.argtypes = [C.c_void_p, _dp]
    L.cafe_b200_set_build_partition.restype = C.c_int
    L.cafe_b200_set_build_partition.argtypes = [C.c_void_p, C.c_int, C.c_int, GATHER_FN, C.c_void_p]
    L.cafe_b200_read_family_table.restype = C.c_int
    L.cafe_b200_read_family_table.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, _i32p, C.c_int64, _i64p, C.c_char_p, C.c_int]
    L.cafe_b200_describe.restype = C.c_int
    L.cafe_b200_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.cafe_b200_destroy.argtypes = [C.c_void_p]
    L.cafe_b200_last_error.restype = C.c_char_p
    L.cafe_b200_last_error.argtypes = [C.c_void_p]
    L.cafe_b200_set_families.restype = C.c_int
    L.cafe_b200_set_families.argtypes = [C.c_void_p, _i32p, C.c_int64]
    L.cafe_b200_set_error_model.restype = C.c_int
    L.cafe_b200_set_error_model.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int]
    L.cafe_b200_set_option.restype = C.c_int
    L.cafe_b200_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.cafe_b200_set_stream.restype = C.c_int
    L.cafe_b200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.cafe_b200_eval.restype = C.c_int
    L.cafe_b200_eval.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _i64p, _i64p, C.c_int64]
    L.cafe_b200_eval_device.restype = C.c_int
    L.cafe_b200_eval_device.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_void_p]
    L.cafe_b200_reconstruct.restype = C.c_int
    L.cafe_b200_reconstruct.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp, _i32p]
    L.cafe_b200_build_matrices.restype = C.c_int
    L.cafe_b200_build_matrices.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp]
    L.cafe_b200_matrix_size.restype = C.c_int
    L.cafe_b200_matrix_size.argtypes = [C.c_void_p]
    L.cafe_b200_prune_roots.restype = C.c_int
    L.cafe_b200_prune_roots.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp]
    L.cafe_b200_root_max.restype = C.c_int
    L.cafe_b200_root_max.argtypes = [C.c_void_p, _dp, C.c_int, _dp]
    L.cafe_b200_pvalues.restype = C.c_int
    L.cafe_b200_pvalues.argtypes = [C.c_int, _dp, C.c_int, C.c_int, _dp, C.c_int64, _dp]
    L.cafe_b200_branch_probabilities.restype = C.c_int
    L.cafe_b200_branch_probabilities.argtypes = [C.c_void_p, _dp, C.c_int, _i32p, C.POINTER(C.c_uint8), _dp]
    L.cafe_b200_launch_count.restype = C.c_int64
    L.cafe_b200_launch_count.argtypes = [C.c_void_p]
    L.cafe_b200_last_timings.restype = C.c_int
    L.cafe_b200_last_timings.argtypes = [C.c_void_p, _dp]
    if L.cafe_b200_abi_version() != 2:
        raise CafeB200Error("libcafe_b200.so ABI version mismatch")
    _lib = L
    return L


def device_count() -> int:
    return load_library().cafe_b200_device_count()


def limits() -> dict:
    lim = _Limits()
    load_library().cafe_b200_get_limits(C.byref(lim))
    return {name: getattr(lim, name) for name, _ in _Limits._fields_}


def _d(a):
    return a.ctypes.data_as(_dp)


def tree_struct(tree):
    """(ctypes struct, arrays it points into) for a flattened tree."""
    arrays = [np.ascontiguousarray(tree.parent, np.int32), np.ascontiguousarray(tree.child_offset, np.int32),
              np.ascontiguousarray(tree.child_list, np.int32), np.ascontiguousarray(tree.leaf_col, np.int32),
              np.ascontiguousarray(tree.branch, np.float64), np.ascontiguousarray(tree.lambda_index, np.int32)]
    a = arrays
    ts = _Tree(len(a[0]), a[0].ctypes.data_as(_ip), a[1].ctypes.data_as(_ip), a[2].ctypes.data_as(_ip), a[3].ctypes.data_as(_ip),
               _d(a[4]), a[5].ctypes.data_as(_ip))
    return ts, arrays


def _count_array(counts: np.ndarray) -> np.ndarray:
    """Counts as the ABI takes them: uint8 / uint16 stay as they are (a quarter / half of the bytes on the wire),
    everything else becomes int32."""
    counts = np.asarray(counts)
    if counts.dtype in (np.uint8, np.uint16, np.int32):
        return np.ascontiguousarray(counts)
    return np.ascontiguousarray(counts, np.int32)


class Engine:
    """One context = one (tree, families, size limits) on one GPU or — ``device`` a list — sharded over several GPUs of
    the box; mirrors what a reference ``model`` object holds between evaluations (src/core.h:122-186)."""

    def __init__(self, tree, counts: np.ndarray, max_family_size: int, max_root_family_size: int, device=0):
        L = load_library()
        self._lib = L
        self.tree = tree
        ts, self._arrays = tree_struct(tree)
        counts = _count_array(counts)
        if counts.ndim != 2 or counts.shape[1] != tree.n_leaves:
            raise ValueError("counts must be [n_families, n_leaves]")
        self.n_families = int(counts.shape[0])
        self.n_leaves = int(counts.shape[1])
        self.n_internal = int((np.asarray(tree.leaf_col) < 0).sum())
        self.mf = int(max_family_size)
        self.mrf = int(max_root_family_size)
        self.devices = [int(d) for d in (device if isinstance(device, (list, tuple)) else [device])]
        self.device = self.devices[0]
        devs = np.ascontiguousarray(self.devices, np.int32)
        handle = C.c_void_p()
        rc = L.cafe_b200_create_multi(C.byref(handle), C.byref(ts), counts.ctypes.data_as(C.c_void_p), counts.dtype.itemsize, self.n_families,
                                      self.n_leaves, self.mf, self.mrf, devs.ctypes.data_as(_ip), len(self.devices))
        if rc != 0:
            raise CafeB200Error(f"cafe_b200_create: {ERR_NAMES.get(rc, rc)}: {L.cafe_b200_last_error(None).decode()}")
        self._h = handle
        self.matrix_size = L.cafe_b200_matrix_size(self._h)

    def describe(self) -> str:
        buf = C.create_string_buffer(1024)
        self._lib.cafe_b200_describe(self._h, buf, 1024)
        return buf.value.decode()

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.cafe_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise CafeB200Error(f"{what}: {ERR_NAMES.get(rc, rc)}: {self._lib.cafe_b200_last_error(self._h).decode()}")

    # -- configuration --------------------------------------------------------------------------
    def set_families(self, counts: np.ndarray):
        counts = _count_array(counts)
        self._check(self._lib.cafe_b200_set_families_ex(self._h, counts.ctypes.data_as(C.c_void_p), counts.dtype.itemsize, counts.shape[0]),
                    "set_families")

    def set_error_model(self, table: Optional[np.ndarray]):
        if table is None:
            self._check(self._lib.cafe_b200_set_error_model(self._h, None, 0, 0), "set_error_model")
            return
        t = np.ascontiguousarray(table, np.float64)
        self._check(self._lib.cafe_b200_set_error_model(self._h, _d(t), t.shape[0], t.shape[1]), "set_error_model")

    def set_rescale(self, on: bool):
        self._check(self._lib.cafe_b200_set_option(self._h, OPT_RESCALE, 1 if on else 0), "set_option")

    def set_max_slots(self, n: int):
        """Cap the on-chip vector storage (>= 2): fewer slots force both tree-walking kernels to spill (tests)."""
        self._check(self._lib.cafe_b200_set_option(self._h, OPT_MAX_SLOTS, int(n)), "set_option")

    def set_build_partition(self, part: int, n_parts: int, gather=None):
        """Distributed matrix build across processes: this context builds slab ``part`` of ``n_parts`` and calls
        ``gather(matrices_ptr, slab_bytes, n_parts, cuda_stream_ptr)`` to all-gather the slabs in place
        (see cafexp_b200.sharded.nccl_matrix_gather).  ``n_parts = 1`` restores the replicated build."""
        if gather is None:
            self._gather_cb = GATHER_FN()
        else:
            def trampoline(_user, matrices, slab_bytes, parts, stream):
                try:
                    gather(matrices, slab_bytes, parts, stream or 0)
                    return 0
                except Exception as e:      # noqa: BLE001 — must not propagate through the C frame
                    import sys
                    sys.stderr.write(f"cafexp_b200: matrix all-gather failed: {e!r}\n")
                    return 1
            self._gather_cb = GATHER_FN(trampoline)          # kept alive as long as the context may call it
        self._check(self._lib.cafe_b200_set_build_partition(self._h, int(part), int(n_parts), self._gather_cb, None), "set_build_partition")

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._lib.cafe_b200_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    # -- the hot path ---------------------------------------------------------------------------
    @staticmethod
    def _lams(lambdas):
        lam = np.ascontiguousarray(np.atleast_2d(np.asarray(lambdas, np.float64)))
        return lam, lam.shape[0], lam.shape[1]

    def infer(self, lambdas, prior, cat_probs=None, mode=BASE_LOGMAX, want_family=True, want_cat=True, failed_cap=1024,
              out_family=None, out_cat=None):
        """One likelihood evaluation.  lambdas: [k][n_lambdas] raw lambda_i*multiplier_k.
        out_family / out_cat: caller-owned float64 arrays (e.g. page-locked) the per-family outputs are written to."""
        lam, k, nl = self._lams(lambdas)
        cp = np.ascontiguousarray(cat_probs if cat_probs is not None else np.ones(k), np.float64)
        pr = np.ascontiguousarray(prior, np.float64)
        if pr.shape[0] < self.mrf:
            raise ValueError("prior must have max_root_family_size entries")
        score = C.c_double()
        nf = C.c_int64()
        fam = (out_family if out_family is not None else np.empty(self.n_families)) if want_family else None
        cat = (out_cat if out_cat is not None else np.empty((self.n_families, k))) if (want_cat and mode == GAMMA_LINSUM) else None
        if fam is not None and (fam.dtype != np.float64 or fam.size < self.n_families or not fam.flags.c_contiguous):
            raise ValueError("out_family must be a contiguous float64 array of n_families entries")
        if cat is not None and (cat.dtype != np.float64 or cat.size < self.n_families * k or not cat.flags.c_contiguous):
            raise ValueError("out_cat must be a contiguous float64 array of n_families x k entries")
        fidx = np.full(failed_cap, -1, np.int64)
        rc = self._lib.cafe_b200_eval(self._h, _d(lam), nl, _d(cp), k, _d(pr), mode, C.byref(score),
                                      _d(fam) if fam is not None else None, _d(cat) if cat is not None else None,
                                      C.byref(nf), fidx.ctypes.data_as(_i64p), failed_cap)
        self._check(rc, "cafe_b200_eval")
        return {"score": score.value, "family_lnl": fam, "cat_lk": cat, "n_failed": nf.value, "failed_idx": fidx[fidx >= 0]}

    def fetch_category_likelihoods(self, k: int) -> np.ndarray:
        """[n_families][k] category likelihoods of the last gamma evaluation (for callers that skipped them in infer)."""
        out = np.empty((self.n_families, k))
        self._check(self._lib.cafe_b200_fetch_category_likelihoods(self._h, k, _d(out)), "cafe_b200_fetch_category_likelihoods")
        return out

    def infer_device(self, lambdas, prior, cat_probs, mode, result_ptr: int):
        """Asynchronous evaluation; [sum lnL, n_failed] are written to device memory at result_ptr."""
        lam, k, nl = self._lams(lambdas)
        cp = np.ascontiguousarray(cat_probs if cat_probs is not None else np.ones(k), np.float64)
        pr = np.ascontiguousarray(prior, np.float64)
        if pr.shape[0] < self.mrf:
            raise ValueError("prior must have max_root_family_size entries")
        if cp.shape[0] != k:
            raise ValueError("one category probability per lambda row")
        self._check(self._lib.cafe_b200_eval_device(self._h, _d(lam), nl, _d(cp), k, _d(pr), mode, C.c_void_p(result_ptr)), "cafe_b200_eval_device")

    def reconstruct(self, lambdas, prior_by_size):
        lam, k, nl = self._lams(lambdas)
        pr = np.ascontiguousarray(prior_by_size, np.float64)
        if pr.shape[0] < min(self.mf, self.mrf) + 1:
            raise ValueError("prior_by_size must have min(mf, mrf)+1 entries")
        states = np.empty((self.n_families, k, self.n_internal), np.int32)
        self._check(self._lib.cafe_b200_reconstruct(self._h, _d(lam), nl, k, _d(pr), states.ctypes.data_as(_i32p)), "cafe_b200_reconstruct")
        return states

    # -- inspection -----------------------------------------------------------------------------
    def root_max(self, lambdas):
        """max_j of every family's root vector under one lambda set (cafe_b200_root_max): the likelihood
        compute_pvalues uses for simulated and observed families alike (src/probability.cpp:308, 399)."""
        lam = np.ascontiguousarray(lambdas, np.float64).ravel()
        out = np.zeros(self.n_families)
        self._check(self._lib.cafe_b200_root_max(self._h, _d(lam), lam.size, _d(out)), "cafe_b200_root_max")
        return out

    def branch_probabilities(self, lambdas, node_sizes, selected=None):
        """compute_viterbi_sum for every (family, node) (cafe_b200_branch_probabilities); -1 where the reference has no value."""
        lam = np.ascontiguousarray(lambdas, np.float64).ravel()
        sizes = np.ascontiguousarray(node_sizes, np.int32)
        assert sizes.shape == (self.n_families, self.tree.n_nodes)
        sel = None if selected is None else np.ascontiguousarray(selected, np.uint8)
        out = np.empty((self.n_families, self.tree.n_nodes))
        self._check(self._lib.cafe_b200_branch_probabilities(self._h, _d(lam), lam.size, sizes.ctypes.data_as(_i32p),
                                                             None if sel is None else sel.ctypes.data_as(C.POINTER(C.c_uint8)), _d(out)),
                    "cafe_b200_branch_probabilities")
        return out

    def build_matrices(self, lambdas):
        lam, k, nl = self._lams(lambdas)
        out = np.empty((k, self.tree.n_nodes, self.matrix_size, self.mf + 1))
        self._check(self._lib.cafe_b200_build_matrices(self._h, _d(lam), nl, k, _d(out)), "cafe_b200_build_matrices")
        return out

    def prune_roots(self, lambdas):
        lam, k, nl = self._lams(lambdas)
        out = np.empty((self.n_families, k, self.mrf))
        self._check(self._lib.cafe_b200_prune_roots(self._h, _d(lam), nl, k, _d(out)), "cafe_b200_prune_roots")
        return out

    @property
    def launches(self) -> int:
        return int(self._lib.cafe_b200_launch_count(self._h))

    def timing_history_ms(self, n: int) -> np.ndarray:
        """[m][4] device times (matrix build, prune, reduce, reconstruct) of the most recent m <= n calls, newest first."""
        ms = np.zeros((n, 4))
        m = self._lib.cafe_b200_timing_history(self._h, n, _d(ms))
        return ms[:max(m, 0)]

    def host_seconds(self) -> dict:
        """Host wall time spent inside the library since create."""
        s3 = np.zeros(3)
        self._lib.cafe_b200_host_seconds(self._h, _d(s3))
        return {"staging": s3[0], "enqueue": s3[1], "wait": s3[2]}

    def last_timings_ms(self) -> dict:
        ms = np.zeros(4)
        self._lib.cafe_b200_last_timings(self._h, _d(ms))
        return {"matrix_build": ms[0], "prune": ms[1], "reduce": ms[2], "reconstruct": ms[3]}


def read_family_table(path: str, leaf_names, id_stride: int = 64):
    """(ids, counts [F][n_leaves] int32) of a CAFE tab-format family table through the library's flat reader
    (cafe_b200_read_family_table; host only — the same result as hostio.read_gene_families, without Python loops)."""
    L = load_library()
    names = (C.c_char_p * len(leaf_names))(*[n.encode() for n in leaf_names])
    n = C.c_int64()
    rc = L.cafe_b200_read_family_table(path.encode(), names, len(leaf_names), None, 0, C.byref(n), None, 0)
    if rc:
        raise CafeB200Error(f"cafe_b200_read_family_table: {L.cafe_b200_last_error(None).decode()}")
    counts = np.zeros((n.value, len(leaf_names)), np.int32)
    ids = C.create_string_buffer(n.value * id_stride)
    rc = L.cafe_b200_read_family_table(path.encode(), names, len(leaf_names), counts.ctypes.data_as(_i32p), n.value, C.byref(n), ids, id_stride)
    if rc:
        raise CafeB200Error(f"cafe_b200_read_family_table: {L.cafe_b200_last_error(None).decode()}")
    raw = ids.raw
    return [raw[i * id_stride:(i + 1) * id_stride].split(b"\0", 1)[0].decode() for i in range(n.value)], counts


def pvalues(cond, observed, device: int = 0) -> np.ndarray:
    """p-value of every observed likelihood against the simulated conditional distributions
    (cafe_b200_pvalues; pvalue / compute_tree_pvalue, src/probability.cpp:379-409).
    cond: [n_root_sizes][n_sim] unsorted; observed: [F]."""
    L = load_library()
    cond = np.ascontiguousarray(cond, np.float64)
    obs = np.ascontiguousarray(observed, np.float64)
    out = np.zeros(len(obs))
    rc = L.cafe_b200_pvalues(device, _d(cond), cond.shape[0], cond.shape[1], _d(obs), len(obs), _d(out))
    if rc:
        raise RuntimeError(f"cafe_b200_pvalues failed ({rc}): {L.cafe_b200_last_error(None).decode()}")
    return out


def neg_inf_safe(x: float) -> float:
    return math.inf if math.isnan(x) else x
