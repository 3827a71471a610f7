"""TEST INFRASTRUCTURE — ctypes binding of oracle/liboracle.so (the CPU restatement) and a runner for
oracle/_ref/ref_harness (the compiled, unmodified reference).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and --impl reference legs),
scripts/make_golden.py.  The product package cafexp_b200 never imports this module.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from typing import Dict, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_HARNESS = os.path.join(HERE, "_ref", "ref_harness")

BASE_LOGMAX = 0
GAMMA_LINSUM = 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i32p = C.POINTER(C.c_int32)


class _Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int), ("parent", _ip), ("child_offset", _ip), ("child_list", _ip),
                ("leaf_col", _ip), ("branch", _dp), ("lambda_index", _ip)]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "cafe_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_chooseln.restype = C.c_double
        L.orc_chooseln.argtypes = [C.c_double, C.c_double]
        L.orc_birthdeath_rate_with_log_alpha.restype = C.c_double
        L.orc_birthdeath_rate_with_log_alpha.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_bd_probability.restype = C.c_double
        L.orc_bd_probability.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_quantise_lambda.restype = C.c_double
        L.orc_quantise_lambda.argtypes = [C.c_double]
        L.orc_quantise_branch.restype = C.c_double
        L.orc_quantise_branch.argtypes = [C.c_double]
        L.orc_is_saturated.restype = C.c_int
        L.orc_is_saturated.argtypes = [C.c_double, C.c_double]
        L.orc_build_matrix.argtypes = [C.c_int, C.c_double, C.c_double, _dp]
        L.orc_incomplete_gamma.restype = C.c_double
        L.orc_incomplete_gamma.argtypes = [C.c_double] * 3
        L.orc_point_normal.restype = C.c_double
        L.orc_point_normal.argtypes = [C.c_double]
        L.orc_point_chi2.restype = C.c_double
        L.orc_point_chi2.argtypes = [C.c_double, C.c_double]
        L.orc_get_gamma.argtypes = [C.c_int, C.c_double, _dp, _dp]
        L.orc_prior_uniform.argtypes = [_ip, _ip, C.c_int, C.c_int, _dp, C.c_int]
        L.orc_prior_poisson.argtypes = [C.c_double, _ip, _ip, C.c_int, C.c_int, _dp, C.c_int]
        L.orc_error_model_replace_epsilon.restype = C.c_int
        L.orc_error_model_replace_epsilon.argtypes = [_dp, C.c_int, C.c_double, C.c_double]
        L.orc_inference_prune.restype = C.c_int
        L.orc_inference_prune.argtypes = [C.POINTER(_Tree), _i32p, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_infer_family_likelihoods.restype = C.c_double
        L.orc_infer_family_likelihoods.argtypes = [C.POINTER(_Tree), _i32p, C.c_int64, C.c_int, _dp, C.c_int, _dp, C.c_int,
                                                   _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   _dp, _dp, C.POINTER(C.c_uint8), C.POINTER(C.c_int64)]
        L.orc_reconstruct.restype = C.c_int
        L.orc_reconstruct.argtypes = [C.POINTER(_Tree), _i32p, C.c_int64, C.c_int, _dp, C.c_int, C.c_int, _dp,
                                      C.c_int, C.c_int, _i32p]
        L.orc_weighted_averages.argtypes = [_i32p, C.c_int, C.c_int, _dp, _dp]
        L.orc_root_max.restype = C.c_int
        L.orc_root_max.argtypes = [C.POINTER(_Tree), _i32p, C.c_int64, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_pvalue.restype = C.c_double
        L.orc_pvalue.argtypes = [C.c_double, _dp, C.c_int]
        L.orc_pvalues.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_int64, _dp]
        L.orc_branch_probabilities.argtypes = [C.POINTER(_Tree), _i32p, C.POINTER(C.c_uint8), C.c_int64, _dp, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_init()
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class TreeHandle:
    """Keeps the numpy arrays alive next to the C struct."""

    def __init__(self, flat):
        self.arrays = [np.ascontiguousarray(flat.parent, np.int32), np.ascontiguousarray(flat.child_offset, np.int32),
                       np.ascontiguousarray(flat.child_list, np.int32), np.ascontiguousarray(flat.leaf_col, np.int32),
                       np.ascontiguousarray(flat.branch, np.float64), np.ascontiguousarray(flat.lambda_index, np.int32)]
        a = self.arrays
        self.struct = _Tree(len(a[0]), _i(a[0]), _i(a[1]), _i(a[2]), _i(a[3]), _d(a[4]), _i(a[5]))
        self.n_internal = int((flat.leaf_col < 0).sum())

    def ref(self):
        return C.byref(self.struct)


def bd_probability(lam, t, s, c):
    return lib().orc_bd_probability(lam, t, s, c)


def build_matrix(n, lam, t):
    out = np.zeros((n, n))
    lib().orc_build_matrix(n, lam, t, _d(out))
    return out


def get_gamma(k, alpha):
    freq = np.zeros(k)
    rate = np.zeros(k)
    lib().orc_get_gamma(k, alpha, _d(freq), _d(rate))
    return freq, rate


def _rootdist_arrays(rootdist: Optional[Dict[int, int]]):
    if not rootdist:
        z = np.zeros(1, np.int32)
        return z, z, 0
    keys = np.asarray(sorted(rootdist), np.int32)
    vals = np.asarray([rootdist[k] for k in sorted(rootdist)], np.int32)
    return keys, vals, len(keys)


def prior_uniform(mrf, rootdist=None, n_out=None):
    keys, vals, n = _rootdist_arrays(rootdist)
    n_out = n_out or mrf
    out = np.zeros(n_out)
    lib().orc_prior_uniform(_i(keys), _i(vals), n, mrf, _d(out), n_out)
    return out


def prior_poisson(poisson_lambda, mrf, rootdist=None, n_out=None):
    keys, vals, n = _rootdist_arrays(rootdist)
    n_out = n_out or mrf
    out = np.zeros(n_out)
    lib().orc_prior_poisson(poisson_lambda, _i(keys), _i(vals), n, mrf, _d(out), n_out)
    return out


def _err_args(err):
    if err is None:
        return None, 0, 0, None
    e = np.ascontiguousarray(err, np.float64)
    return _d(e), e.shape[0], e.shape[1], e


def inference_prune(flat, counts_row, lambdas, mf, mrf, err=None):
    th = TreeHandle(flat)
    row = np.ascontiguousarray(counts_row, np.int32)
    lam = np.ascontiguousarray(lambdas, np.float64)
    out = np.zeros(mrf)
    ep, er, en, keep = _err_args(err)
    rc = lib().orc_inference_prune(th.ref(), row.ctypes.data_as(_i32p), _d(lam), lam.size, ep, er, en, mf, mrf, _d(out))
    if rc:
        raise ValueError("count outside 0..max_family_size")
    return out


def infer(flat, counts, lambdas, cat_probs, prior, mf, mrf, mode, err=None):
    """lambdas: [k][n_lambdas] raw (lambda_i * multiplier_k).  Returns dict(score, family_lnl, cat_lk, failed)."""
    th = TreeHandle(flat)
    counts = np.ascontiguousarray(counts, np.int32)
    F, nl = counts.shape
    lam = np.ascontiguousarray(np.atleast_2d(lambdas), np.float64)
    k, n_lambdas = lam.shape
    cp = np.ascontiguousarray(cat_probs, np.float64)
    pr = np.ascontiguousarray(prior, np.float64)
    fam = np.zeros(F)
    cat = np.zeros((F, k))
    failed = np.zeros(F, np.uint8)
    nf = C.c_int64(0)
    ep, er, en, keep = _err_args(err)
    score = lib().orc_infer_family_likelihoods(th.ref(), counts.ctypes.data_as(_i32p), F, nl, _d(lam), n_lambdas, _d(cp), k,
                                               _d(pr), ep, er, en, mf, mrf, mode, _d(fam), _d(cat),
                                               failed.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nf))
    return {"score": score, "family_lnl": fam, "cat_lk": cat, "failed": failed, "n_failed": nf.value}


def reconstruct(flat, counts, lambdas, prior, mf, mrf):
    th = TreeHandle(flat)
    counts = np.ascontiguousarray(counts, np.int32)
    F, nl = counts.shape
    lam = np.ascontiguousarray(np.atleast_2d(lambdas), np.float64)
    k, n_lambdas = lam.shape
    pr = np.ascontiguousarray(prior, np.float64)
    states = np.zeros((F, k, th.n_internal), np.int32)
    lib().orc_reconstruct(th.ref(), counts.ctypes.data_as(_i32p), F, nl, _d(lam), n_lambdas, k, _d(pr), mf, mrf,
                          states.ctypes.data_as(_i32p))
    return states


def root_max(flat, counts, lambdas, mf, mrf):
    """max_j of the root vector per family (src/probability.cpp:308, 399); lambdas: [n_lambdas] raw."""
    th = TreeHandle(flat)
    counts = np.ascontiguousarray(counts, np.int32)
    F, nl = counts.shape
    lam = np.ascontiguousarray(lambdas, np.float64).ravel()
    out = np.zeros(F)
    rc = lib().orc_root_max(th.ref(), counts.ctypes.data_as(_i32p), F, nl, _d(lam), lam.size, mf, mrf, _d(out))
    if rc:
        raise ValueError("count outside 0..max_family_size")
    return out


def pvalues(cond, observed):
    """cond: [n_root_sizes][n_sim] unsorted simulated likelihoods; observed: [F].  src/probability.cpp:310, 379-409."""
    cond = np.array(cond, np.float64, order="C", copy=True)
    obs = np.ascontiguousarray(observed, np.float64)
    out = np.zeros(len(obs))
    lib().orc_pvalues(_d(cond), cond.shape[0], cond.shape[1], _d(obs), len(obs), _d(out))
    return out


def branch_probabilities(flat, node_sizes, lambdas, mf, mrf, selected=None):
    """compute_viterbi_sum per (family, node), src/gene_family_reconstructor.cpp:361-400; -1 = no value."""
    th = TreeHandle(flat)
    sizes = np.ascontiguousarray(node_sizes, np.int32)
    lam = np.ascontiguousarray(lambdas, np.float64).ravel()
    sel = None if selected is None else np.ascontiguousarray(selected, np.uint8)
    out = np.zeros(sizes.shape)
    lib().orc_branch_probabilities(th.ref(), sizes.ctypes.data_as(_i32p), None if sel is None else sel.ctypes.data_as(C.POINTER(C.c_uint8)),
                                   sizes.shape[0], _d(lam), lam.size, mf, mrf, _d(out))
    return out


def have_ref() -> bool:
    return os.path.exists(REF_HARNESS)


def run_ref(cmd: str, threads: Optional[int] = None, **kw) -> dict:
    """Run the compiled reference harness; returns its JSON line."""
    argv = [REF_HARNESS, cmd]
    for key, val in kw.items():
        if val is None or val is False:
            continue
        argv.append("--" + key)
        if val is not True:
            argv.append(repr(val) if isinstance(val, float) else str(val))
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    res = subprocess.run(argv, check=True, capture_output=True, text=True, env=env)
    line = [l for l in res.stdout.splitlines() if l.startswith("{\"")][-1]
    return json.loads(line)
