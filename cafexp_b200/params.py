"""Host-side parameter helpers of the likelihood path, mirroring the reference's scalar host code so that a Python
host can drive the C ABI the way the reference's C++ does (the GPU only ever receives k multipliers and a prior table):

    get_gamma(k, alpha)       discrete-gamma rate multipliers, category means       src/gamma.cpp:15-241  (PAML, Yang 1994)
    prior_uniform(...)        uniform_distribution::compute over a root distribution src/root_distribution.cpp:15-31,
                                                                                     src/root_equilibrium_distribution.cpp:20-32
    prior_poisson(...)        poisson_distribution::compute                          src/poisson.cpp:19-36,
                                                                                     src/root_equilibrium_distribution.h:45-51

Plain double arithmetic with the C library's exp / log / pow / lgamma (ctypes for lgamma: CPython's own differs in the
last bits), the published algorithms (AS 32 incomplete gamma, AS 70 normal point, AS 91 chi-square point) with the
reference's constants and tolerances, including its truncated ln 2 and its use of the previous continued-fraction
convergent.  Checked against the reference's outputs in tests/test_host_cpu.py.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
from typing import Dict, Optional, Tuple

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.lgamma.restype = ctypes.c_double
_libm.lgamma.argtypes = [ctypes.c_double]


def _lgamma(x: float) -> float:
    return _libm.lgamma(x)


def incomplete_gamma(x: float, alpha: float, ln_gamma_alpha: float) -> float:
    """AS 32 (src/gamma.cpp:66-116): series for x <= 1 or x < alpha, continued fraction otherwise."""
    accurate, overflow = 1e-8, 1e30
    if x == 0:
        return 0.0
    if x < 0 or alpha <= 0:
        return -1.0
    factor = math.exp(alpha * math.log(x) - x - ln_gamma_alpha)
    if not (x > 1 and x >= alpha):
        total, term, rn = 1.0, 1.0, alpha
        while True:
            rn += 1
            term *= x / rn
            total += term
            if not term > accurate:
                break
        return total * (factor / alpha)
    a, count = 1 - alpha, 0.0
    b = a + x + 1
    p0, p1, p2, p3 = 1.0, x, x + 1, x * b
    gin = p2 / p3
    while True:
        a += 1
        b += 2
        count += 1
        an = a * count
        p4 = b * p2 - an * p0
        p5 = b * p3 - an * p1
        if p5 != 0:
            rn = p4 / p5
            dif = abs(gin - rn)
            if dif <= accurate and dif <= accurate * rn:
                return 1 - factor * gin          # the previous convergent, as the reference
            gin = rn
        p0, p1, p2, p3 = p2, p3, p4, p5
        if abs(p4) >= overflow:
            p0, p1, p2, p3 = p0 / overflow, p1 / overflow, p2 / overflow, p3 / overflow


def point_normal(prob: float) -> float:
    """AS 70 (src/gamma.cpp:203-215)."""
    a = (-.322232431088, -1.0, -.342242088547, -.0204231210245, -.453642210148e-4)
    b = (.0993484626060, .588581570495, .531103462366, .103537752850, .0038560700634)
    tail = prob if prob < 0.5 else 1 - prob
    if tail < 1e-20:
        return -9999.0
    y = math.sqrt(math.log(1 / (tail * tail)))
    z = y + ((((y * a[4] + a[3]) * y + a[2]) * y + a[1]) * y + a[0]) / ((((y * b[4] + b[3]) * y + b[2]) * y + b[1]) * y + b[0])
    return -z if prob < 0.5 else z


def point_chi2(prob: float, v: float) -> float:
    """AS 91 (src/gamma.cpp:129-186): e = .5e-6, ln 2 truncated to .6931471805 as in the reference."""
    e, aa = .5e-6, .6931471805
    p = prob
    if p < .000002 or p > .999998 or v <= 0:
        return -1.0
    g = _lgamma(v / 2)
    xx = v / 2
    c = xx - 1
    if v < -1.24 * math.log(p):
        ch = math.pow(p * xx * math.exp(g + xx * aa), 1 / xx)
        if ch - e < 0:
            return ch
    elif v <= .32:
        ch = 0.4
        a = math.log(1 - p)
        while True:
            q = ch
            p1 = 1 + ch * (4.67 + ch)
            p2 = ch * (6.73 + ch * (6.66 + ch))
            t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2
            ch -= (1 - math.exp(a + g + .5 * ch + c * aa) * p2 / p1) / t
            if not abs(q / ch - 1) - .01 > 0:
                break
    else:
        x = point_normal(p)
        p1 = 0.222222 / v
        ch = v * math.pow(x * math.sqrt(p1) + 1 - p1, 3.0)
        if ch > 2.2 * v + 6:
            ch = -2 * (math.log(1 - p) - c * math.log(.5 * ch) + g)
    while True:
        q = ch
        p1 = .5 * ch
        t = incomplete_gamma(p1, xx, g)
        if t < 0:
            return -1.0
        p2 = p - t
        t = p2 * math.exp(xx * aa + g + p1 - c * math.log(ch))
        b = t / ch
        a = 0.5 * t - b * c
        s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420
        s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520
        s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520
        s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040
        s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520
        s6 = (120 + c * (346 + 127 * c)) / 5040
        ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))))
        if not abs(q / ch - 1) > e:
            break
    return ch


def get_gamma(k: int, alpha: float) -> Tuple[np.ndarray, np.ndarray]:
    """(category probabilities, rate multipliers): k equiprobable categories, multipliers = category means
    (src/gamma.cpp:15-52 with median == 0, :225-241 with beta == alpha)."""
    beta = alpha
    factor = alpha / beta * k
    lnga1 = _lgamma(alpha + 1)
    freq = [0.0] * k
    rate = [0.0] * k
    for i in range(k - 1):
        freq[i] = point_chi2((i + 1.0) / k, 2.0 * alpha) / (2.0 * beta)
    for i in range(k - 1):
        freq[i] = incomplete_gamma(freq[i] * beta, alpha + 1, lnga1)
    if k == 1:
        rate[0] = factor            # a single category: the whole distribution, mean 1
    else:
        rate[0] = freq[0] * factor
        rate[k - 1] = (1 - freq[k - 2]) * factor
        for i in range(1, k - 1):
            rate[i] = (freq[i] - freq[i - 1]) * factor
    return np.full(k, 1.0 / k), np.asarray(rate)


def _expand_rootdist(rootdist: Optional[Dict[int, int]], mrf: int):
    """root_distribution::vectorize / vectorize_uniform (src/root_distribution.cpp:15-31)."""
    if not rootdist:
        return [1] * mrf
    out = []
    for size in sorted(rootdist):
        out.extend([size] * max(int(rootdist[size]), 0))
    return out


def prior_uniform(mrf: int, rootdist: Optional[Dict[int, int]] = None, n_out: Optional[int] = None) -> np.ndarray:
    """(double)(float)(list[val] / sum) for val = 0 .. n_out-1, 0 past the end of the list
    (src/root_equilibrium_distribution.cpp:20-32)."""
    lst = _expand_rootdist(rootdist, mrf)
    n_out = mrf if n_out is None else n_out
    total = np.float32(sum(lst))
    out = np.zeros(n_out)
    for v in range(min(n_out, len(lst))):
        out[v] = float(np.float32(lst[v]) / total)
    return out


def prior_poisson(poisson_lambda: float, mrf: int, rootdist: Optional[Dict[int, int]] = None, n_out: Optional[int] = None) -> np.ndarray:
    """(double)(float) poisspdf(val, lambda) for val below the root-distribution list length
    (src/poisson.cpp:19-36, src/root_equilibrium_distribution.h:45-51)."""
    n = len(_expand_rootdist(rootdist, mrf))
    n_out = mrf if n_out is None else n_out
    out = np.zeros(n_out)
    for v in range(min(n_out, n)):
        out[v] = float(np.float32(math.exp(v * math.log(poisson_lambda) - _lgamma(float(v + 1)) - poisson_lambda)))
    return out
