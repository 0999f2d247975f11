"""Host-side emulation of the K-build's per-entry arithmetic (gumbi_b200/csrc/kbuild_persist.cuh::kb4_eval), bit for bit where the kernel
works on bit patterns: the exp table with biased high words, the one-instruction binary exponent, the two-sided integer clip on the
high word, the group-wise underflow flush.  The CUDA kernel itself is parity-tested on the device (tests/test_gpu_parity.py); this file
keeps the integer tricks honest on a box without a GPU (the fused multiply-adds are emulated in extended precision, so agreement is to
a few ulps, not to the bit)."""
from decimal import Decimal, getcontext

import numpy as np
import pytest

TAB, TAB_LOG2 = 2048, 11
SHIFT = 20 - TAB_LOG2
MAGIC = 6755399441055744.0           # 1.5 * 2^52
LOG2E = np.longdouble("1.442695040888963407359924681001892137")
LN2 = np.longdouble("0.693147180559945309417232121458176568")
getcontext().prec = 50


def _hi(x):
    return (np.asarray(x, dtype=np.float64).view(np.int64) >> 32).astype(np.int32)


def _lo(x):
    return np.asarray(x, dtype=np.float64).view(np.int64) & 0xFFFFFFFF


def _hilo(hi, lo):
    return ((hi.astype(np.int64) << 32) | lo).view(np.float64)


def _fma(a, b, c):
    return (np.longdouble(a) * np.longdouble(b) + np.longdouble(c)).astype(np.float64)


def constants(kind):
    """kb4_set_constants: zs = 1/2 for ExpQuad (exp(-z/2)), 1 for the Matern family (exp(-w))."""
    zs = 0.5 if kind == "ExpQuad" else 1.0
    c = {"ExpQuad": 1.0, "Matern52": 5.0}[kind]
    zmin = 0.0 if kind == "ExpQuad" else c * 1e-12
    zmax = 2000.0 if kind == "ExpQuad" else 1.0e6
    return dict(zs=zs, cA=-float(LOG2E * np.longdouble(zs * TAB)), cR=float(LN2 / np.longdouble(zs * TAB)), q3=-zs ** 3 / 6.0,
                hi_zmin=int(_hi(zmin)), hi_zmax=int(_hi(zmax)))


def table(eta2):
    """sTab: eta^2 2^(j/2048) with the high word biased by -(j << 9)."""
    t = eta2 * np.exp2(np.arange(TAB) / TAB)
    return _hi(t) - (np.arange(TAB, dtype=np.int32) << SHIFT), _lo(t)


def kb4_eval(z, kind, eta2):
    k = constants(kind)
    zs = k["zs"]
    z = np.asarray(z, dtype=np.float64)
    # two-sided clip on the high word, low word untouched
    x = _hilo(np.minimum(np.maximum(_hi(z), np.int32(k["hi_zmin"])), np.int32(k["hi_zmax"])), _lo(z))
    if kind != "ExpQuad":
        x = np.sqrt(x)                                        # the kernel: MUFU seed + third-order step, residual 2^-67
    t = _fma(x, k["cA"], MAGIC)
    n = _lo(t).astype(np.uint32).view(np.int32)
    rr = _fma(t - MAGIC, k["cR"], x)
    # half a table step (+ the double rounding of this emulation: long double sum, then double -- up to 2^-11 of a step; the kernel's FMA rounds once)
    assert np.all(np.abs(rr) <= float(LN2) / (2 * TAB * zs) * (1 + 1e-3))
    m = rr * _fma(_fma(rr, k["q3"], 0.5 * zs * zs), rr, -zs)
    thi, tlo = table(eta2)
    j = n & (TAB - 1)
    with np.errstate(over="ignore"):
        hi = (thi[j].astype(np.int64) + (n.astype(np.int64) << SHIFT))
    assert np.all((hi > -2 ** 31) & (hi < 2 ** 31)), "the 32-bit exponent arithmetic must not wrap"
    hi = hi.astype(np.int32)
    flush = hi < 0x00100000
    Ts = _hilo(np.where(flush, 0, hi).astype(np.int32), np.where(flush, 0, tlo[j]))
    e = _fma(Ts, m, Ts)
    if kind == "Matern52":
        e = e * _fma(_fma(1.0 / 3.0, x, 1.0), x, 1.0)
    return e


def reference(z, kind, eta2):
    out = []
    for v in np.asarray(z, dtype=np.float64):
        if kind == "ExpQuad":
            out.append(Decimal(eta2) * (Decimal(max(float(v), 0.0)) * Decimal("-0.5")).exp())
        else:
            w = Decimal(max(float(v), 5e-12)).sqrt()
            out.append(Decimal(eta2) * (1 + w + w * w / 3) * (-w).exp())
    return np.array([float(o) for o in out])


@pytest.mark.parametrize("kind", ["ExpQuad", "Matern52"])
@pytest.mark.parametrize("eta2", [1.3, 1e-200, 1e100])
def test_entry_arithmetic_matches_a_50_digit_evaluation(kind, eta2):
    rng = np.random.default_rng(3)
    z = np.concatenate([rng.uniform(0, 60, 4000), rng.uniform(0, 3, 2000), [0.0, 1e-300, 7e-12, 59.999]])
    if kind == "Matern52":
        z = np.concatenate([z ** 2, [5e-12, 6e-12]])
    got, ref = kb4_eval(z, kind, eta2), reference(z, kind, eta2)
    # relative error = absolute error of the exp argument (up to 60 here): |x| 2^-54 from the one-step reduction, for Matern |w| 2^-53
    # from the rounding of the square root itself, + a few ulps
    np.testing.assert_allclose(got, ref, rtol=1.5e-14)


@pytest.mark.parametrize("kind", ["ExpQuad", "Matern52"])
def test_clip_cap_and_flush(kind):
    big = np.array([1e3, 1386.0, 5e3, 1e6, 1e9, 1e15, 1e300])
    with np.errstate(over="ignore"):
        z = big if kind == "ExpQuad" else big ** 2          # the last one overflows to +inf: clipped like any huge distance
    z = np.concatenate([z, [-1e-16, -3e-13, 0.0, 5e-324]])
    for eta2 in (1.3, 1e-250, 1e100):
        got = kb4_eval(z, kind, eta2)
        assert np.all(np.isfinite(got)) and np.all(got >= 0.0)
        # exp argument beyond -745: the oracle's exp underflows to 0 and the kernel returns an exact 0 (for any eta^2 < 1e125)
        n_big = len(big)
        arg = 0.5 * big if kind == "ExpQuad" else big
        assert np.all(got[:n_big][arg > 746] == 0.0)
        # negative rounding residues and zeros are clipped: value at r = 0 (+ the 1e-12 of euclidean_dist under the square root)
        at0 = reference([0.0], kind, eta2)[0]
        np.testing.assert_allclose(got[n_big:], at0, rtol=5e-12)


def test_table_bias_cancels_the_index_bits():
    thi, _ = table(1.0)
    for n in (0, -1, -2047, -2048, -2049, -123456, -2_000_000):
        j = n & (TAB - 1)
        hi = int(thi[j]) + (n << SHIFT)
        want = int(_hi(np.exp2(j / TAB))) + ((n >> TAB_LOG2) << 20)
        assert hi == want
