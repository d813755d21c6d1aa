"""CPU emulation of fast_exp_full (csrc/hh_fastnormal.cuh): the constants are read from the source, the FMA steps are
emulated in 80-bit arithmetic, and the result must stay within 1.5 ulp of exp over the range the generators use."""
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hedgehog.jl_b200", "csrc", "hh_fastnormal.cuh")


def _constants():
    text = open(SRC).read()
    m = re.search(r"kExpF = \{([^}]*)\}", text)
    vals = [eval(v.strip().replace("\n", "")) for v in m.group(1).split(",")]
    return vals  # magic, scale, neg_hi, neg_lo, inv6, inv24


def fast_exp_emulated(x):
    magic, scale, neg_hi, neg_lo, inv6, inv24 = _constants()
    L = np.longdouble
    t = (L(x) * L(scale) + L(magic)).astype(np.float64)
    nf = t - magic
    n = nf.astype(np.int64)
    r = (L(nf) * L(neg_hi) + L(x)).astype(np.float64)
    r = (L(nf) * L(neg_lo) + L(r)).astype(np.float64)
    tab = np.exp2(np.arange(256) / 256.0)
    e = tab[n & 255]
    p = (L(r) * L(inv24) + L(inv6)).astype(np.float64)
    p = (L(p) * L(r) + L(0.5)).astype(np.float64)
    p = (L(p) * L(r) + L(1.0)).astype(np.float64)
    v = (L(e * r) * L(p) + L(e)).astype(np.float64)
    return np.ldexp(v, (n >> 8).astype(np.int32))


def test_constants_are_the_two_word_split_of_ln2_over_256():
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    magic, scale, neg_hi, neg_lo, inv6, inv24 = _constants()
    c = Decimal(2).ln() / 256
    assert magic == 2.0 ** 52 + 2.0 ** 51
    assert -neg_hi == float(c) and -neg_lo == float(c - Decimal(float(c)))
    assert scale == float(Decimal(256) / Decimal(2).ln())
    assert inv6 == 1.0 / 6 and inv24 == 1.0 / 24


def test_emulated_fast_exp_is_within_1p5_ulp():
    rng = np.random.default_rng(0)
    for lo, hi in ((-2.0, 2.0), (-20.0, 20.0), (-699.0, 699.0), (4.0, 6.0)):
        x = rng.uniform(lo, hi, 400_000)
        got = fast_exp_emulated(x)
        ref = np.exp(np.longdouble(x))
        err_ulp = np.abs(np.longdouble(got) - ref) / np.longdouble(np.spacing(ref.astype(np.float64)))
        assert float(err_ulp.max()) < 1.5, (lo, hi, float(err_ulp.max()))
    # exact at the table nodes' neighbourhood and symmetric handling of negative n
    assert fast_exp_emulated(np.array([0.0]))[0] == 1.0
