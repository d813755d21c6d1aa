#!/usr/bin/env python
"""Coefficients of Debye's uniform expansion of I_nu (DLMF 10.41.3, 10.41.7-9), as pasted into log_besseli_debye
(hedgehog.jl_b200/csrc/hh_bessel.cuh):  u_k(t) = t^k P_k(t^2),
    u_{k+1}(t) = 1/2 t^2 (1 - t^2) u_k'(t) + 1/8 int_0^t (1 - 5 s^2) u_k(s) ds,   u_0 = 1,
in exact rational arithmetic.   python tools/gen_debye.py [K=8]   prints P_0 .. P_K, lowest power first."""
import sys
from fractions import Fraction as Fr


def debye_polys(K):
    us = [{0: Fr(1)}]
    for _ in range(K):
        nxt = {}
        for p, c in us[-1].items():
            if p > 0:
                nxt[p + 1] = nxt.get(p + 1, 0) + Fr(1, 2) * c * p
                nxt[p + 3] = nxt.get(p + 3, 0) - Fr(1, 2) * c * p
            nxt[p + 1] = nxt.get(p + 1, 0) + Fr(1, 8) * c / (p + 1)
            nxt[p + 3] = nxt.get(p + 3, 0) - Fr(5, 8) * c / (p + 3)
        us.append({p: c for p, c in nxt.items() if c != 0})
    return us


if __name__ == "__main__":
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    for k, u in enumerate(debye_polys(K)):
        assert set(u) <= {k + 2 * j for j in range(k + 1)}
        print("      " + ", ".join(repr(float(u.get(k + 2 * j, 0))) for j in range(k + 1)) + ",   // u_%d" % k)
