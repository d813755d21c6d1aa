"""An independent numpy restatement of SURVEY.md Appendix A (A.1-A.4, A.6), vectorised over paths, against the C oracle in
parity mode (caller-supplied normals). Two restatements in two languages written from the same cited reference lines:
a transcription slip in either shows up as a per-path difference. Tolerance 1e-13 (libm vs numpy exp/log last bits)."""
import math

import numpy as np
import pytest

from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec
from helpers import gbm_model, heston_model, rel_err


def np_heston_em(m, z, split, anti):
    """heston.jl:7-31 + EM{split} [upstream]: K = u + dt f(u); u' = K + g(K or u) dW; dW = sqrt(dt) M z"""
    n, M, _ = z.shape
    dt = m.T / M
    sq = math.sqrt(dt)
    out = []
    for sign in ((1.0, -1.0) if anti else (1.0,)):
        x = np.full(n, math.log(m.S0))
        v = np.full(n, m.V0)
        for k in range(M):
            dW1 = sign * (sq * m.m11 * z[:, k, 0] + sq * m.m12 * z[:, k, 1])
            dW2 = sign * (sq * m.m21 * z[:, k, 0] + sq * m.m22 * z[:, k, 1])
            vp = np.maximum(v, 0.0)
            K1 = x + dt * (m.r - 0.5 * vp)
            K2 = v + dt * (m.kappa * (m.theta - vp))
            s = np.sqrt(np.maximum(K2 if split else v, 0.0))
            x = K1 + s * dW1
            v = K2 + (m.xi * s) * dW2
        out.append(np.exp(x))
    return np.concatenate(out)


def np_gbm(m, z, scheme, anti):
    n, M, _ = z.shape
    if scheme == abi.HH_SCHEME_EXACT_TERMINAL:  # montecarlo.jl:293-303, 384-390 (Q1: sqrt(alpha) in the mean)
        a = m.T
        c = math.sqrt(a) if (m.flags & abi.HH_FLAG_Q1_SQRT_MEAN) else a
        mu = math.log(m.S0) + (m.r - m.sigma ** 2 / 2) * c
        X = mu + m.sigma * math.sqrt(a) * z[:, 0, 0]
        return np.concatenate([np.exp(X), np.exp(2 * mu - X)]) if anti else np.exp(X)
    dt = m.T / M
    sq = math.sqrt(dt)
    drift = m.r - 0.5 * (m.sigma * m.sigma)
    out = []
    for sign in ((1.0, -1.0) if anti else (1.0,)):
        if scheme == abi.HH_SCHEME_EM:  # heston.jl:33-52
            x = np.full(n, math.log(m.S0))
            for k in range(M):
                x = (x + dt * drift) + m.sigma * (sign * (sq * z[:, k, 0]))
            out.append(np.exp(x))
        else:  # montecarlo.jl:140-159 + GeometricBrownianMotionProcess [upstream]; antithetic: sigma -> -sigma (:270-284)
            S = np.full(n, m.S0)
            for k in range(M):
                S = S + S * (np.exp(drift * dt + sign * (m.sigma * sq * z[:, k, 0])) - 1.0)
            out.append(S)
    return np.concatenate(out)


@pytest.mark.parametrize("anti", [0, 1])
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("corr", ["cholesky", "sym_sqrt", "svd"])
def test_heston_em_restatements_agree(oracle, anti, split, corr):
    m = heston_model(corr=corr, split=split, xi=0.5)
    z = np.random.Generator(np.random.Philox(31)).standard_normal((400, 30, 2))
    sim = SimSpec(n_paths=400, n_steps=30, vr=anti, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    _, term = oracle.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    assert rel_err(term, np_heston_em(m, z, split, anti)) < 1e-13


@pytest.mark.parametrize("anti", [0, 1])
@pytest.mark.parametrize("scheme,steps,T", [(abi.HH_SCHEME_EM, 12, 1.0), (abi.HH_SCHEME_EXACT_STEPS, 12, 1.0),
                                            (abi.HH_SCHEME_EXACT_TERMINAL, 1, 366 / 365)])
def test_gbm_restatements_agree(oracle, anti, scheme, steps, T):
    m = gbm_model(T=T)
    z = np.random.Generator(np.random.Philox(32)).standard_normal((500, steps, 1))
    sim = SimSpec(n_paths=500, n_steps=steps, scheme=scheme, vr=anti, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    res, term = oracle.mc_european(m, sim, [(100.0, 1.0), (100.0, -1.0)], math.exp(-m.r * m.T), want_terminal=True)
    want = np_gbm(m, z, scheme, anti)
    assert rel_err(term, want) < 1e-13
    # payoff, antithetic pair average, discount, mean   (payoffs.jl:154-156, montecarlo.jl:428-432, 489-490)
    for r, cp in zip(res, (1.0, -1.0)):
        pay = np.maximum(cp * (want - 100.0), 0.0)
        if anti:
            pay = 0.5 * (pay[:500] + pay[500:])
        assert abs(r.price - math.exp(-m.r * m.T) * pay.mean()) < 1e-12 * max(r.price, 1.0)


def test_lsm_backward_induction_restatements_agree(oracle):
    """least_squares_montecarlo.jl:99-136 with numpy.polyfit (QR-based least squares, like Polynomials.fit)."""
    m = gbm_model()
    n, M, deg = 3000, 10, 3
    z = np.random.Generator(np.random.Philox(33)).standard_normal((n, M, 1))
    sim = SimSpec(n_paths=n, n_steps=M, scheme=abi.HH_SCHEME_EXACT_STEPS, rng_mode=abi.HH_RNG_NORMALS, normals=z)
    D = math.exp(-m.r * m.T / M)
    o, tau, val, paths = oracle.lsm_american(m, sim, (100.0, -1.0), deg, D, want_stopping=True, want_paths=True)
    G = paths.T  # (M+1, n)
    t_ = np.full(n, M)
    v_ = np.maximum(100.0 - G[M], 0.0)
    for t in range(M - 1, 0, -1):
        y = D ** (t_ - t) * v_
        e = np.maximum(100.0 - G[t], 0.0)
        itm = e > 0
        if not itm.any():
            continue
        beta = np.polyfit(G[t][itm], y[itm], deg)
        cont = np.polyval(beta, G[t])
        ex = itm & (e > cont)
        t_[ex] = t
        v_[ex] = e[ex]
    price = float(np.mean(D ** t_ * v_))
    flips = int(np.sum(t_ != tau))
    assert flips <= 2, flips
    assert abs(o.price - price) < (1e-9 if flips == 0 else 1e-4) * price
