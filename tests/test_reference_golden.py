"""Parity against the REAL Hedgehog.jl package, when its outputs are available.

`tools/dump_reference.jl` runs the unmodified reference on small fixed inputs (known Brownian increments through the
reference's own NoiseGrid hook, montecarlo.jl:252-263; `solve` for the exact GBM law and for LSM; the Broadie-Kaya
characteristic-function pieces) and writes `tests/golden/reference_dump/`. This image has no Julia toolchain, so the dump
cannot be produced here: until someone runs the script on a box with Julia >= 1.10 and commits the directory, the tests
below SKIP LOUDLY and per-path parity with the package stays "unpinned" (DESIGN.md section 2).

    julia --project=/path/to/Hedgehog.jl tools/dump_reference.jl /path/to/Hedgehog.jl
    python -m pytest tests/test_reference_golden.py -q            # oracle vs the dump (CPU)
    python -m pytest tests/test_reference_golden.py -q -m gpu     # CUDA vs the dump

What always runs here is `test_consumer_on_a_synthetic_dump`: the checks are exercised end to end on a dump of the same
format written FROM THE ORACLE (layouts, orderings, tolerances of this file — not parity evidence).
HH_REFERENCE_DUMP=<dir> points the tests at another directory."""
import math
import os

import numpy as np
import pytest

from oracle import bk_ref as B
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP_DIR = os.environ.get("HH_REFERENCE_DUMP") or os.path.join(ROOT, "tests", "golden", "reference_dump")
HOWTO = ("no reference dump at %s: PARITY WITH THE JULIA PACKAGE IS UNPINNED. Produce it on a box with Julia >= 1.10 and "
         "Hedgehog's dependencies: julia --project=<Hedgehog.jl> tools/dump_reference.jl <Hedgehog.jl>" % DUMP_DIR)


# ---- the dump format (tools/dump_reference.jl) -------------------------------------------------------------------------
def load_dump(path):
    arrays, meta = {}, {}
    with open(os.path.join(path, "manifest.txt")) as f:
        for line in f:
            parts = line.rstrip("\n").split(" ")
            if parts[0] == "array":
                name, dtype, dims = parts[1], parts[2], [int(x) for x in parts[3:]]
                raw = np.fromfile(os.path.join(path, f"{name}.{dtype}"), dtype="<f8" if dtype == "f64" else "<i8")
                arrays[name] = raw.reshape(dims, order="F")          # Julia arrays are column-major
            elif parts[0] == "meta":
                meta[parts[1]] = " ".join(parts[2:])
    return arrays, meta


def write_dump(path, arrays, meta):
    os.makedirs(path, exist_ok=True)
    lines = []
    for name, a in arrays.items():
        a = np.asarray(a)
        dtype = "i64" if a.dtype.kind in "iu" else "f64"
        a.astype("<i8" if dtype == "i64" else "<f8").ravel(order="F").tofile(os.path.join(path, f"{name}.{dtype}"))
        lines.append(f"array {name} {dtype} " + " ".join(str(d) for d in a.shape))
    lines += [f"meta {k} {v}" for k, v in meta.items()]
    with open(os.path.join(path, "manifest.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


def params_of(meta, key):
    return {k: float(v) for k, v in (kv.split("=") for kv in meta[key].split())}


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


# ---- engines: the oracle and the CUDA library behind one small interface -------------------------------------------------
class OracleSide:
    name = "oracle"

    def __init__(self):
        self.e = O.OracleEngine()

    def heston(self, S0, r, T, V0, kappa, theta, xi, rho, split, M):
        m = O.heston_model(S0, r, T, V0, kappa, theta, xi, rho, split=split)
        m.m11, m.m12, m.m21, m.m22 = M
        return m

    def gbm(self, S0, r, T, sigma, q1=True):
        m = O.o_model()
        m.kind, m.flags = O.HH_MODEL_GBM, O.HH_FLAG_SPLIT_STEP | (O.HH_FLAG_Q1_SQRT_MEAN if q1 else 0)
        m.S0, m.r, m.T, m.sigma = S0, r, T, sigma
        return m

    def sim(self, **kw):
        return O.OSim(**kw)

    C = O

    def terminal(self, model, sim):
        _, t = self.e.mc_european(model, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
        return t

    def terminal_v(self, model, sim):
        return self.e.heston_terminal_v(model, sim)

    def lsm_paths(self, model, sim, payoff, degree, D):
        out, tau, val, paths = self.e.lsm_american(model, sim, payoff, degree, D, want_stopping=True, want_paths=True)
        return out.price, tau, val, paths.T          # [n_steps + 1, ncols] like the reference Matrix


class CudaSide(OracleSide):
    name = "cuda"

    def __init__(self):
        import hedgehog_jl_b200 as hh
        from hedgehog_jl_b200 import _abi as abi
        from hedgehog_jl_b200.engine import SimSpec
        self.hh, self.abi, self.SimSpec = hh, abi, SimSpec
        self.e = hh.default_engine(0)
        self.C = abi

    def heston(self, S0, r, T, V0, kappa, theta, xi, rho, split, M):
        m = self.abi.hh_model()
        m.kind, m.flags = self.abi.HH_MODEL_HESTON, (self.abi.HH_FLAG_SPLIT_STEP if split else 0)
        m.S0, m.r, m.T, m.V0, m.kappa, m.theta, m.xi, m.rho = S0, r, T, V0, kappa, theta, xi, rho
        m.m11, m.m12, m.m21, m.m22 = M
        return m

    def gbm(self, S0, r, T, sigma, q1=True):
        m = self.abi.hh_model()
        m.kind, m.flags = self.abi.HH_MODEL_GBM, self.abi.HH_FLAG_SPLIT_STEP | (self.abi.HH_FLAG_Q1_SQRT_MEAN if q1 else 0)
        m.S0, m.r, m.T, m.sigma = S0, r, T, sigma
        return m

    def sim(self, **kw):
        return self.SimSpec(**kw)

    def terminal_v(self, model, sim):
        return None   # the C ABI returns S_T only; V_T is checked on the oracle side

    def lsm_paths(self, model, sim, payoff, degree, D):
        out, tau, val, paths = self.e.lsm_american(model, sim, payoff, degree, D, want_stopping=True, want_paths=True)
        return out.price, tau, val, paths.T


FACTORS = {   # candidate factors M of [1 rho; rho 1] (M M^T = Gamma), row-major (m11, m12, m21, m22)
    "cholesky": lambda rho: (1.0, 0.0, rho, math.sqrt(1 - rho * rho)),
    "sym_sqrt": lambda rho: ((math.sqrt(1 + rho) + math.sqrt(1 - rho)) / 2, (math.sqrt(1 + rho) - math.sqrt(1 - rho)) / 2,
                             (math.sqrt(1 + rho) - math.sqrt(1 - rho)) / 2, (math.sqrt(1 + rho) + math.sqrt(1 - rho)) / 2),
    "svd": lambda rho: (math.sqrt((1 + rho) / 2), math.sqrt((1 - rho) / 2), math.sqrt((1 + rho) / 2), -math.sqrt((1 - rho) / 2)),
    "svd_flipped": lambda rho: (-math.sqrt((1 + rho) / 2), -math.sqrt((1 - rho) / 2), -math.sqrt((1 + rho) / 2), math.sqrt((1 - rho) / 2)),
}
IDENTITY = (1.0, 0.0, 0.0, 1.0)


# ---- the checks (shared by the real dump and the synthetic self-check) ---------------------------------------------------
def check_heston_em(side, A, meta, tag):
    """Known increments (NoiseGrid): per-path (log S_T, V_T) of StochasticDiffEq's EM() against the restated step."""
    p = params_of(meta, f"{tag}_params")
    T, steps = float(meta[f"{tag}_T"]), int(meta[f"{tag}_steps"])
    W, uT = A[f"{tag}_W"], A[f"{tag}_uT"]                       # [2, steps + 1, npaths], [2, npaths]
    npaths = W.shape[2]
    dW = W[:, 1:, :] - W[:, :-1, :]                               # what NoiseGrid hands to the integrator
    Z = np.ascontiguousarray((dW / math.sqrt(T / steps)).transpose(2, 1, 0))   # Z[path][step][component], identity factor
    C = side.C
    errs = {}
    for split in (True, False):
        m = side.heston(p["S0"], p["r"], T, p["V0"], p["kappa"], p["theta"], p["xi"], p["rho"], split, IDENTITY)
        sim = side.sim(n_paths=npaths, n_steps=steps, scheme=C.HH_SCHEME_EM, rng_mode=C.HH_RNG_NORMALS, normals=Z)
        e = rel(side.terminal(m, sim), np.exp(uT[0]))
        v = side.terminal_v(m, sim)
        if v is not None:
            e = max(e, float(np.max(np.abs(v - uT[1]) / np.maximum(np.abs(uT[1]), 1e-6))))
        errs[split] = e
    assert errs[True] < 1e-12, (f"{side.name}/{tag}: split-step EM (the shipped default HH_FLAG_SPLIT_STEP) does not reproduce the "
                                f"package: rel err {errs[True]:.2e} (split off: {errs[False]:.2e})")
    return errs


def check_gbm_em(side, A, meta):
    p = params_of(meta, "gbm_em_params")
    T, steps = float(meta["gbm_em_T"]), int(meta["gbm_em_steps"])
    W, uT = A["gbm_em_W"], A["gbm_em_uT"]
    npaths = W.shape[2]
    Z = np.ascontiguousarray(((W[:, 1:, :] - W[:, :-1, :]) / math.sqrt(T / steps)).transpose(2, 1, 0))
    C = side.C
    sim = side.sim(n_paths=npaths, n_steps=steps, scheme=C.HH_SCHEME_EM, rng_mode=C.HH_RNG_NORMALS, normals=Z)
    assert rel(side.terminal(side.gbm(p["S0"], p["r"], T, p["sigma"]), sim), np.exp(uT[0])) < 1e-12


def check_correlation_factor(side, A, meta):
    """The package's CorrelatedWienerProcess on replayed normals: which factor of [1 rho; rho 1] does it apply?"""
    if not meta.get("heston_em_corr", "").startswith("ok"):
        pytest.skip("heston_em_corr not in the dump: " + meta.get("heston_em_corr", "absent"))
    p = params_of(meta, "heston_em_params")
    T, steps = float(meta["heston_em_T"]), int(meta["heston_em_steps"])
    Z = np.ascontiguousarray(A["heston_em_Z"].transpose(2, 1, 0))
    uT = A["heston_em_corr_uT"]
    C = side.C
    errs = {}
    for name, f in FACTORS.items():
        m = side.heston(p["S0"], p["r"], T, p["V0"], p["kappa"], p["theta"], p["xi"], p["rho"], True, f(p["rho"]))
        sim = side.sim(n_paths=Z.shape[0], n_steps=steps, scheme=C.HH_SCHEME_EM, rng_mode=C.HH_RNG_NORMALS, normals=Z)
        errs[name] = rel(side.terminal(m, sim), np.exp(uT[0]))
    best = min(errs, key=errs.get)
    assert errs[best] < 1e-12, f"no candidate factor reproduces the package's correlated increments: {errs}"
    return best, errs


def check_gbm_exact_terminal(side, A, meta):
    """solve(prob, MonteCarlo(LognormalDynamics(), BlackScholesExact(), cfg)) at T = 366/365: the sqrt(alpha)-in-the-mean
    quirk (montecarlo.jl:302) must be on for parity."""
    T = float(meta["gbm_exact_terminal_T"])
    Z = np.ascontiguousarray(A["gbm_exact_terminal_Z"].reshape(-1, 1, 1))
    C = side.C
    sim = side.sim(n_paths=Z.shape[0], n_steps=1, scheme=C.HH_SCHEME_EXACT_TERMINAL, rng_mode=C.HH_RNG_NORMALS, normals=Z)
    got = side.terminal(side.gbm(100.0, 0.05, T, 0.2, q1=True), sim)
    assert rel(got, A["gbm_exact_terminal_ST"]) < 1e-12
    if abs(T - 1.0) > 1e-9:
        assert rel(side.terminal(side.gbm(100.0, 0.05, T, 0.2, q1=False), sim), A["gbm_exact_terminal_ST"]) > 1e-9
    if "gbm_exact_terminal_anti_plus" in A:
        sim = side.sim(n_paths=Z.shape[0], n_steps=1, scheme=C.HH_SCHEME_EXACT_TERMINAL, rng_mode=C.HH_RNG_NORMALS, normals=Z,
                       vr=C.HH_VR_ANTITHETIC)
        t = side.terminal(side.gbm(100.0, 0.05, T, 0.2, q1=True), sim)
        n = Z.shape[0]
        assert rel(t[:n], A["gbm_exact_terminal_anti_plus"]) < 1e-12 and rel(t[n:], A["gbm_exact_terminal_anti_minus"]) < 1e-12


def check_lsm(side, A, meta, tag):
    """solve(prob, LSM(...)): the dumped spot grid is regenerated from backed-out normals (exact GBM steps), then the
    backward induction must reproduce stopping_info up to ties of the strict comparison, and the price to 1e-9."""
    p = params_of(meta, f"{tag}_params")
    grid = A[f"{tag}_spot_paths"]                                 # [steps + 1, ncols]
    steps, ncols = grid.shape[0] - 1, grid.shape[1]
    anti = tag.endswith("antithetic")
    n = ncols // 2 if anti else ncols
    dt = p["T"] / steps
    Z = (np.log(grid[1:, :n] / grid[:-1, :n]) - (p["r"] - 0.5 * p["sigma"] ** 2) * dt) / (p["sigma"] * math.sqrt(dt))
    Z = np.ascontiguousarray(Z.T.reshape(n, steps, 1))
    C = side.C
    sim = side.sim(n_paths=n, n_steps=steps, scheme=C.HH_SCHEME_EXACT_STEPS, rng_mode=C.HH_RNG_NORMALS, normals=Z,
                   vr=C.HH_VR_ANTITHETIC if anti else C.HH_VR_NONE)
    D = math.exp(-p["r"] * dt)
    price, tau, val, paths = side.lsm_paths(side.gbm(p["S0"], p["r"], p["T"], p["sigma"]), sim, (p["K"], p["cp"]),
                                            int(meta[f"{tag}_degree"]), D)
    assert rel(paths, grid) < 1e-11                               # the log / exp round trip of the backed-out normals
    ref_tau, ref_val = A[f"{tag}_tau"], A[f"{tag}_value"]
    flips = int(np.sum(tau != ref_tau))
    assert flips <= max(1, ncols // 200), f"{side.name}/{tag}: {flips} of {ncols} stopping times differ from the package"
    same = tau == ref_tau
    assert rel(val[same], ref_val[same]) < 1e-11
    ref_price = float(meta[f"{tag}_price"])
    assert abs(price - ref_price) <= 1e-9 * abs(ref_price) + 2.0 * flips * p["K"] / ncols


def check_gbm_exact_steps(side, A, meta):
    if not meta.get("gbm_exact_steps", "").startswith("ok"):
        pytest.skip("gbm_exact_steps not in the dump: " + meta.get("gbm_exact_steps", "absent"))
    Zg, S = A["gbm_exact_steps_Z"], A["gbm_exact_steps_S"]        # [steps, npaths], [steps + 1, npaths]
    steps, n = Zg.shape
    C = side.C
    sim = side.sim(n_paths=n, n_steps=steps, scheme=C.HH_SCHEME_EXACT_STEPS, rng_mode=C.HH_RNG_NORMALS,
                   normals=np.ascontiguousarray(Zg.T.reshape(n, steps, 1)))
    _, _, _, paths = side.lsm_paths(side.gbm(100.0, 0.05, 1.0, 0.2), sim, (100.0, -1.0), 2, math.exp(-0.05 / steps))
    assert rel(paths, S) < 1e-12


def check_bk_oracle(A, meta, tag):
    """evaluate_chf / moments_from_cf / the series length / the inversion of the package against oracle/bk_ref.py."""
    p = params_of(meta, f"{tag}_params")
    out, phis, mom = A[f"{tag}_out"], A[f"{tag}_phis"], A[f"{tag}_moments"]
    for i in range(out.shape[1]):
        VT, u, mean, var, h, J, x, resid, logIk, _ = out[:, i]
        cf = B.HestonCF(p["kappa"], p["theta"], p["xi"], p["V0"], VT, p["tau"])
        assert abs(cf.logI_k.real - logIk) < 1e-11 * max(1.0, abs(logIk))
        th = math.nan
        for k, a in enumerate((1e-2, 0.0, -1e-2)):
            phi, th = cf.evaluate(a, th)
            assert abs(phi - complex(mom[2 * k, i], mom[2 * k + 1, i])) < 1e-12
        m2, v2 = B.moments_from_cf(cf)
        assert m2 == pytest.approx(mean, rel=1e-9) and v2 == pytest.approx(var, rel=1e-6, abs=1e-12)
        th = math.nan
        for j in range(1, int(J) + 1):
            phi, th = cf.evaluate(h * j, th)
            assert abs(phi - complex(phis[0, j - 1, i], phis[1, j - 1, i])) < 1e-12, (tag, i, j)
        series = B.cf_series(cf, h)
        assert len(series) == int(J)
        assert abs(B.cdf_from_series(series, x, h) - u - resid) < 1e-10      # the same F at the package's root
        o = B.sample_integral_V(cf, u)
        if o["status"] == 0:   # the restated secant passes the package's own acceptance test, like the package's root does
            assert abs(B.cdf_from_series(series, o["x"], h) - u) <= 1e-4 and abs(resid) <= 1e-4


def check_bk_cuda(cuda, abi, A, meta, tag):
    p = params_of(meta, f"{tag}_params")
    out, phis = A[f"{tag}_out"], A[f"{tag}_phis"]
    n = out.shape[1]
    m = abi.hh_model()
    m.kind, m.flags = abi.HH_MODEL_HESTON, abi.HH_FLAG_SPLIT_STEP
    m.S0, m.r, m.T, m.V0, m.kappa, m.theta, m.xi, m.rho = p["S0"], p["r"], p["tau"], p["V0"], p["kappa"], p["theta"], p["xi"], p["rho"]
    m.m11, m.m12, m.m21, m.m22 = FACTORS["cholesky"](p["rho"])
    na = phis.shape[1]
    a = out[4][:, None] * np.arange(1, na + 1)[None, :]
    got = cuda.bk_chf(m, p["tau"], np.full(n, p["V0"]), out[0], a)
    for i in range(n):
        J = int(out[5, i])
        assert np.max(np.abs(got[i, :J] - (phis[0, :J, i] + 1j * phis[1, :J, i]))) < 1e-12
    g = cuda.bk_integral(m, p["tau"], np.full(n, p["V0"]), out[0], out[1])
    assert np.array_equal(g["J"].astype(int), out[5].astype(int))
    assert rel(g["mean"], out[2]) < 1e-7 and np.max(np.abs(g["h"] / out[4] - 1)) < 5e-3
    ok = g["status"] != 2
    # the package's root satisfies |F(x) - u| <= 1e-4 (secant; F' ~ 1 / sd) or sits in a bracket of width 1e-4 (bisection,
    # xtol = atol, sample_from_cf.jl:128); the kernel's root is exact to rounding
    sd = np.sqrt(np.maximum(out[3], 1e-12))
    assert np.all(np.abs(g["x"][ok] - out[6][ok]) <= 1e-3 * sd[ok] + 1.1e-4)


# ---- the real dump --------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dump():
    if not os.path.exists(os.path.join(DUMP_DIR, "manifest.txt")):
        pytest.skip(HOWTO)
    return load_dump(DUMP_DIR)


def test_oracle_matches_the_package(dump):
    A, meta = dump
    side = OracleSide()
    print("reference:", {k: v for k, v in meta.items() if k.startswith(("hedgehog", "julia", "dep_"))})
    for tag in ("heston_em", "heston_em_wild"):
        print(tag, "rel err with split on / off:", check_heston_em(side, A, meta, tag))
    check_gbm_em(side, A, meta)
    check_gbm_exact_terminal(side, A, meta)
    for tag in ("lsm", "lsm_antithetic"):
        check_lsm(side, A, meta, tag)
    for tag in ("bk_c2", "bk_q8", "bk_case1"):
        check_bk_oracle(A, meta, tag)


def test_oracle_correlation_factor_is_the_packages(dump):
    A, meta = dump
    best, errs = check_correlation_factor(OracleSide(), A, meta)
    print("factor of [1 rho; rho 1] applied by CorrelatedWienerProcess:", best, errs)
    # any factor gives the same law; the shipped default is Cholesky — if the package applies another one, only the
    # parity-mode mapping Z -> dW changes: set hh_model.m11..m22 accordingly in the host layers (api.py, HedgehogB200.jl)
    assert best in FACTORS


def test_oracle_gbm_exact_steps_form(dump):
    A, meta = dump
    check_gbm_exact_steps(OracleSide(), A, meta)


@pytest.mark.gpu
def test_cuda_matches_the_package(dump):
    A, meta = dump
    side = CudaSide()
    for tag in ("heston_em", "heston_em_wild"):
        check_heston_em(side, A, meta, tag)
    check_gbm_em(side, A, meta)
    check_gbm_exact_terminal(side, A, meta)
    for tag in ("lsm", "lsm_antithetic"):
        check_lsm(side, A, meta, tag)
    for tag in ("bk_c2", "bk_q8", "bk_case1"):
        check_bk_cuda(side.e, side.abi, A, meta, tag)


# ---- self-check of this consumer on a dump of the same format written from the ORACLE (not parity evidence) -----------------
def synthetic_dump(path):
    e = O.OracleEngine()
    rng = np.random.default_rng(1)
    A, meta = {}, {"hedgehog_version": "synthetic (written from oracle/ by tests/test_reference_golden.py)"}

    def em(tag, p, ncomp, M):
        steps, n, T = 16, 48, 1.0
        Z = rng.standard_normal((ncomp, steps, n))
        W = np.zeros((ncomp, steps + 1, n))
        Mm = np.array(M).reshape(2, 2)[:ncomp, :ncomp]
        for k in range(steps):
            W[:, k + 1, :] = W[:, k, :] + math.sqrt(T / steps) * (Mm @ Z[:, k, :])
        dW = W[:, 1:, :] - W[:, :-1, :]
        Zi = np.ascontiguousarray((dW / math.sqrt(T / steps)).transpose(2, 1, 0))
        if ncomp == 2:
            m = O.heston_model(p["S0"], p["r"], T, p["V0"], p["kappa"], p["theta"], p["xi"], p["rho"], split=True)
            m.m11, m.m12, m.m21, m.m22 = IDENTITY
        else:
            m = OracleSide().gbm(p["S0"], p["r"], T, p["sigma"])
        sim = O.OSim(n_paths=n, n_steps=steps, scheme=O.HH_SCHEME_EM, rng_mode=O.HH_RNG_NORMALS, normals=Zi)
        _, t = e.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
        uT = np.zeros((ncomp, n))
        uT[0] = np.log(t)
        if ncomp == 2:
            uT[1] = e.heston_terminal_v(m, sim)
        A[f"{tag}_Z"], A[f"{tag}_W"], A[f"{tag}_uT"], A[f"{tag}_ST"] = Z, W, uT, t
        meta[f"{tag}_T"], meta[f"{tag}_steps"] = repr(T), steps
        meta[f"{tag}_params"] = " ".join(f"{k}={v}" for k, v in p.items())
        return Z, m

    hp = dict(S0=100.0, r=0.03, V0=0.04, kappa=2.0, theta=0.04, xi=0.3, rho=-0.7, K=100.0, cp=1.0)
    Zh, _ = em("heston_em", hp, 2, FACTORS["cholesky"](-0.7))
    em("heston_em_wild", dict(S0=100.0, r=0.03, V0=0.01, kappa=0.5, theta=0.01, xi=1.0, rho=-0.9, K=100.0, cp=1.0), 2,
       FACTORS["cholesky"](-0.9))
    em("gbm_em", dict(S0=100.0, r=0.05, sigma=0.2, K=100.0, cp=1.0), 1, IDENTITY)
    # "the package's" correlated process: pretend it applies the SVD factor
    m = O.heston_model(100.0, 0.03, 1.0, 0.04, 2.0, 0.04, 0.3, -0.7, split=True)
    m.m11, m.m12, m.m21, m.m22 = FACTORS["svd"](-0.7)
    Zc = np.ascontiguousarray(Zh.transpose(2, 1, 0))
    sim = O.OSim(n_paths=Zc.shape[0], n_steps=16, scheme=O.HH_SCHEME_EM, rng_mode=O.HH_RNG_NORMALS, normals=Zc)
    _, t = e.mc_european(m, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
    A["heston_em_corr_uT"] = np.vstack([np.log(t), e.heston_terminal_v(m, sim)])
    meta["heston_em_corr"] = "ok"
    # exact terminal law at T = 366/365
    T = 366 / 365
    Z1 = rng.standard_normal(500)
    g = OracleSide().gbm(100.0, 0.05, T, 0.2, q1=True)
    for vr, names in ((O.HH_VR_NONE, ("gbm_exact_terminal_ST",)), (O.HH_VR_ANTITHETIC, ("gbm_exact_terminal_anti_plus", "gbm_exact_terminal_anti_minus"))):
        sim = O.OSim(n_paths=500, n_steps=1, scheme=O.HH_SCHEME_EXACT_TERMINAL, rng_mode=O.HH_RNG_NORMALS,
                     normals=np.ascontiguousarray(Z1.reshape(-1, 1, 1)), vr=vr)
        _, t = e.mc_european(g, sim, [(100.0, 1.0)], 1.0, want_terminal=True)
        for k, nm in enumerate(names):
            A[nm] = t[k * 500:(k + 1) * 500]
    A["gbm_exact_terminal_Z"] = Z1
    meta["gbm_exact_terminal_T"] = repr(T)
    # LSM
    for tag, vr in (("lsm", O.HH_VR_NONE), ("lsm_antithetic", O.HH_VR_ANTITHETIC)):
        sim = O.OSim(n_paths=200, n_steps=10, scheme=O.HH_SCHEME_EXACT_STEPS, base_seed=5, vr=vr)
        gm = OracleSide().gbm(100.0, 0.05, 1.0, 0.2)
        out, tau, val, paths = e.lsm_american(gm, sim, (100.0, -1.0), 3, math.exp(-0.05 / 10), want_stopping=True, want_paths=True)
        A[f"{tag}_spot_paths"], A[f"{tag}_tau"], A[f"{tag}_value"] = paths.T, tau.astype(np.int64), val
        meta[f"{tag}_price"], meta[f"{tag}_degree"] = repr(out.price), 3
        meta[f"{tag}_params"] = "S0=100 r=0.05 sigma=0.2 K=100 cp=-1 T=1 steps=10"
    # exact steps on known normals
    Zg = rng.standard_normal((10, 20))
    sim = O.OSim(n_paths=20, n_steps=10, scheme=O.HH_SCHEME_EXACT_STEPS, rng_mode=O.HH_RNG_NORMALS,
                 normals=np.ascontiguousarray(Zg.T.reshape(20, 10, 1)))
    _, _, _, paths = e.lsm_american(OracleSide().gbm(100.0, 0.05, 1.0, 0.2), sim, (100.0, -1.0), 2, math.exp(-0.005),
                                    want_stopping=True, want_paths=True)
    A["gbm_exact_steps_Z"], A["gbm_exact_steps_S"] = Zg, paths.T
    meta["gbm_exact_steps"] = "ok"
    # Broadie-Kaya pieces
    from scipy import stats
    sets = {"bk_c2": dict(S0=100.0, V0=0.04, kappa=2.0, theta=0.04, xi=0.3, rho=-0.7, r=0.03, tau=1 / 12),
            "bk_q8": dict(S0=100.0, V0=1.5, kappa=0.04, theta=0.3, xi=-0.6, rho=0.04, r=0.05, tau=364 / 365),
            "bk_case1": dict(S0=100.0, V0=0.010201, kappa=6.21, theta=0.019, xi=0.61, rho=-0.7, r=0.0319, tau=1.0)}
    for tag, p in sets.items():
        d, lam_s, c = B.vt_params(p["kappa"], p["theta"], p["xi"], 1.0, p["tau"])
        n = 6
        out, mom, series_all = np.zeros((10, n)), np.zeros((6, n)), []
        for i in range(n):
            VT = c * stats.ncx2.rvs(d, lam_s * p["V0"], random_state=rng)
            u = rng.uniform(0.02, 0.98)
            cf = B.HestonCF(p["kappa"], p["theta"], p["xi"], p["V0"], VT, p["tau"])
            th = math.nan
            for k, a in enumerate((1e-2, 0.0, -1e-2)):
                phi, th = cf.evaluate(a, th)
                mom[2 * k, i], mom[2 * k + 1, i] = phi.real, phi.imag
            o = B.sample_integral_V(cf, u)
            series = B.cf_series(cf, o["h"])
            series_all.append(series)
            out[:, i] = (VT, u, o["mean"], o["var"], o["h"], len(series), o["x"], B.cdf_from_series(series, o["x"], o["h"]) - u,
                         cf.logI_k.real, B.cdf_from_series(series, o["max_guess"], o["h"]) - u)
        J = max(len(s) for s in series_all)
        phis = np.full((2, J, n), np.nan)
        for i, s in enumerate(series_all):
            phis[0, :len(s), i], phis[1, :len(s), i] = np.real(s), np.imag(s)
        A[f"{tag}_out"], A[f"{tag}_phis"], A[f"{tag}_moments"] = out, phis, mom
        meta[f"{tag}_params"] = " ".join(f"{k}={v!r}" for k, v in p.items())
    write_dump(path, A, meta)


def test_consumer_on_a_synthetic_dump(tmp_path):
    """Not parity evidence: exercises the loader and every check of this file on a dump written from the oracle."""
    synthetic_dump(str(tmp_path))
    A, meta = load_dump(str(tmp_path))
    side = OracleSide()
    for tag in ("heston_em", "heston_em_wild"):
        errs = check_heston_em(side, A, meta, tag)
        assert errs[False] > 1e-9          # the two step forms are distinguishable on these inputs
    check_gbm_em(side, A, meta)
    best, errs = check_correlation_factor(side, A, meta)
    assert best == "svd" and errs["cholesky"] > 1e-6   # the synthetic "package" was given the SVD factor
    check_gbm_exact_terminal(side, A, meta)
    for tag in ("lsm", "lsm_antithetic"):
        check_lsm(side, A, meta, tag)
    check_gbm_exact_steps(side, A, meta)
    for tag in ("bk_c2", "bk_q8", "bk_case1"):
        check_bk_oracle(A, meta, tag)


@pytest.mark.gpu
def test_cuda_consumer_on_a_synthetic_dump(tmp_path):
    """The CUDA side of the same plumbing (kernels against an oracle-written dump: the parity the other tests hold too)."""
    synthetic_dump(str(tmp_path))
    A, meta = load_dump(str(tmp_path))
    side = CudaSide()
    for tag in ("heston_em", "heston_em_wild"):
        check_heston_em(side, A, meta, tag)
    check_gbm_em(side, A, meta)
    check_gbm_exact_terminal(side, A, meta)
    for tag in ("lsm", "lsm_antithetic"):
        check_lsm(side, A, meta, tag)
    check_gbm_exact_steps(side, A, meta)
    for tag in ("bk_c2", "bk_q8", "bk_case1"):
        check_bk_cuda(side.e, side.abi, A, meta, tag)
