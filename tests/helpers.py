"""Shared builders for the parity tests: the reference's own test set-ups as hh_model/SimSpec."""
import datetime as dt

import numpy as np

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi
from hedgehog_jl_b200.engine import SimSpec


def heston_model(S0=100.0, r=0.03, T=1.0, V0=0.04, kappa=2.0, theta=0.04, xi=0.3, rho=-0.7, corr="cholesky", split=True):
    m = abi.hh_model()
    m.kind = abi.HH_MODEL_HESTON
    m.flags = abi.HH_FLAG_SPLIT_STEP if split else 0
    m.S0, m.r, m.T = S0, r, T
    m.V0, m.kappa, m.theta, m.xi, m.rho = V0, kappa, theta, xi, rho
    (m.m11, m.m12, m.m21, m.m22), _ = hh.corr_factor(rho, corr)
    return m


def gbm_model(S0=100.0, r=0.05, sigma=0.2, T=1.0, q1=True):
    m = abi.hh_model()
    m.kind = abi.HH_MODEL_GBM
    m.flags = abi.HH_FLAG_SPLIT_STEP | (abi.HH_FLAG_Q1_SQRT_MEAN if q1 else 0)
    m.S0, m.r, m.T, m.sigma = S0, r, T, sigma
    return m


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def results_tuple(res):
    return [(r.sum, r.sumsq, r.n, r.price, r.std_error, r.n_nonfinite) for r in res]
