#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric: Heston Euler-Maruyama path-steps/sec (config C2).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference arm: CPU implementation of the same path

A "step" is one pass of the hot path over one batch of synthetic input: pricing the C2 contract
(S0=K=100, r=0.03, V0=0.04, kappa=2, theta=0.04, xi=0.3, rho=-0.7, T=1, call) with `--paths` trajectories
x 252 Euler-Maruyama steps per GPU in Float64 (weak scaling: per-GPU work fixed, disjoint Philox streams).

Keys (see the task contract):
  value     whole-job path-steps/s, kernels only (inputs are a few scalars, already on the device)
  e2e       the same through the public API hedgehog_jl_b200.solve(problem, method) with host buffers
  roofline  ALGORITHMIC FP64 work (25 FLOP per path-step) / average kernel time, against the FP64
            DFMA peak measured in this run (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline  the CPU oracle (a C restatement of the reference's arithmetic, "port") on a bounded sample
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "heston_em_path_steps_per_sec"
UNIT = "path-steps/s"
FLOP_PER_PATH_STEP = 25.0  # 17 (SDE update) + 8 (Box-Muller scaling); transcendentals not counted
# executed by heston_fast2_kernel per path-step, from ncu (profiles/r1_c_ncu_heston_fast_v2.csv): 31.1 FP64 instructions
# = 51.2 FLOP (DFMA counted twice), 90.6 instructions in all
EXEC_FLOP_PER_PATH_STEP = 51.16
EXEC_FP64_INSTR_PER_PATH_STEP = 31.09
EXEC_INSTR_PER_PATH_STEP = 90.6
WORKLOAD = "C2 Heston EM European call: 1e8 paths x 252 steps per GPU, f64, NoVarianceReduction"
EULER_BIAS_252 = 0.005651  # price(252 steps) - Carr-Madan, 2e8 paths, std error 0.0005 (tools/euler_bias.py)
CARR_MADAN_C2 = 9.242536279428904  # oracle/anchors.py heston_price(100,100,.03,1,.04,2,.04,.3,-.7), CarrMadan(1, 32)


def c2_problem(hh, total_paths, nsteps, precision="f64", ensemble=False, base_seed=42):
    import datetime as dt
    payoff = hh.VanillaOption(100.0, dt.date(2020, 12, 31), hh.European(), hh.Call(), hh.Spot())  # 365 days: T = 1
    market = hh.HestonInputs(dt.date(2020, 1, 1), 0.03, 100.0, 0.04, 2.0, 0.04, 0.3, -0.7)
    prob = hh.PricingProblem(payoff, market)
    cfg = hh.SimulationConfig(total_paths, steps=nsteps, base_seed=base_seed)
    method = hh.MonteCarlo(hh.HestonDynamics(), hh.EulerMaruyama(), cfg, precision=precision, ensemble=ensemble)
    return prob, method


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.15 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = max(float(r[2]) for r in rows)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": pw, "samples": len(rows),
                "reasons": reasons}


def cpu_sample(paths, nsteps, seconds=None, reps=1):
    """Time the CPU oracle on `paths` x `nsteps` of the C2 workload, all host threads. Returns (path-steps/s, info)."""
    import hedgehog_jl_b200 as hh
    from hedgehog_jl_b200 import _abi as abi
    from hedgehog_jl_b200.engine import SimSpec
    from oracle import oracle as O
    eng = O.OracleEngine()
    m = abi.hh_model()
    m.kind, m.flags = abi.HH_MODEL_HESTON, abi.HH_FLAG_SPLIT_STEP
    m.S0, m.r, m.T = 100.0, 0.03, 1.0
    m.V0, m.kappa, m.theta, m.xi, m.rho = 0.04, 2.0, 0.04, 0.3, -0.7
    (m.m11, m.m12, m.m21, m.m22), _ = hh.corr_factor(m.rho, "cholesky")
    D = math.exp(-0.03)
    done, t_used, price = 0, 0.0, None
    rep = 0
    while True:
        sim = SimSpec(n_paths=paths, n_steps=nsteps, scheme=abi.HH_SCHEME_EM, base_seed=42 + rep, path_offset=rep * paths)
        t0 = time.perf_counter()
        res, _ = eng.mc_european(m, sim, [(100.0, 1.0)], D)
        t_used += time.perf_counter() - t0
        done += paths * nsteps
        price = res[0].price
        rep += 1
        if seconds is None:
            if rep >= reps:
                break
        elif t_used >= seconds:
            break
    return done / t_used, {"cores": eng.threads, "paths": paths * rep, "seconds": t_used, "price": price}


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU path. The Julia package cannot run in this image (no julia), so this is
    the C restatement (oracle/, OpenMP over all host cores); see BASELINE.md §3."""
    if rank != 0:
        return
    sample_paths = args.ref_paths
    for _ in range(args.warmup):
        cpu_sample(max(sample_paths // 10, 1000), args.nsteps)
    t0 = time.perf_counter()
    total = 0
    cores = 1
    for k in range(args.steps):
        v, info = cpu_sample(sample_paths, args.nsteps)
        total += sample_paths * args.nsteps
        cores = info["cores"]
    dt_s = time.perf_counter() - t0
    value = total / dt_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "paths_per_gpu": args.paths, "n_steps": args.nsteps,
                   "sample": f"each step prices {sample_paths} of the workload's trajectories x {args.nsteps} steps on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x {sample_paths} paths x {args.nsteps} steps, OpenMP over all host threads; "
                                   "C restatement of the reference arithmetic (the Julia package cannot run here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--paths", type=int, default=100_000_000, help="trajectories per GPU per step (C2: 1e8)")
    ap.add_argument("--nsteps", type=int, default=252)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-paths", type=int, default=400_000)
    ap.add_argument("--skip-ensemble", action="store_true")
    ap.add_argument("--skip-f32", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import hedgehog_jl_b200 as hh
    from hedgehog_jl_b200 import _abi as abi
    from hedgehog_jl_b200.api import _model_of, _scheme_of, _sim_of

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = hh.default_engine(local_rank)
    # a non-default torch stream shared with the library, so torch.cuda.Event brackets exactly its kernels
    # (torch's legacy default stream is the NULL handle, which hh_set_stream treats as "use the ctx stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fp64_peak, _ = eng.fp64_peak()

    K, W = args.steps, args.warmup
    total_paths = args.paths * world
    prob, method = c2_problem(hh, total_paths, args.nsteps, args.precision)
    mdl = _model_of(prob, method)
    payoffs = [(100.0, 1.0)]
    D = hh.df(prob.market_inputs.rate, prob.payoff.expiry)

    def sim_for(k):
        s = _sim_of(method, _scheme_of(method), (rank, world))
        s.base_seed = 42 + k  # every step simulates fresh trajectories
        return s

    # ---- kernels only ("value") --------------------------------------------------------------------------------
    for k in range(W):
        eng.mc_european_launch(mdl, sim_for(1000 + k), payoffs)
        eng.mc_european_collect(D)
    sampler = ClockSampler(local_rank)
    time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record(stream)
    for k in range(K):
        eng.mc_european_launch(mdl, sim_for(k), payoffs)
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    last = eng.mc_european_collect(D)[0]
    clocks = sampler.stop(t_wall0, t_wall1)
    path_steps_per_step = float(total_paths) * args.nsteps
    value = path_steps_per_step * K / (ms * 1e-3)
    per_launch_ms = ms / K
    achieved_tflops = FLOP_PER_PATH_STEP * float(args.paths) * args.nsteps / (per_launch_ms * 1e-3) * 1e-12

    # ---- end to end through the public API ("e2e") ---------------------------------------------------------------
    for k in range(min(W, 2)):
        hh.solve(*c2_problem(hh, total_paths, args.nsteps, args.precision, base_seed=2000 + k), engine=eng)
    barrier()
    t0 = time.perf_counter()
    e2e_price = None
    for k in range(K):
        sol = hh.solve(*c2_problem(hh, total_paths, args.nsteps, args.precision, base_seed=3000 + k), engine=eng)
        e2e_price = sol.price
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = path_steps_per_step * K / e2e_s
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 16 * len(payoffs) + 0,
           "d2h_bytes_per_step": 24 * len(payoffs),
           "note": "hedgehog_jl_b200.solve(problem, method): scalars + payoff array in, [sum, sumsq, nonfinite] out, "
                   "allreduce of the partial sums across ranks; MonteCarloSolution.ensemble not requested"}
    if world == 1 and not args.skip_ensemble:
        # the reference also returns the terminal price vector (MonteCarloSolution.ensemble): 8 B per trajectory D2H,
        # staged through a pinned double buffer into the caller's pageable array (one untimed call sizes the buffers)
        sol = hh.solve(*c2_problem(hh, total_paths, args.nsteps, args.precision, ensemble=True, base_seed=3999), engine=eng)
        del sol
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(2):
            sol = hh.solve(*c2_problem(hh, total_paths, args.nsteps, args.precision, ensemble=True, base_seed=4000 + k), engine=eng)
            del sol
        torch.cuda.synchronize()
        dt_e = (time.perf_counter() - t0) / 2
        e2e["with_ensemble"] = {"value": path_steps_per_step / dt_e, "unit": UNIT, "d2h_bytes_per_step": 8 * total_paths,
                                "seconds": dt_e}

    # ---- Float32 fast mode of the same workload (config C2 "Float32 fast mode"): reported beside, not as `value` ------
    f32 = None
    if args.precision == "f64" and not args.skip_f32:
        prob32, method32 = c2_problem(hh, total_paths, args.nsteps, "f32")
        mdl32 = _model_of(prob32, method32)

        def sim32(k):
            s = _sim_of(method32, _scheme_of(method32), (rank, world))
            s.base_seed = 7000 + k
            return s
        eng.mc_european_launch(mdl32, sim32(0), payoffs)
        eng.mc_european_collect(D)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(K):
            eng.mc_european_launch(mdl32, sim32(1 + k), payoffs)
        e1.record(stream)
        barrier()
        ms32 = max_over_ranks(e0.elapsed_time(e1))
        r32 = eng.mc_european_collect(D)[0]
        f32 = {"value": path_steps_per_step * K / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32 / K, "price": r32.price,
               "std_error": r32.std_error, "note": "f32 state and normals (32-bit uniforms, MUFU lg2/sin/cos/sqrt), f64 payoff sums"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": per_launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD
                       if args.paths == 100_000_000 and args.nsteps == 252 and args.precision == "f64" else
                       f"Heston EM European call: {args.paths} paths x {args.nsteps} steps per GPU",
                       "paths_per_gpu": args.paths, "n_steps": args.nsteps, "rng": "Philox4x32-10 in-kernel, Box-Muller f64",
                       "l2": "not applicable: the kernel reads no HBM input (state in registers); every step uses a new seed",
                       "sharding": "contiguous global path index blocks per rank, no data-path collective"},
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp64_peak,
                         "traffic": 25088.0,  # dram bytes per launch (ncu, profiles/r1_g_ncu_heston_v3.csv): tables only, no data stream
                         "peak_source": "DFMA-chain microbenchmark run by this bench (hh_bench_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "convention": "algorithmic 25 FLOP per path-step (log/sincos/sqrt expansions NOT counted)",
                         "executed": {
                             "tflops": EXEC_FLOP_PER_PATH_STEP * float(args.paths) * args.nsteps / (per_launch_ms * 1e-3) * 1e-12,
                             "frac_of_fp64_peak": EXEC_FLOP_PER_PATH_STEP * float(args.paths) * args.nsteps
                             / (per_launch_ms * 1e-3) * 1e-12 / fp64_peak,
                             "flop_per_path_step": EXEC_FLOP_PER_PATH_STEP,
                             "source": "ncu op_{dadd,dmul,dfma} counts of this kernel, profiles/r1_c_ncu_heston_fast_v2.csv"},
                         "limiter": "SMSP dispatch port: an FP64 instruction holds it 2-3 cycles (profiles/"
                                    "r1_b_ubench_issue_pipes.txt), so cycles per warp-step ~ 2.2 x 31 FP64 + 1.1 x 60 other; "
                                    "measured 148 (DESIGN.md section 4)"} if args.precision == "f64" else
            {"bound": "fp32+sfu", "achieved": achieved_tflops, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
             "convention": "algorithmic 25 FLOP per path-step; f32 fast mode (MUFU-bound), no FP32 peak measured"},
            "e2e": e2e, "gpu_launches": 2 * K, "clocks": clocks, "f32_fast_mode": f32,
            "check": {"price": last.price, "std_error": last.std_error, "carr_madan": CARR_MADAN_C2,
                      "n_nonfinite": last.n_nonfinite, "e2e_price": e2e_price,
                      "euler_bias_at_252_steps": EULER_BIAS_252,
                      "z_vs_carr_madan_plus_bias": (last.price - CARR_MADAN_C2 - EULER_BIAS_252) / last.std_error
                      if args.nsteps == 252 and last.std_error > 0 else None,
                      "note": "the scheme's O(dt) discretisation bias, measured with 2e8 paths per step count "
                              "(profiles/r1_i_euler_bias_c2.json: +0.0217, +0.0112, +0.0057, +0.0023, +0.0004 at 63..1008 steps)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, info = cpu_sample(100_000, args.nsteps, seconds=args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port",
                                    "sample": f"{info['paths']} paths x {args.nsteps} steps in {info['seconds']:.1f} s, "
                                              "C restatement (oracle/) with OpenMP on all host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
