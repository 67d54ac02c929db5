#!/usr/bin/env python
"""bench.py — LMM forward-rate path-steps/s (BASELINE.json metric) on N B200s, plus the Bermudan swaption wall time.

Workload at N = 1: configs[3] "LIBOR Market Model 40 forward rates, 3-factor covariance, 0.5y steps, 4M paths" (C4), the
configuration the headline metric is quoted on; under torchrun every rank simulates its own contiguous block of
4M paths of ONE logical simulation of N*4M paths (weak scaling, MT19937 jump-ahead, no data-path collective).
One step = {Brownian generation (MT19937 jump-ahead + AS241) + fused log-Euler evolution} of all paths, every realization
X[t][j][path] stored in HBM (getProcessValue semantics).  The C5 Bermudan swaption (1M paths per GPU) is timed separately
and reported under "bermudan".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

N_LIBORS, N_FACTORS, PERIOD = 40, 3, 0.5
SCHEME_NAMES = {0: "EULER", 1: "PREDICTOR_CORRECTOR", 2: "EULER_FUNCTIONAL", 3: "PREDICTOR_CORRECTOR_FUNCTIONAL"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe).  The sampler is started before the
    warm-up (a first nvidia-smi query on a fresh box can take a second), samples every 20 ms with a timestamp, and finish() keeps the
    samples that fall inside the timed windows."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def wait_first_sample(self, timeout=5.0):
        t0 = time.time()
        while not self.samples and time.time() - t0 < timeout:
            time.sleep(0.01)

    def finish(self, windows):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        for ts, s in self.samples:
            if not any(a <= ts <= b for a, b in windows):
                continue
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "window": "device-timed steps + end-to-end steps (20 ms period)"}


def market(pkg):
    from common import lmm_setup
    return lmm_setup(pkg, n_libors=N_LIBORS, n_factors=N_FACTORS, period=PERIOD, dt=PERIOD)


def reference_inputs(orc):
    """The C4 market data for the reference arm, built WITHOUT the product package (numpy + the oracle's time grid only): four-parameter
    volatility sigma_j(t_i) = (a + b tau) exp(-c tau) + d (0 once fixed), correlation exp(-0.1 |T_i - T_j|) reduced to 3 factors like
    LinearAlgebra.factorReduction (eigen-decomposition, largest eigenvalues, first entry positive, row renormalisation, second pass).
    tests/test_cpu_host.py checks that it agrees with the product's own tables."""
    import math
    sim = orc.time_discretization(0.0, int(round(N_LIBORS * PERIOD / PERIOD)), PERIOD)
    tenor = orc.time_discretization(0.0, N_LIBORS, PERIOD)
    T, N = sim.size - 1, tenor.size - 1
    a, b, c, d = 0.2, 0.0, 0.25, 0.3
    sigma = np.zeros((T, N))
    for t in range(T):
        for j in range(N):
            ttm = tenor[j] - sim[t]
            sigma[t, j] = 0.0 if ttm <= 0 else (b * ttm + a) * math.exp(c * (-ttm)) + d

    def factor_matrix(corr, F):
        ev, V = np.linalg.eigh(corr)
        order = np.argsort(-ev, kind="stable")
        fm = np.zeros((corr.shape[0], F))
        for f in range(F):
            v = V[:, order[f]]
            sign = 1.0 if v[0] > 0.0 else -1.0
            fm[:, f] = sign * math.sqrt(max(float(ev[order[f]]), 0.0) / float(np.sum(v * v))) * v
        return fm
    corr = np.array([[math.exp(-0.1 * abs(tenor[r] - tenor[c_])) for c_ in range(N)] for r in range(N)])
    fm = factor_matrix(corr, N_FACTORS)
    for row in range(N):
        fm[row] = fm[row] / math.sqrt(float(np.sum(fm[row] * fm[row])))
    fm = factor_matrix(fm @ fm.T, N_FACTORS)
    return {"sim": sim, "tenor": tenor, "sigma": sigma, "factor_matrix": fm, "L0": np.full(N, 0.05), "T": T}


def run_reference(args):
    """Reference arm: the reference's CPU path (oracle port, all host threads) on a bounded sample of the same workload.  Nothing of the
    product package is on this path: inputs from reference_inputs(), arithmetic in oracle/."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = graft.load_oracle()
    s = reference_inputs(orc)
    cores = os.cpu_count() or 1
    T = s["T"]
    sample = args.cpu_paths
    sec, _ = orc.time_lmm_fused(3141, s["sim"], s["tenor"], N_FACTORS, min(sample, 20000), s["L0"], s["sigma"], s["factor_matrix"], args.scheme, cores)
    inner = max(1, int(4.0 / max(sec * sample / min(sample, 20000), 1e-3)))       # each timed step ~4 s of CPU work
    times = []
    for _ in range(args.warmup + args.steps):
        sec = 0.0
        for r in range(inner):
            dt_, _ = orc.time_lmm_fused(3141 + r, s["sim"], s["tenor"], N_FACTORS, sample, s["L0"], s["sigma"], s["factor_matrix"], args.scheme, cores)
            sec += dt_
        times.append(sec)
    times = times[args.warmup:]
    tot = sum(times)
    value = inner * sample * T * len(times) / tot
    line = {
        "impl": "reference", "metric": "LMM forward-rate path-steps/sec", "value": value, "unit": "path-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C4 LMM 40 forward rates x 3 factors x 40 steps (0.5y), scheme %s; bounded sample of %d x %d paths per step" % (SCHEME_NAMES[args.scheme], inner, sample)},
        "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d x %d paths x %d steps per timed step; oracle C++ restatement of the reference's arithmetic, fused per path, path-parallel over %d threads "
                                   "(faster than the reference's own execution shape; the JVM reference cannot run here: no JVM)" % (inner, sample, T, cores)},
        "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--paths", type=int, default=4_000_000, help="paths per GPU (C4: 4M)")
    ap.add_argument("--bermudan-paths", type=int, default=1_000_000, help="paths per GPU for the C5 Bermudan swaption")
    ap.add_argument("--scheme", type=int, default=2, help="0 EULER, 1 PREDICTOR_CORRECTOR, 2 EULER_FUNCTIONAL (reference default), 3 PC_FUNCTIONAL")
    ap.add_argument("--cpu-paths", type=int, default=0, help="sample size of the CPU baseline (0 = auto)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-bermudan", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="skip the C1/C2/C3 side measurements")
    ap.add_argument("--bermudan-strong-paths", type=int, default=8_000_000, help="total paths of the strong-scaling Bermudan valuation (C5: 8M)")
    args = ap.parse_args()
    if args.cpu_paths == 0:
        args.cpu_paths = 20000 * max(1, (os.cpu_count() or 1))
    if args.impl == "reference":
        return run_reference(args)

    pkg = graft.load_package()
    nv = pkg.native
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    shard = pkg.from_environment()
    nv.init(local_rank)                                      # raises without a GPU: there is no CPU fallback
    import torch
    dist = torch.distributed if world > 1 else None

    def barrier():
        nv.synchronize()
        if dist is not None:
            dist.barrier()
        nv.synchronize()

    s = market(pkg)
    T, N, F = s["T"], s["N"], s["F"]
    P_local = args.paths
    P_global = P_local * world
    factory = pkg.RandomVariableCudaFactory(shard)
    model = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], factory, s["cov"], None, {"measure": "SPOT", "stateSpace": "LOGNORMAL"})
    fixing = [5.0 + 0.5 * i for i in range(10)]
    payment = [5.5 + 0.5 * i for i in range(10)]
    swaption = pkg.Swaption(5.0, fixing, payment, [0.05] * 10)

    def step(seed, price=False):
        bm = pkg.BrownianMotionCuda(s["sim"], F, P_global, seed, factory)
        process = pkg.EulerSchemeFromProcessModel(model, bm, args.scheme)
        process.getProcessValue(T, N - 1)                   # triggers generation + evolution (lazy like the reference)
        if price:
            return swaption.getValue(pkg.LIBORMonteCarloSimulationFromLIBORModel(process))
        return None

    # ---- kernel-level timing: device events on the library's stream, K steps back to back ------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for w in range(args.warmup):
        step(1000 + w)
    sampler.wait_first_sample()
    barrier()
    launches0 = nv.launch_count()
    win_a = time.time()
    nv.timer_start()
    for k in range(args.steps):
        step(3141 + k)
    ms = nv.timer_stop_ms()
    win_b = time.time()
    launches = nv.launch_count() - launches0
    barrier()
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=shard.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = P_global * T * args.steps / (ms * 1e-3)

    # ---- end to end through the host API: host tables in, price out, per step ---------------------------------------------
    for w in range(min(2, args.warmup)):
        step(2000 + w, price=True)                           # warm-up of the priced path (pool blocks of the product's temporaries)
    barrier()
    win_c = time.time()
    t0 = time.perf_counter()
    prices = [step(3141 + k, price=True) for k in range(args.steps)]
    nv.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.finish([(win_a, win_b), (win_c, time.time())])
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=shard.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = P_global * T * args.steps / e2e_s
    h2d = 8 * (T * F + T + T * N * F + T * N + 4 * N) + 4 * T + 8 * (T * F + (T + 1) * N) + 624 * 4
    d2h = 2 * 8 * 148 * 4 * 6                               # double-double partials of the reductions behind the price

    # ---- per-kernel durations for the roofline (separate, instrumented steps) ----------------------------------------------
    bm_ms, eu_ms = [], []
    for k in range(min(3, args.steps)):
        bm = pkg.BrownianMotionCuda(s["sim"], F, P_global, 7000 + k, factory)
        nv.synchronize()
        nv.timer_start()
        bm.getBrownianIncrement(0, 0)
        bm_ms.append(nv.timer_stop_ms())
        process = pkg.EulerSchemeFromProcessModel(model, bm, args.scheme)
        nv.timer_start()
        process.getProcessValue(T, N - 1)
        eu_ms.append(nv.timer_stop_ms())
        del process, bm
    live = sum(max(0, N - (t + 1)) for t in range(T))
    euler_bytes = P_local * (8 * live + 8 * F * sum(1 for t in range(T) if N - (t + 1) > 0))
    bm_bytes = P_local * T * F * 8
    peak, peak_src = measured_peaks()
    eu = float(np.mean(eu_ms))
    achieved = euler_bytes / (eu * 1e-3) / 1e9
    import ctypes as C
    tf = C.c_double()
    dfma_sampler = ClockSampler(local_rank)
    dfma_sampler.start()
    dfma_sampler.wait_first_sample()
    t_d0 = time.time()
    for _ in range(3):
        nv.check(nv.load().fmb_bench_dfma_tflops(C.byref(tf)))
    dfma_clocks = dfma_sampler.finish([(t_d0, time.time())])
    # DRAM traffic per launch from the committed ncu --set full capture (profiles/r01_top_kernels_ncu.json), valid for the 4 M-path launch
    traffic = None
    try:
        if P_local == 4_000_000 and args.scheme == 2:
            for k in json.load(open(os.path.join(ROOT, "profiles", "r02_top_kernels_ncu.json"))):
                if "eulerLmmKernel" in k["Kernel Name"] and float(k["gpu__time_duration.sum"].split()[0]) > 10.0:      # the 4 M-path launch
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                    traffic = sum(float(k[m].split()[0]) * scale[k[m].split()[1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        traffic = None
    # FP64 work of the kernel: 56 FP64 instructions (DFMA/DADD/DMUL/DSETP) per live rate-step on the kernel's hot path, counted from the
    # executed-instruction column of the ncu source page (profiles/r01_notes.md, profiles/tools/hot_path.py); DFMA-equivalent flops = 2 each
    fp64_instr = 56.0 * live * P_local
    fp64_achieved_tflops = 2.0 * fp64_instr / (eu * 1e-3) / 1e12
    # The dominant kernel is FP64-pipe bound (double log + exp + division per rate-step; ncu: sm__pipe_fp64_cycles_active 68 %, DRAM 20 %), so
    # the binding roofline is the FP64 pipe: DFMA-equivalent flops of the kernel's hot path / the DFMA peak measured in this run.  The
    # HBM fraction (algorithmic bytes / measured copy bandwidth) is reported next to it in the same object.
    roofline = {"bound": "fp64", "kernel": "eulerLmmKernel<3,1,0,1,0>", "achieved": fp64_achieved_tflops, "peak": tf.value,
                "unit": "TFLOP/s", "frac": fp64_achieved_tflops / tf.value,
                "peak_source": "fmb_bench_dfma_tflops measured in this run (8 independent DFMA chains per thread); MEASURED_PEAKS.json has no FP64 entry",
                "fp64_instructions_per_rate_step": 56, "avg_launch_ms": eu,
                "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": euler_bytes},
                "traffic": traffic, "traffic_source": "static: profiles/r02_top_kernels_ncu.json (ncu --set full capture of this kernel at this size, not re-measured in this run)" if traffic else None,
                "dfma_clocks": dfma_clocks}
    fp64 = {"euler_ms": eu, "brownian_ms": float(np.mean(bm_ms)), "brownian_achieved_gbs": bm_bytes / (float(np.mean(bm_ms)) * 1e-3) / 1e9,
            "brownian_hbm_frac": bm_bytes / (float(np.mean(bm_ms)) * 1e-3) / 1e9 / peak}

    # ---- optional FAST floating-point mode (not the headline: the headline is STRICT), same step, same sizes ---------------------------
    fast = None
    try:
        nv.set_fp_mode(1)
        step(900)
        nv.synchronize()
        nv.timer_start()
        for k in range(args.steps):
            step(4141 + k)
        fast_ms = nv.timer_stop_ms() / args.steps
        fast = {"ms_per_step": fast_ms, "value": P_global * T / (fast_ms * 1e-3),
                "what": "fmb_set_fp_mode(1): FMA contraction, functional schemes carry the log-state; paths within 1e-12 of the oracle (tests)"}
    finally:
        nv.set_fp_mode(0)

    # ---- C5: Bermudan swaption wall time (simulation + backward induction with regression, price on the host) ------------------
    # weak: bermudan_paths per GPU (1 M: the north star's 8 M paths on 8 GPUs); strong: the SAME 8 M paths on however many GPUs run.
    bermudan = None
    parity_inputs = None
    if not args.skip_bermudan:
        from common import bermudan_spec
        b = bermudan_spec(s)
        product = pkg.BermudanSwaption(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])

        def bermudan_run(paths_total, fac, reps=3, with_swaption=False):
            mdl = model if fac is factory else pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], fac, s["cov"], None,
                                                                                      {"measure": "SPOT", "stateSpace": "LOGNORMAL"})
            walls, price, sw, launches_b = [], None, None, 0
            for rep in range(reps):
                if fac is factory:
                    barrier()
                else:
                    nv.synchronize()
                l0 = nv.launch_count()
                t0 = time.perf_counter()
                bm = pkg.BrownianMotionCuda(s["sim"], F, paths_total, 3141, fac)
                sim = pkg.LIBORMonteCarloSimulationFromLIBORModel(pkg.EulerSchemeFromProcessModel(mdl, bm, args.scheme))
                price = product.getValue(sim)
                nv.synchronize()
                walls.append(time.perf_counter() - t0)
                launches_b = nv.launch_count() - l0
                if with_swaption and rep == reps - 1:
                    sw = swaption.getValue(sim)
                del sim, bm
            wall = min(walls[1:]) if len(walls) > 1 else walls[0]
            if dist is not None and fac is factory:
                t = torch.tensor([wall], dtype=torch.float64, device=shard.device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                wall = float(t.item())
            return {"wall_ms": 1e3 * wall, "paths": paths_total, "price": price, "launches": int(launches_b)}, sw

        Pb = args.bermudan_paths * world
        weak, sw_sharded = bermudan_run(Pb, factory, with_swaption=True)
        strong, _ = bermudan_run(args.bermudan_strong_paths, factory)
        nv.load().fmb_pool_trim()
        bermudan = dict(weak, exercise_dates=20, basis_functions=6, scaling="weak", paths_per_gpu=args.bermudan_paths,
                        exchange=(("native: peer-memory stores over NVLink + flags, one kernel per exchange on the compute stream" if getattr(shard, "peer_exchange", False)
                                   else "native NCCL all-gather on the compute stream") if shard.native_comm else ("shared-memory mailbox" if shard._mailbox is not None else "torch.distributed"))
                        if world > 1 else "none (1 GPU)",
                        strong={"wall_ms": strong["wall_ms"], "paths": strong["paths"], "price": strong["price"], "scaling": "strong"})
        parity_inputs = (Pb, weak["price"], sw_sharded, product, bermudan_run)

    # ---- CPU baseline on the host cores (rank 0, N = 1 only; bounded sample) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        orc = graft.load_oracle()
        cores = os.cpu_count() or 1
        sec, reps = 0.0, 0
        while sec < 10.0 and reps < 200:                      # bounded sample: ~10 s of CPU work
            dt_, _ = orc.time_lmm_fused(3141 + reps, s["sim"].times, s["tenor"].times, F, args.cpu_paths, s["L0"], s["sigma"], s["factor_matrix"], args.scheme, cores)
            sec += dt_
            reps += 1
        shaped_paths = 20000
        sec_shaped, _ = orc.time_lmm_reference_shaped(3141, s["sim"].times, s["tenor"].times, F, shaped_paths, s["L0"], s["sigma"], s["factor_matrix"], args.scheme)
        cpu = {"value": reps * args.cpu_paths * T / sec, "unit": "path-steps/s", "cores": cores, "kind": "port",
               "sample": "%d x %d paths x %d steps, oracle port fused per path, %d threads, %.1f s" % (reps, args.cpu_paths, T, cores, sec),
               "reference_shaped": {"value": shaped_paths * T / sec_shaped, "cores": 1,
                                    "sample": "%d paths, one array pass + allocation per RandomVariable op, single sequential MT stream, %.1f s" % (shaped_paths, sec_shaped)}}

    # ---- multi-rank parity (N > 1): the sharded prices against rank 0 recomputing the same logical simulation on ONE GPU, and a 20 000-path
    #      sharded valuation against the CPU oracle, through the native exchange and through the host-side exchange -------------------
    shard_parity = None
    if world > 1 and parity_inputs is not None:
        Pb, price_sharded, sw_sharded, product, bermudan_run = parity_inputs

        def small_sharded(fac):
            mdl = pkg.LIBORMarketModelFromCovarianceModel.of(s["tenor"], None, s["L0"], s["df"], fac, s["cov"], None, {"measure": "SPOT", "stateSpace": "LOGNORMAL"})
            bm = pkg.BrownianMotionCuda(s["sim"], F, 20_000, 3141, fac)
            sim = pkg.LIBORMonteCarloSimulationFromLIBORModel(pkg.EulerSchemeFromProcessModel(mdl, bm, args.scheme))
            return product.getValue(sim), swaption.getValue(sim)

        p20_native = small_sharded(factory)
        exchanges = None
        if shard.native_comm:
            import ctypes as C
            ex = C.c_uint64()
            nv.check(nv.load().fmb_comm_info(None, None, C.byref(ex)))
            exchanges = ex.value
            barrier()
            nv.check(nv.load().fmb_comm_shutdown())              # from here on the library reduces locally; shards are merged on the host
            shard.native_comm = False
        p20_host = small_sharded(pkg.RandomVariableCudaFactory(shard))           # torch.distributed (or mailbox) exchange of the same partials
        barrier()
        if rank == 0:
            local = pkg.RandomVariableCudaFactory(pkg.LOCAL)
            single, sw_single = bermudan_run(Pb, local, reps=2, with_swaption=True)
            orc = graft.load_oracle()
            from common import lmm_oracle
            ref = lmm_oracle(orc, s, 20_000, scheme=args.scheme)
            r = ref.bermudan(b["is_exercise"], b["fixing"], b["lengths"], b["payment"], b["notionals"], b["swaprates"])
            ref_sw, _, _ = ref.swaption(5.0, fixing, payment, [0.05] * 10)
            rel = lambda x, y: abs(x - y) / abs(y)
            shard_parity = {
                "paths": Pb, "bermudan_sharded": price_sharded, "bermudan_single_gpu": single["price"],
                "bermudan_rel_vs_single_gpu": rel(price_sharded, single["price"]),
                "swaption_rel_vs_single_gpu": rel(sw_sharded, sw_single),
                "rel_vs_oracle_20k_paths": {"bermudan_native_exchange": rel(p20_native[0], r["price"]), "swaption_native_exchange": rel(p20_native[1], ref_sw),
                                            "bermudan_host_exchange": rel(p20_host[0], r["price"]), "swaption_host_exchange": rel(p20_host[1], ref_sw)},
                "single_gpu_same_paths_wall_ms": single["wall_ms"], "sharded_wall_ms": bermudan["wall_ms"],
                "speedup_vs_single_gpu_same_paths": single["wall_ms"] / bermudan["wall_ms"],
                "native_exchanges_total": exchanges, "tolerance": 1e-10,
                "ok": bool(max(rel(price_sharded, single["price"]), rel(sw_sharded, sw_single), rel(p20_native[0], r["price"]), rel(p20_native[1], ref_sw),
                               rel(p20_host[0], r["price"]), rel(p20_host[1], ref_sw)) <= 1e-10)}
        barrier()

    # ---- the other BASELINE.json configurations (N = 1 only; 1 warm-up + 2 timed repetitions each, device events) -----------------------
    configs = None
    if world == 1 and not args.skip_configs:
        sys.path.insert(0, os.path.join(ROOT, "profiles"))
        import bench_configs as bc
        configs = bc.run_all(pkg)
        if bermudan is not None:
            configs.append({"config": "C5 LMM Bermudan swaption, %d paths (simulate + 20 exercise dates x 6 basis functions, price on the host)" % bermudan["paths"],
                            "paths": bermudan["paths"], "steps": T, "ms": bermudan["wall_ms"], "path_steps_per_s": bermudan["paths"] * T / (bermudan["wall_ms"] * 1e-3)})
            configs.append({"config": "C5 LMM Bermudan swaption, %d paths on this GPU count" % bermudan["strong"]["paths"], "paths": bermudan["strong"]["paths"], "steps": T,
                            "ms": bermudan["strong"]["wall_ms"], "path_steps_per_s": bermudan["strong"]["paths"] * T / (bermudan["strong"]["wall_ms"] * 1e-3)})

    # ---- SURVEY §8f rank 4 on the device: a calibration loop over the fused simulation and an AAD sweep (N = 1 only) --------------------
    f4 = None
    if world == 1 and not args.skip_configs:
        f4 = bc.next_rows(pkg)

    if rank == 0:
        line = {
            "metric": "LMM forward-rate path-steps/sec", "value": value, "unit": "path-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4 LMM 40 forward rates x 3 factors x 40 steps (0.5y), %d paths per GPU, scheme %s, spot measure, log-normal, STRICT fp (no FMA contraction)"
                                   % (P_local, SCHEME_NAMES[args.scheme]),
                       "paths_total": P_global, "l2": "inputs/outputs far larger than L2: %.1f GB written per step per GPU" % ((euler_bytes + bm_bytes) / 1e9),
                       "step": "Brownian generation (MT19937 jump-ahead + AS241) + fused Euler evolution, all X[t][j][path] stored"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "path-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "BrownianMotionCuda + EulerSchemeFromProcessModel + Swaption.getValue through the host API, price on the host each step",
                    "price": prices[-1]},
            "roofline": roofline, "kernels": fp64, "fast_mode": fast, "bermudan": bermudan, "shard_parity": shard_parity, "configs": configs, "f4": f4,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
