#!/usr/bin/env python
"""bench.py -- headline measurement of the LatticeMC hot path on B200 (contract: see the task's bench.py section).

Metric (BASELINE.json): KMC vacancy hops/s (primary, walker-sharded, weak scaling) and CMC swap trials/s
(secondary object "cmc") per box, next to the reference's own CPU path timed on this box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--gpus N] ...                the reference's CPU implementation (oracle/_ref)

One "step" = one pass of the hot path over one batch: every walker of the rank advances `--hops` KMC steps
(12 candidate barriers + Arrhenius rates + event select + residence time + jump per hop) in ONE kernel launch.
Workload = BASELINE configs[2] as written: 8192 independent single-vacancy Al-2%Mg-2%Zn walkers on 8x8x8 FCC cells
(2048 sites), T_w = 400..600 K, synthetic JSON coefficients, walker w on GPU w mod N -- STRONG scaling
(the headline at N > 1); the weak-scaling line (8192 walkers on every GPU) is reported next to it as `weak_scaling`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_KMC_STEP = 3792 + 18   # 12 events x 316 B + state update
BYTES_PER_EVENT = 316            # 60 B occupancy + 240 B neighbour indices + 16 B output
BYTES_PER_TRIAL = 440            # CMC/SA swap trial
FACTOR = 8                       # 8x8x8 cells = 2048 sites per walker
P_MG = P_ZN = 0.02


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel, **shape):
    """DRAM bytes per launch of `kernel` from the committed ncu capture with the same launch shape, else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            for row in json.load(f).get(kernel, []):
                if all(row.get(k) == v for k, v in shape.items()):
                    return row.get("dram_bytes")
    except (OSError, ValueError):
        pass
    return None


def walker_occupancy(first_walker, n_walkers, factor=FACTOR, stride=1, p_mg=P_MG, p_zn=P_ZN):
    """Random Al-Mg-Zn alloy per walker (seed 42 + global walker index first + stride * i), one vacancy each, REASSIGNED id order."""
    n = 4 * factor ** 3
    out = np.empty((n_walkers, n), dtype=np.uint8)
    for w in range(n_walkers):
        u = np.random.default_rng(42 + first_walker + stride * w).random(n)
        row = out[w]
        row[:] = 1
        row[u < p_mg + p_zn] = 3
        row[u < p_mg] = 2
        row[n // 2 + 3] = 0
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 8 and r[4 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    """One reference trajectory: mc::KineticMcFirstOmp::Simulate (LRU predictor, as shipped) on one walker, 1 OMP thread."""
    seed, steps, json_path = args
    from oracle import ref_lib as R
    occ = walker_occupancy(seed, 1)[0]
    # walker_occupancy is in REASSIGNED order; the reference config is built in GenerateFCC order then reassigned
    from latticemontecarlo_b200 import synth
    perm = synth.generate_to_reassigned_permutation(FACTOR)      # perm[new] = old
    occ_gen = np.empty_like(occ)
    occ_gen[perm] = occ
    cfg = R.RefConfig.fcc(FACTOR, occ_gen, reassign=True)
    t0 = time.perf_counter()
    out = R.kmc_first_omp(cfg, json_path, temperature=400.0 + 200.0 * (seed % 8192) / 8191.0, maximum_steps=steps - 1,
                          seed=seed + 1, threads=1, trace=False)
    return out["seconds"], time.perf_counter() - t0


def reference_kmc_throughput(json_path, steps_per_walker, n_procs, first_seed=0):
    """Aggregate hops/s of n_procs independent reference trajectories running concurrently (one per host core)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_ref_worker, [(first_seed + p, steps_per_walker, json_path) for p in range(n_procs)])
    wall = time.perf_counter() - t0
    simulate_s = max(r[0] for r in res)          # slowest trajectory's time inside Simulate() (constructors excluded)
    return n_procs * steps_per_walker / simulate_s, simulate_s, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref_lib as R
    from latticemontecarlo_b200 import synth
    if not R.build():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/liblmc_ref.so missing and /root/reference absent"}))
        return 0
    R.lib()                                   # mapped in this process too (the trajectories run in forked workers that inherit it)
    cores = host_cores()
    with tempfile.TemporaryDirectory() as d:
        js = os.path.join(d, "quartic_coefficients.json")
        synth.write_synthetic_json(js)
        steps_per_walker = args.ref_hops
        for _ in range(args.warmup if args.warmup < 2 else 1):     # one short warm-up pass (page-in, LRU is per process anyway)
            reference_kmc_throughput(js, max(200, steps_per_walker // 10), cores)
        rates, times = [], []
        for k in range(args.steps):
            rate, sim_s, _ = reference_kmc_throughput(js, steps_per_walker, cores, first_seed=1000 * k)
            rates.append(rate)
            times.append(sim_s)
    value = sum(cores * steps_per_walker for _ in rates) / sum(times)
    sample = "%d concurrent mc::KineticMcFirstOmp trajectories (LRU predictor as shipped, 1 OMP thread each) x %d hops, 8x8x8 cell" % (
        cores, steps_per_walker)
    line = {
        "impl": "reference", "metric": "kmc_vacancy_hops_per_s", "value": value, "unit": "hops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "hops/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "hops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def kmc_kernel_name(lanes, resident_occupancy=False, handoff=False):
    """The first-order kernel the library chose: thousands of walkers -> half-warp per walker (with the tail of the launch
    handed to the block-per-walker kernel); a few per SM -> block per walker (small cells: with the walker's occupancy
    resident in shared memory)."""
    if lanes == 0:
        return "kmc_run_kernel + kmc_team_run_kernel<8> for the tail of the launch" if handoff else "kmc_run_kernel"
    return "kmc_team_run_kernel<%d%s>" % (lanes, ", occupancy in shared memory" if resident_occupancy else "")


def workload_config(args, n_gpus):
    return {"workload": "BASELINE configs[2]: batched KMC, %d independent single-vacancy Al-2%%Mg-2%%Zn walkers on 8x8x8 FCC "
                        "(2048 sites), T=400..600 K, walker w on GPU w mod %d, synthetic JSON coefficients (K=24/32)" % (args.walkers, n_gpus),
            "walkers_total": args.walkers, "walkers_per_gpu": -(-args.walkers // n_gpus), "hops_per_walker_per_step": args.hops,
            "sites_per_walker": 4 * FACTOR ** 3, "l2": "flushed between timed steps (256 MiB write)",
            "parallelism": "walker-sharded x%d (w mod N), no data-path collective" % n_gpus}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    from latticemontecarlo_b200 import build as _build, capi, sharding, synth
    _build.build()
    if capi.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    n_gpus = world
    H = args.hops
    n_sites = 4 * FACTOR ** 3
    tmp = tempfile.TemporaryDirectory()
    js = os.path.join(tmp.name, "quartic_coefficients.json")
    synth.write_synthetic_json(js)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def measure_kmc(first_walker, stride, W, walkers_total, steps, p_mg=P_MG, p_zn=P_ZN, with_e2e=True, sample_clocks=False):
        """KMC throughput of this rank's W walkers (global indices first_walker + stride * i): resident (CUDA events on the
        engine stream, L2 flushed between steps) and end to end through the C ABI with pinned host buffers."""
        engine = capi.Engine(FACTOR, id_order=capi.ORDER_REASSIGNED, n_walkers=W, device=local_rank)
        engine.load_coefficients(js)
        occ_pinned = torch.empty((W, n_sites), dtype=torch.uint8, pin_memory=True)
        occ_np = occ_pinned.numpy()
        occ_np[:] = walker_occupancy(first_walker, W, stride=stride, p_mg=p_mg, p_zn=p_zn)
        out_pinned = torch.empty((W, n_sites), dtype=torch.uint8, pin_memory=True)
        temps = 400.0 + 200.0 * (first_walker + stride * np.arange(W)) / max(1, walkers_total - 1)
        stream = torch.cuda.ExternalStream(engine.cuda_stream(), device=local_rank)
        engine.set_occupancy_all(occ_np)
        engine.kmc_reset()

        def step_resident():
            flush.fill_(1)                       # evict the walkers' occupancy from L2 between timed steps
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                engine.kmc_run(H, temperatures=temps, seed=20260101)
                e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1), engine.last_kernel_ms()

        for _ in range(max(3, args.warmup)):
            step_resident()
        steps_before = engine.kmc_state()["steps"].copy()
        launches0 = engine.launch_count()
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        wall0 = time.perf_counter()
        timings = [step_resident() for _ in range(steps)]
        barrier()
        wall = time.perf_counter() - wall0
        if sampler:
            sampler.__exit__(None, None, None)
        launches = engine.launch_count() - launches0
        st = engine.kmc_state()
        advanced = st["steps"] - steps_before
        step_ms, kernel_ms = sharding.max_over_ranks([sum(t[0] for t in timings), sum(t[1] for t in timings)], device="cuda")
        res = {"engine": engine, "occ_pinned": occ_pinned, "temps": temps, "step_ms": step_ms, "kernel_ms": kernel_ms,
               "local_kernel_ms": sum(t[1] for t in timings), "wall": wall, "launches": launches, "clocks": sampler,
               "walkers_advanced_all_steps": int(np.count_nonzero(advanced == H * steps)), "min_steps_advanced": int(advanced.min()),
               "age_hops": [int(steps_before.min()), int(st["steps"].max())], "lanes": engine.kmc_last_launch_lanes(),
               "resident_occupancy": engine.kmc_last_launch_resident_occupancy(), "handoff": engine.kmc_last_launch_handoff()}
        if with_e2e:
            # ---- end to end through the C ABI with host buffers: H2D occupancy, reset, run, D2H state + occupancy
            def step_e2e():
                t0 = time.perf_counter()
                engine.set_occupancy_all(occ_np)
                engine.kmc_reset()
                engine.kmc_run(H, temperatures=temps, seed=20260101)
                engine.kmc_state()
                capi._check(capi.lib().lmc_engine_get_occupancy_all(engine.h, out_pinned.numpy().ctypes.data_as(capi.C.c_void_p),
                                                                    capi.C.c_int64(W * n_sites)))
                return time.perf_counter() - t0

            for _ in range(2):
                step_e2e()
            barrier()
            e2e_s = sum(step_e2e() for _ in range(steps))
            barrier()
            res["e2e_s"] = sharding.max_over_ranks([e2e_s], device="cuda")[0]
        return res

    # ---- headline: BASELINE configs[2] as written -- args.walkers walkers in total, walker w on GPU w mod N (strong scaling)
    W_total = args.walkers
    W = len(sharding.interleaved_walkers(W_total, rank, n_gpus))          # walker w on GPU w mod N
    head = measure_kmc(rank, n_gpus, W, W_total, args.steps, sample_clocks=True)
    engine, occ_pinned, temps = head["engine"], head["occ_pinned"], head["temps"]
    hops_total = float(W_total) * H * args.steps
    value = hops_total / (head["step_ms"] * 1e-3)
    e2e_value = hops_total / head["e2e_s"]
    advanced_all = sharding.sum_over_ranks([head["walkers_advanced_all_steps"]], device="cuda")[0]
    peak, peak_kind = measured_peaks()
    achieved = (float(W) * H * args.steps * BYTES_PER_KMC_STEP) / (head["local_kernel_ms"] * 1e-3) / 1e9     # this GPU, dominant kernel
    line = {
        "metric": "kmc_vacancy_hops_per_s", "value": value, "unit": "hops/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": head["step_ms"] / args.steps, "higher_is_better": True,
        "scaling": "strong" if n_gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, n_gpus),
        "e2e": {"value": e2e_value, "unit": "hops/s", "h2d_bytes_per_step": int(W * n_sites + W * 8),
                "d2h_bytes_per_step": int(W * n_sites + W * 44)},
        "gpu_launches": int(head["launches"]),
        "walkers_advanced_all_steps": int(advanced_all), "walkers_total": W_total,
        "walker_age_hops_during_timed_region": head["age_hops"],
        "roofline": {"kernel": kmc_kernel_name(head["lanes"], head["resident_occupancy"], head["handoff"]), "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": recorded_traffic("kmc_run_kernel", walkers=W, hops=H, handoff=bool(head["handoff"])) if head["lanes"] == 0 else recorded_traffic("kmc_team_run_kernel", walkers=W, hops=H, lanes=head["lanes"]), "peak_kind": peak_kind,
                     "algorithmic_bytes_per_launch": int(W * H * BYTES_PER_KMC_STEP),
                     "note": "effective bandwidth: 3810 algorithmic B per KMC step (SURVEY 8(d)); geometry comes from constant "
                             "offset tables and walkers are L2-resident, so physical DRAM traffic is far below this"
                             + ("; the launch duration covers the throughput kernel, the latency kernel that finishes the walkers still "
                                "running when all but ~19 per SM were through, and the two one-block helper kernels between them" if head["handoff"] else "")},
        "wall_s_timed_region": head["wall"],
    }
    clocks = head["clocks"]
    if n_gpus > 1:
        # weak-scaling extra: args.walkers walkers on EVERY GPU (the round-1 headline), same w mod N dealing of seeds / temperatures
        weak = measure_kmc(rank, n_gpus, W_total, W_total * n_gpus, max(3, args.steps // 2), with_e2e=False)
        wsteps = max(3, args.steps // 2)
        line["weak_scaling"] = {"value": float(W_total) * n_gpus * H * wsteps / (weak["step_ms"] * 1e-3), "unit": "hops/s", "scaling": "weak",
                                "walkers_per_gpu": W_total, "walkers_total": W_total * n_gpus, "ms_per_step": weak["step_ms"] / wsteps,
                                "roofline_frac_this_gpu": float(W_total) * H * wsteps * BYTES_PER_KMC_STEP / (weak["local_kernel_ms"] * 1e-3) / 1e9 / peak}
        weak["engine"].close()
    if rank == 0:
        line["barrier_eval"] = bench_barrier_eval(engine, torch, W, peak)          # on the walkers as the timed region left them
    if rank == 0 and not args.no_age:
        line["value_vs_age"] = bench_value_vs_age(engine, occ_pinned, W, H, temps)
        rich = measure_kmc(0, 1, W_total if n_gpus == 1 else W, W_total, 3, p_mg=0.10, p_zn=0.10, with_e2e=False) if n_gpus == 1 else None
        if rich:
            wr = W_total
            line["alloy_10Mg_10Zn"] = {"value": float(wr) * H * 3 / (rich["step_ms"] * 1e-3), "unit": "hops/s", "walkers": wr,
                                       "walkers_advanced_all_steps": rich["walkers_advanced_all_steps"],
                                       "note": "same driver on Al-10%Mg-10%Zn walkers: the table walk is O(non-solvent sites + their pairs)"}
            rich["engine"].close()
    if rank == 0 and n_gpus == 1 and not args.no_age:
        line["low_occupancy"] = bench_low_occupancy(local_rank, js, max(1, W_total // 8), H)
    cmc_multi = None
    if world > 1 and not args.no_cmc:
        cmc_multi = bench_cmc_multi_gpu(torch, dist, local_rank, rank, world, js)      # collective: every rank takes part
    if rank == 0:
        line["clocks"] = clocks.summary()
        if n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(js)
        if not args.no_chain:
            line["single_trajectory"] = bench_single_trajectory(local_rank, js)
            line["chain"] = bench_chain(engine, W, temps, occ_pinned, peak, js, n_gpus == 1 and not args.no_cpu_baseline)
        if not args.no_cmc:
            line["cmc"] = bench_cmc(torch, local_rank, js, peak, n_gpus == 1 and not args.no_cpu_baseline)
            if cmc_multi is not None:
                line["cmc"]["multi_gpu"] = cmc_multi
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bench_value_vs_age(engine, occ_pinned, W, H, temps):
    """The rate as the walkers age (trajectories find solute: longer table walks): kernel rate of one launch of H hops at
    walker ages of about 0, 20k and 200k hops."""
    out = []
    engine.set_occupancy_all(occ_pinned.numpy())
    engine.kmc_reset()
    target, done = [0, 20000, 200000], 0
    for age in target:
        while done < age:
            n = min(8192, age - done)
            engine.kmc_run(n, temperatures=temps, seed=20260101)
            done += n
        engine.kmc_run(H, temperatures=temps, seed=20260101)
        ms = engine.last_kernel_ms()
        done += H
        out.append({"age_hops": age, "value": W * H / (ms * 1e-3), "kernel_ms": ms})
    return out


def bench_barrier_eval(engine, torch, W, peak):
    """Batched barrier evaluation with resident inputs: all 12 candidate events of every walker's vacancy."""
    st = engine.kmc_state()
    vac = st["vacancy"]
    nbrs = np.stack([engine.neighbors(1, int(v)) for v in vac[:256]])
    # neighbour ids for all walkers: translation by the vacancy differs per walker, so compute on the host once
    all_nbrs = np.empty((W, 12), dtype=np.int64)
    all_nbrs[:256] = nbrs
    for w in range(256, W):
        all_nbrs[w] = engine.neighbors(1, int(vac[w]))
    n = W * 12
    dev = torch.device("cuda")
    d_w = torch.from_numpy(np.repeat(np.arange(W, dtype=np.int32), 12)).to(dev)
    d_i = torch.from_numpy(np.repeat(vac, 12)).to(dev)
    d_j = torch.from_numpy(all_nbrs.reshape(-1)).to(dev)
    d_ea = torch.empty(n, dtype=torch.float64, device=dev)
    d_de = torch.empty(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    times = []
    for k in range(8):
        engine.eval_barriers_dev(n, d_w.data_ptr(), d_i.data_ptr(), d_j.data_ptr(), d_ea.data_ptr(), d_de.data_ptr())
        ms = engine.last_kernel_ms()
        if k >= 3:
            times.append(ms)
    ms = sum(times) / len(times)
    achieved = n * BYTES_PER_EVENT / (ms * 1e-3) / 1e9
    out = {"kernel": "barrier_kernel", "events": n, "ms": ms, "events_per_s": n / (ms * 1e-3), "achieved_gbs": achieved,
           "frac_of_hbm_peak": achieved / peak, "bytes_per_event": BYTES_PER_EVENT, "finite": bool(torch.isfinite(d_ea).all().item()),
           "note": "arbitrary (vacancy, neighbour) events, one thread per event"}
    # the same kernel on a batch that fills the device several times over (98k threads are a third of one wave)
    big = 16
    d_wb, d_ib, d_jb = d_w.repeat(big), d_i.repeat(big), d_j.repeat(big)
    d_eab = torch.empty(n * big, dtype=torch.float64, device=dev)
    d_deb = torch.empty(n * big, dtype=torch.float64, device=dev)
    times = []
    for k in range(8):
        engine.eval_barriers_dev(n * big, d_wb.data_ptr(), d_ib.data_ptr(), d_jb.data_ptr(), d_eab.data_ptr(), d_deb.data_ptr())
        if k >= 3:
            times.append(engine.last_kernel_ms())
    msb = sum(times) / len(times)
    out["large_batch"] = {"events": n * big, "ms": msb, "events_per_s": n * big / (msb * 1e-3),
                          "frac_of_hbm_peak": n * big * BYTES_PER_EVENT / (msb * 1e-3) / 1e9 / peak,
                          "matches_small_batch": bool(torch.equal(d_eab[:n], d_ea) and torch.equal(d_deb[-n:], d_de)),
                          "note": "the same events 16 times over, still grouped by vacancy (12 consecutive threads share one "
                                  "neighbourhood: L1 hits); events in arbitrary order run at 0.27-0.29 (tools/barrier_probe.py)"}
    del d_wb, d_ib, d_jb, d_eab, d_deb
    # the event-list form (KineticMcFirstOmp::BuildEventList for a batch of vacancies): one box scan per vacancy.  Every
    # walker holds one vacancy, so a large batch is built by listing each (walker, vacancy) item `reps` times.
    reps = 16
    d_v = torch.from_numpy(np.tile(vac, reps)).to(dev)
    d_wv = torch.from_numpy(np.tile(np.arange(W, dtype=np.int32), reps)).to(dev)
    d_nb = torch.empty(W * reps * 12, dtype=torch.int64, device=dev)
    d_ea2 = torch.empty(W * reps * 12, dtype=torch.float64, device=dev)
    d_de2 = torch.empty(W * reps * 12, dtype=torch.float64, device=dev)
    times = []
    for k in range(8):
        engine.eval_vacancy_events_dev(W * reps, d_wv.data_ptr(), d_v.data_ptr(), d_nb.data_ptr(), d_ea2.data_ptr(), d_de2.data_ptr())
        if k >= 3:
            times.append(engine.last_kernel_ms())
    ms2 = sum(times) / len(times)
    n2 = W * reps * 12
    # same events, same numbers (the list form is ordered by neighbour id: compare as sorted sets per vacancy)
    same = bool(torch.allclose(torch.sort(d_ea2[:n].view(W, 12), dim=1).values, torch.sort(d_ea.view(W, 12), dim=1).values, rtol=0, atol=1e-12))
    achieved2 = n2 * BYTES_PER_EVENT / (ms2 * 1e-3) / 1e9
    out["event_lists"] = {"kernel": "vacancy_events_kernel", "events": n2, "ms": ms2, "events_per_s": n2 / (ms2 * 1e-3),
                          "achieved_gbs": achieved2, "frac_of_hbm_peak": achieved2 / peak, "bytes_per_event": BYTES_PER_EVENT,
                          "matches_barrier_kernel": same,
                          "note": "12 events of each of %d (walker, vacancy) items per launch, half-warp per item, one box scan" % (W * reps)}
    return out


def bench_low_occupancy(device, json_path, walkers, hops):
    """One GPU's share of the job when it is spread over 8 GPUs (1024 of 8192 walkers: 7 per SM).  A KMC step is a serial
    dependency, so this is a latency measurement: the half-warp kernel against the block-per-walker kernel the library
    picks at this occupancy (LMC_KMC_TEAM_LANES is read at every launch)."""
    from latticemontecarlo_b200 import capi
    eng = capi.Engine(FACTOR, id_order=capi.ORDER_REASSIGNED, n_walkers=walkers, device=device)
    eng.load_coefficients(json_path)
    occ = walker_occupancy(0, walkers, stride=8)
    temps = 400.0 + 200.0 * (8 * np.arange(walkers)) / max(1, 8 * walkers - 1)
    out = {"walkers": walkers, "hops_per_launch": hops, "unit": "hops/s"}
    saved = os.environ.get("LMC_KMC_TEAM_LANES")
    try:
        for name, forced in (("half_warp_kernel", "0"), ("library_choice", None)):
            if forced is None:
                os.environ.pop("LMC_KMC_TEAM_LANES", None)
            else:
                os.environ["LMC_KMC_TEAM_LANES"] = forced
            eng.set_occupancy_all(occ)
            eng.kmc_reset()
            eng.kmc_run(hops, temperatures=temps, seed=20260101)
            ms = []
            for _ in range(3):
                eng.kmc_run(hops, temperatures=temps, seed=20260101)
                ms.append(eng.last_kernel_ms())
            st = eng.kmc_state()
            out[name] = {"kernel": kmc_kernel_name(eng.kmc_last_launch_lanes(), eng.kmc_last_launch_resident_occupancy(), eng.kmc_last_launch_handoff()), "value": walkers * hops / (min(ms) * 1e-3), "kernel_ms": min(ms),
                         "us_per_step": min(ms) * 1e3 / hops, "walkers_advanced_all_steps": int(np.count_nonzero(st["steps"] == 4 * hops)),
                         "vacancy_digest": int(np.bitwise_xor.reduce(st["vacancy"] * (np.arange(walkers) + 1)))}
    finally:
        if saved is None:
            os.environ.pop("LMC_KMC_TEAM_LANES", None)
        else:
            os.environ["LMC_KMC_TEAM_LANES"] = saved
    out["same_trajectories"] = out["half_warp_kernel"]["vacancy_digest"] == out["library_choice"]["vacancy_digest"]
    eng.close()
    return out


def bench_single_trajectory(device, json_path):
    """BASELINE configs[0] / configs[4]: ONE vacancy trajectory (a step is a serial dependency, so this is a latency
    number): 10^6-site Al-2%Mg-2%Zn supercell (f = 63), time-temperature ramp + rate corrector, first- and second-order KMC."""
    from latticemontecarlo_b200 import capi, synth
    f, steps = 63, 20000
    eng = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=device)
    eng.load_coefficients(json_path)
    occ = synth.random_alloy(f, P_MG, P_ZN, seed=42, vacancy_site=4 * f ** 3 // 2 + 3)
    tt = np.array([[0.0, 300.0], [1e-3, 500.0], [1e-1, 700.0]])
    out = {"sites": 4 * f ** 3, "steps_per_launch": steps, "unit": "hops/s",
           "workload": "single vacancy, %d sites, T(t) ramp 300-700 K + rate corrector" % (4 * f ** 3)}
    for name, second_order in (("first_order", False), ("second_order_chain", True)):
        eng.set_occupancy(occ)
        eng.kmc_reset()
        eng.kmc_run(2000, time_temperature=tt, rate_corrector=True, seed=7, second_order=second_order)
        ms = []
        for _ in range(3):
            eng.kmc_run(steps, time_temperature=tt, rate_corrector=True, seed=7, second_order=second_order)
            ms.append(eng.last_kernel_ms())
        st = eng.kmc_state()
        out[name] = {"value": steps / (min(ms) * 1e-3), "kernel_ms": min(ms), "us_per_step": min(ms) * 1e3 / steps,
                     "kernel": "kmc_chain_run_kernel" if second_order else kmc_kernel_name(eng.kmc_last_launch_lanes(), eng.kmc_last_launch_resident_occupancy(), eng.kmc_last_launch_handoff()),
                     "time_reached_s": float(st["time"][0]), "temperature_reached_K": float(st["temperature"][0])}
    eng.close()
    return out


def bench_chain(engine, W, temps, occ_pinned, peak, json_path, with_cpu):
    """Secondary metric: second-order KMC (mc::KineticMcChainOmpi, the method script/kmc_param.txt selects) on the same
    walkers: 144 barrier evaluations per hop, one thread block (12 half-warps = the reference's 12 ranks) per walker."""
    from latticemontecarlo_b200 import capi
    hops = 256
    n = int(occ_pinned.numel())
    ms = []
    for it in range(4):
        capi._check(capi.lib().lmc_engine_set_occupancy_all(engine.h, occ_pinned.numpy().ctypes.data_as(capi.C.c_void_p), capi.C.c_int64(n)))
        engine.kmc_reset()
        engine.kmc_chain_run(hops, temperatures=temps, seed=20260101)
        if it:
            ms.append(engine.last_kernel_ms())
    st = engine.kmc_state()
    rate = W * hops / (sum(ms) / len(ms) * 1e-3)
    achieved = rate * 12 * BYTES_PER_KMC_STEP / 1e9
    out = {"metric": "kmc_chain_hops_per_s", "unit": "hops/s", "value": rate, "walkers": W, "hops_per_walker": hops,
           "kernel_ms": sum(ms) / len(ms), "barrier_evaluations_per_hop": 144, "all_walkers_advanced": bool((st["steps"] == hops).all()),
           "roofline": {"kernel": "kmc_chain_run_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": recorded_traffic("kmc_chain_run_kernel", walkers=W, hops=hops),
                        "note": "12 x the first-order algorithmic bytes per hop (144 events)"}}
    if with_cpu:
        try:
            from oracle import ref_lib as R
            from latticemontecarlo_b200 import synth
            occ = walker_occupancy(0, 1)[0]
            perm = synth.generate_to_reassigned_permutation(FACTOR)
            occ_gen = np.empty_like(occ)
            occ_gen[perm] = occ
            cfg = R.RefConfig.fcc(FACTOR, occ_gen, reassign=True)
            steps = 3000
            res = R.kmc_chain_ompi(cfg, json_path, temperature=500.0, maximum_steps=steps - 1, seed=1, trace=False)
            out["cpu_baseline"] = {"value": steps / res["seconds"], "unit": "hops/s", "cores": min(12, host_cores()), "kind": "reference",
                                   "sample": "one mc::KineticMcChainOmpi trajectory, its 12 MPI ranks as 12 threads of one process "
                                             "(in-process MPI shim), %d hops on one 8x8x8 walker; %.1f s in Simulate()" % (steps, res["seconds"])}
        except Exception as exc:
            out["cpu_baseline"] = {"value": None, "unit": "hops/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (exc,)}
    return out


def bench_cmc(torch, device, json_path, peak, with_cpu):
    """Secondary metric: CMC swap trials/s on BASELINE configs[1] (one 40x40x40 Al-2%Mg-2%Zn lattice, 800 K), configs[3]
    (SimulatedAnnealing on 100x100x100 = 4M sites) and 148 independent 20^3 replicas at 600..1000 K, with both drivers:
      `domain`  lmc_cmc_domain_run -- spatial domains, pairs drawn inside a domain core, one grid barrier per sweep (the
                north_star's decomposition; semantic change stated in include/lmc_b200.h; ensemble parity with
                mc::CanonicalMcSerial in tests/test_gpu_cmc_stat.py)
      `global_pairs`  lmc_cmc_run -- the reference's global pair draw, batches of mutually non-interfering trials.
    Rates are kernel rates (CUDA events on the engine stream); `fresh` = the first trials from the random alloy, `aged` =
    after about 300 trials per site at 800 K (solute has clustered: longer table walks); `e2e` includes H2D occupancy,
    reset, run and D2H occupancy through the C ABI."""
    from latticemontecarlo_b200 import capi, synth
    out = {"metric": "cmc_swap_trials_per_s", "unit": "trials/s", "bytes_per_trial": BYTES_PER_TRIAL}
    cases = (("single_lattice_40x40x40", 40, 1, None, 8),                              # BASELINE configs[1]
             ("single_lattice_100x100x100_sa", 100, 1, (900.0, 40000000000), 8),       # configs[3]: SimulatedAnnealing, 4M sites
             ("replicas_148x_20x20x20", 20, 148, None, 8))
    for name, f, replicas, sa, trials_per_site in cases:
        n_sites = 4 * f ** 3 * replicas
        eng = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=replicas, device=device)
        eng.load_coefficients(json_path)
        occ = np.stack([synth.random_alloy(f, P_MG, P_ZN, seed=1000 + r, vacancy_site=None) for r in range(replicas)])
        pinned = torch.empty(occ.shape, dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[:] = occ
        temps = np.linspace(600.0, 1000.0, replicas) if replicas > 1 else np.array([800.0])
        eng.set_occupancy_all(pinned.numpy())
        if name == "single_lattice_40x40x40":
            out["swap_de_eval"] = bench_swap_de_eval(torch, eng, occ[0], peak)
        entry = {"replicas": replicas, "sites": 4 * f ** 3, "driver": "SimulatedAnnealing schedule" if sa else "CanonicalMc at fixed temperature"}

        def timed(run, per_launch, launches):
            kernel_ms, done = [], []
            for _ in range(launches):
                s0 = eng.cmc_state()["steps"].sum()
                run(per_launch)
                kernel_ms.append(eng.last_kernel_ms())
                done.append(int(eng.cmc_state()["steps"].sum() - s0))
            return sum(done) / (sum(kernel_ms) * 1e-3), float(np.mean(kernel_ms)), int(np.mean(done))

        # ---- domain driver
        per_launch = trials_per_site * 4 * f ** 3                    # per replica
        run_dom = lambda n: eng.cmc_domain_run(n, temperatures=temps, seed=5)
        eng.cmc_reset(*(sa or ()))
        run_dom(per_launch // 4)                                      # warm-up
        eng.set_occupancy_all(pinned.numpy()); eng.cmc_reset(*(sa or ()))
        fresh, fresh_ms, fresh_n = timed(run_dom, per_launch, 3)
        run_dom(300 * 4 * f ** 3 - 3 * per_launch)                    # age the alloy
        aged, aged_ms, aged_n = timed(run_dom, per_launch, 3)
        st = eng.cmc_state()
        t0 = time.perf_counter()
        n_e2e = 0
        for _ in range(3):
            eng.set_occupancy_all(pinned.numpy())
            eng.cmc_reset(*(sa or ()))
            run_dom(per_launch)
            n_e2e += int(eng.cmc_state()["steps"].sum())
            eng.get_occupancy_all()
        e2e = n_e2e / (time.perf_counter() - t0)
        entry["domain"] = {"value": fresh, "aged": aged, "e2e": e2e, "kernel_ms": fresh_ms, "aged_kernel_ms": aged_ms, "trials_per_launch": fresh_n,
                           "accept_ratio_aged": float(st["accepted"].sum() / max(1, st["steps"].sum())), "shape": eng.cmc_domain_last_shape(),
                           "roofline": {"kernel": "cmc_domain_kernel", "bound": "hbm", "achieved": fresh * BYTES_PER_TRIAL / 1e9, "peak": peak, "unit": "GB/s",
                                        "frac": fresh * BYTES_PER_TRIAL / 1e9 / peak, "frac_aged": aged * BYTES_PER_TRIAL / 1e9 / peak,
                                        "traffic": recorded_traffic("cmc_domain_kernel", replicas=replicas, factor=f),
                                        "note": "effective bandwidth (440 algorithmic B per trial, SURVEY 8(d)); a domain lives in shared memory for "
                                                "a whole sweep, so the physical traffic is one read + one write of the occupancy per sweep"}}
        # ---- global-pair driver (the reference's proposal)
        trials = 200000 if f == 40 else (2000000 if f == 100 else 20000)
        sa_g = (sa[0], 40000000) if sa else ()
        run_glob = lambda n: eng.cmc_run(n, temperatures=temps, seed=5)
        eng.set_occupancy_all(pinned.numpy()); eng.cmc_reset(*sa_g)
        run_glob(trials // 4)
        g_rate, g_ms, g_n = timed(run_glob, trials, 5)
        st = eng.cmc_state()
        kernel = "cmc_grid_kernel" if replicas == 1 else "cmc_run_kernel"     # one lattice: whole-GPU cooperative kernel
        entry["global_pairs"] = {"value": g_rate, "kernel_ms": g_ms, "trials_per_launch": g_n,
                                 "accept_ratio": float(st["accepted"].sum() / max(1, st["steps"].sum())),
                                 "roofline": {"kernel": kernel, "bound": "hbm", "achieved": g_rate * BYTES_PER_TRIAL / 1e9, "peak": peak, "unit": "GB/s",
                                              "frac": g_rate * BYTES_PER_TRIAL / 1e9 / peak,
                                              "traffic": recorded_traffic(kernel, replicas=replicas, factor=f, trials=trials)}}
        entry["value"] = fresh
        out[name] = entry
        eng.close()
    out["value"] = out["single_lattice_40x40x40"]["value"]
    if with_cpu:
        out["cpu_baseline"] = cpu_baseline_cmc(json_path)
    return out


def bench_swap_de_eval(torch, eng, occ, peak, n=1 << 22):
    """Batched EnergyChangePredictorPairSite::GetDeFromLatticeIdPair with resident inputs: n random UNLIKE-species site pairs
    of the 256k-site lattice, drawn like CanonicalMcAbstract::GenerateLatticeIdJumpPair (redraw while the two species are
    equal, CanonicalMcAbstract.cpp:43-51), so every pair costs two 43-site gathers; ~42/N of the pairs are coupled."""
    rng = np.random.default_rng(11)
    n_sites = occ.size
    a = rng.integers(0, n_sites, n)
    b = rng.integers(0, n_sites, n)
    same = np.nonzero(occ[a] == occ[b])[0]
    while same.size:
        b[same] = rng.integers(0, n_sites, same.size)
        same = same[occ[a[same]] == occ[b[same]]]
    dev = torch.device("cuda")
    d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    d_de = torch.empty(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    times = []
    for k in range(8):
        eng.eval_swap_de_dev(n, 0, d_a.data_ptr(), d_b.data_ptr(), d_de.data_ptr())
        ms = eng.last_kernel_ms()
        if k >= 3:
            times.append(ms)
    ms = sum(times) / len(times)
    achieved = n * BYTES_PER_TRIAL / (ms * 1e-3) / 1e9
    check = eng.eval_swap_de(a[:4096], b[:4096])               # the host-buffer entry point on the same pairs
    return {"kernel": "swap_de_rows_kernel", "pairs": n, "ms": ms, "pairs_per_s": n / (ms * 1e-3), "achieved_gbs": achieved,
            "frac_of_hbm_peak": achieved / peak, "bytes_per_trial": BYTES_PER_TRIAL,
            "finite": bool(torch.isfinite(d_de).all().item()),
            "matches_host_entry_point": bool(np.array_equal(check, d_de[:4096].cpu().numpy())),
            "note": "random unlike-species pairs of one 40x40x40 lattice (every pair is evaluated: 2 x 43-site gathers)"}


def bench_cmc_multi_gpu(torch, dist, local_rank, rank, world, json_path):
    """BASELINE configs[3] / [1] over all GPUs of the job: ONE lattice annealed by `world` ranks with the domain driver
    (x slabs of the domain grid, halo / migration rows written into peer memory inside the persistent kernel:
    include/lmc_b200.h, lmc_cmc_domain_attach_peers).  Every rank takes part; the rate uses the slowest rank's kernel time;
    all ranks must end identical AND equal to the single-GPU run of the same seed, which rank 0 computes in the same job."""
    import hashlib
    from latticemontecarlo_b200 import capi, sharding, synth
    out = {}
    for name, f, sa, trials in (("single_lattice_100x100x100_sa", 100, (900.0, 40000000000), 8 * 4000000), ("single_lattice_40x40x40", 40, (), 8 * 256000)):
        occ = synth.random_alloy(f, P_MG, P_ZN, seed=1000, vacancy_site=None)

        def digest(e):
            st = e.cmc_state()
            return hashlib.sha256(e.get_occupancy(0).tobytes()).hexdigest() + "%.17g %d %d %.17g" % (st["energy"][0], st["steps"][0], st["accepted"][0], st["temperature"][0])

        def run(e, collective):
            e.set_occupancy(occ)
            e.cmc_reset(*sa)
            e.cmc_domain_run(trials // 4, seed=5)
            ms, done = [], []
            for _ in range(3):
                if collective:
                    dist.barrier(); torch.cuda.synchronize()
                s0 = int(e.cmc_state()["steps"][0])
                e.cmc_domain_run(trials, seed=5)
                ms.append(sharding.max_over_ranks([e.last_kernel_ms()], device="cuda")[0] if collective else e.last_kernel_ms())
                done.append(int(e.cmc_state()["steps"][0]) - s0)
            return sum(done) / (sum(ms) * 1e-3), float(np.mean(ms)), int(np.mean(done)), digest(e)

        eng = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=local_rank)
        eng.load_coefficients(json_path)
        sharding.attach_cmc_domain_peers(eng, rank, world)
        dist.barrier(); torch.cuda.synchronize()
        rate, ms, done, dg = run(eng, True)
        shape = eng.cmc_domain_last_shape()
        digests = [None] * world
        dist.all_gather_object(digests, dg)
        eng.close()
        single = None
        if rank == 0:
            ref = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=local_rank)
            ref.load_coefficients(json_path)
            r1, ms1, _, dg1 = run(ref, False)
            ref.close()
            single = {"value": r1, "kernel_ms": ms1, "equals_multi_gpu_run": dg1 == dg}
        out[name] = {"value": rate, "unit": "trials/s", "n_gpus": world, "sites": 4 * f ** 3, "scaling": "strong",
                     "driver": "SimulatedAnnealing schedule" if sa else "CanonicalMc at fixed temperature", "trials_per_launch": done, "kernel_ms": ms,
                     "shape": shape, "ranks_identical": len(set(digests)) == 1, "single_gpu_same_job": single,
                     "speedup_vs_single_gpu_same_job": rate / single["value"] if single else None,
                     "exchange": "domain rows written into the next sweep's holders through NVLink peer mappings inside cmc_domain_kernel; "
                                 "sweep totals as flag-carrying lines (inter-GPU barrier); no NCCL call on the data path"}
        dist.barrier()
    return out


def cpu_baseline_cmc(json_path):
    """mc::CanonicalMcOmp (batch = OMP thread count) on this box's host cores.  The per-trial cost does not depend on the
    lattice size, so the sample uses a 20x20x20 cell (the reference's per-site tables for 256k sites take minutes to build)."""
    try:
        from oracle import ref_lib as R
        from latticemontecarlo_b200 import synth
        if not R.build():
            return {"value": None, "unit": "trials/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref unavailable"}
        cores = host_cores()
        f = 20
        occ = synth.random_alloy(f, P_MG, P_ZN, seed=1000, vacancy_site=None)
        cfg = R.RefConfig.fcc(f, occ, reassign=False)
        steps = 40000
        res = R.cmc_omp(cfg, json_path, temperature=800.0, maximum_steps=steps, seed=3, threads=cores)
        return {"value": res["steps"] / res["seconds"], "unit": "trials/s", "cores": cores, "kind": "reference",
                "sample": "mc::CanonicalMcOmp, %d OMP threads, %d trials on a 20x20x20 Al-2%%Mg-2%%Zn cell at 800 K (%.1f s in Simulate())"
                          % (cores, res["steps"], res["seconds"])}
    except Exception as exc:
        return {"value": None, "unit": "trials/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (exc,)}


def cpu_baseline(json_path):
    """The reference's own KMC driver on this box's host cores, bounded sample (about 10-20 s of CPU work)."""
    try:
        from oracle import ref_lib as R
        if not R.build():
            return {"value": None, "unit": "hops/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref unavailable"}
        cores = host_cores()
        steps = 20000
        rate, sim_s, wall = reference_kmc_throughput(json_path, steps, cores)
        return {"value": rate, "unit": "hops/s", "cores": cores, "kind": "reference",
                "sample": "%d concurrent mc::KineticMcFirstOmp trajectories (LRU on, 1 OMP thread each) x %d hops on the 8x8x8 "
                          "workload; %.1f s inside Simulate(), %.1f s wall incl. predictor construction" % (cores, steps, sim_s, wall)}
    except Exception as exc:  # the baseline must never take the GPU number down with it
        return {"value": None, "unit": "hops/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (exc,)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--walkers", type=int, default=8192, help="walkers of the job (strong scaling: w mod N); also walkers per GPU of the weak-scaling extra")
    ap.add_argument("--hops", type=int, default=2048, help="KMC steps per walker per bench step")
    ap.add_argument("--ref-hops", type=int, default=20000, help="reference arm: hops per trajectory per step (same length as cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cmc", action="store_true", help="skip the secondary CMC measurement")
    ap.add_argument("--no-age", action="store_true", help="skip the value-vs-walker-age and rich-alloy extras")
    ap.add_argument("--no-chain", action="store_true", help="skip the secondary second-order KMC measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
