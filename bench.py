#!/usr/bin/env python
"""bench.py - individual-sites per second per EM iteration on B200.

One "step" = one EM iteration of the hot path on a fixed synthetic input (the reference's iter_EM,
EM.cpp:139-289).  Workloads (BASELINE.json configs, weak scaling: the per-GPU work is fixed):

  --config 1 (default)  configs[1] per GPU: 100 individuals x 1,000,000 sites, --freq_est 1, start values
                        --freq 0.1 --indF 0.1,0.2: E-step + lockstep L-BFGS-B update of every (F, alpha) around
                        batched objective launches + per-site allele-frequency EM fused with the emission refresh.
  --config 2            configs[2] at 8 GPUs: 125 individuals x 10,000,000 sites per GPU, same step.
  --config 4            configs[4] at 8 GPUs: 1,250 individuals x 1,000,000 sites per GPU, --indF_fixed
                        --alpha_fixed --freq_est 0: the step is the fixed-parameter iteration (forward-backward +
                        posterior); the Viterbi tract decoding is timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|4]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "ind-sites/sec per EM iteration"
UNIT = "ind-sites/s"
CONFIGS = {   # individuals per GPU, sites, free parameters
    1: dict(n_ind=100, n_sites=1_000_000, fixed=False, name="configs[1]"),
    2: dict(n_ind=125, n_sites=10_000_000, fixed=False, name="configs[2]"),
    4: dict(n_ind=1250, n_sites=1_000_000, fixed=True, name="configs[4]"),
}
START_FREQ, START_F, START_ALPHA = 0.1, 0.1, 0.2
# algorithmic work per unit, stated in DESIGN.md "Measurement"
ESTEP_BYTES_PER_IND_SITE = 24.0          # read 2 emissions, write 1 posterior
FREQ_FLOPS_PER_IND_PASS = 13.75          # odds-form est_maf contribution: 8 FP64 instructions (5.75 of them FMA)
FREQ_INSTR_PER_IND_PASS = 8.0            # -> at most 13.75/16 of the DFMA peak; passes per site are counted by the kernel
LKL_FLOPS_PER_IND_SITE_POINT = 14.0      # factored 2x2 update: 2 ADD + 4 FMA + 4 MUL per objective point and site
EXP_FLOPS = 21.0                         # kappa = expm1(alpha d), table tier: 13 FP64 instructions, 8 of them FMA; 3 per 5 points
FP64_LANES_PER_SM = 64                   # B200: 64 FP64 FMA lanes per SM -> arithmetic peak = SMs x 64 x 2 x clock


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS))
    ap.add_argument("--n_ind", type=int, default=None, help="individuals per GPU (default: the config's)")
    ap.add_argument("--n_sites", type=int, default=None)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_parity_check", action="store_true")
    ap.add_argument("--trace", action="store_true", help="multi-rank: per-phase stream times of every rank (extra untimed leg)")
    ap.add_argument("--cpu_sites", type=int, default=None,
                    help="sites of the bounded CPU sample (reference arm); default: sized from a calibration "
                         "iteration so that the W + K iterations fit --cpu_budget_s, at most 20,000")
    ap.add_argument("--cpu_budget_s", type=float, default=float(os.environ.get("NFH_REF_BUDGET_S", "150")),
                    help="wall-clock target of the whole reference arm (W + K iterations)")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    a.n_ind = a.n_ind or cfg["n_ind"]
    a.n_sites = a.n_sites or cfg["n_sites"]
    a.fixed = cfg["fixed"]
    a.cfg_name = cfg["name"]
    return a


# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed regions (B200_PROFILING.md): NVML every
    10 ms when pynvml is importable, else `nvidia-smi -lms 100`."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0, uuid=None):
        self.rows, self.proc, self.index, self.uuid = [], None, index, uuid
        self.nvml, self.handle, self.stop_flag, self.t = None, None, threading.Event(), None
        self.sm, self.mx, self.pw, self.reasons = [], [], [], set()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = (pynvml.nvmlDeviceGetHandleByUUID(self.uuid) if self.uuid
                           else pynvml.nvmlDeviceGetHandleByIndex(self.index))
            self.nvml = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.pw.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag.set()
            self.t.join(timeout=2)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
            for r in self.rows:
                f = [x.strip() for x in r.split(",")]
                if len(f) < 7:
                    continue
                try:
                    self.sm.append(float(f[0])); self.mx.append(float(f[1])); self.pw.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(self.NAMES, f[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "power_w_max": max(self.pw) if self.pw else None, "samples": len(self.sm),
                "source": "nvml" if self.nvml else "nvidia-smi", "reasons": sorted(self.reasons)}


def ncu_traffic():
    """DRAM bytes per individual-site measured by ncu (committed under profiles/), or None."""
    for name in ("ncu_traffic_r02.json", "ncu_traffic_r01f.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                return json.load(fh)["bytes_per_ind_site"]
        except Exception:  # noqa: BLE001
            continue
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_text(args, n_total):
    if args.fixed:
        flags = "--indF FILE --indF_fixed --alpha_fixed --freq FILE --freq_est 0 (true parameter values)"
    else:
        flags = "--freq_est 1, start --freq 0.1 --indF 0.1,0.2"
    return (f"{args.cfg_name} per GPU: {args.n_ind} individuals x {args.n_sites} sites, depth-2 GL, {flags}; "
            f"{n_total} individuals total")


# ---------------------------------------------------------------------------
def reference_arm(args, as_cpu_baseline=False):
    """The UNMODIFIED reference (oracle/_ref, built from /root/reference) timed on host cores: one step = one
    iter_EM() on a bounded sample of the workload (same generator, same start values, same flags)."""
    from _oracle import Ref
    import ngsf_hmm_b200  # noqa: F401
    from ngsf_hmm_b200 import sim
    cores = os.cpu_count() or 1
    N = min(args.n_ind, 100)                    # the reference holds ~250 B per individual-site
    threads = min(cores, N)
    steps, warm = (1, 0) if as_cpu_baseline else (args.steps, args.warmup)
    ref = Ref()

    def make_state(S):
        d = sim.simulate_torch(N, S, device="cpu", seed=1002, site_chunk=1 << 14)
        if args.fixed:
            return ref.state(d["log_gl"].numpy(), d["dist_mb"], np.clip(d["true_freq"], 0.01, 0.49),
                             np.clip(d["true_F"], 1e-6, 1 - 1e-6), d["true_alpha"], freq_est=0, n_threads=threads,
                             indF_fixed=True, alpha_fixed=True)
        return ref.state(d["log_gl"].numpy(), d["dist_mb"], START_FREQ, START_F, START_ALPHA, freq_est=1,
                         n_threads=threads)

    # The sample is bounded by TIME: the driver chooses K and W, and a reference iteration costs ~1e-5 s per
    # individual-site, so a fixed site count would run for many minutes at large K.  One calibration iteration on
    # 2,000 sites (the first EM iteration, the one with the most BFGS rounds) gives the rate; the sample then gets
    # as many sites as W + K iterations can cover within --cpu_budget_s, between 2,000 and 20,000.
    sized = "--cpu_sites"
    S = args.cpu_sites
    if S is None:
        S_cal = 2000
        st = make_state(S_cal)
        t0 = time.perf_counter()
        st.iter_EM()
        t_cal = time.perf_counter() - t0
        st.close()
        S = int(0.8 * args.cpu_budget_s / max(steps + warm, 1) / max(t_cal, 1e-6) * S_cal)
        S = max(2000, min(20000, S // 500 * 500))
        sized = (f"a calibration iteration on {S_cal} sites ({t_cal:.2f} s) and a budget of {args.cpu_budget_s:.0f} s "
                 f"for {warm} + {steps} iterations")
    st = make_state(S)
    for _ in range(warm):
        st.iter_EM()
    t0 = time.perf_counter()
    for _ in range(steps):
        st.iter_EM()
    dt = (time.perf_counter() - t0) / steps
    st.close()
    value = N * S / dt
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
          "sample": f"{N} individuals x {S} sites of the same synthetic workload, {warm} warm-up + {steps} timed "
                    f"iter_EM() call(s) of the unmodified reference in-process (oracle/_ref), --n_threads {threads} "
                    f"of {cores} host cores; its frequency loop is serial; sites chosen by {sized}.  A per-unit RATE "
                    f"on a sample, not the same job timed twice"}
    if as_cpu_baseline:
        return cb
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args, args.n_ind * max(args.gpus, 1)),
                       "step": ("one fixed-parameter EM iteration = forward-backward E-step with posteriors"
                                if args.fixed else
                                "one EM iteration = E-step + BFGS(F,alpha) per individual + freq EM + emission refresh"),
                       "sample": f"each step runs on a bounded sample of that workload: {N} individuals x {S} sites "
                                 f"(rank 0's host cores only; see cpu_baseline.sample)"},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ---------------------------------------------------------------------------
def upload_synthetic(ctx, sim, dev, n_total, n_sites, seed=1002):
    """Synthetic GL for this rank's SITE block, all individuals, generated on the GPU chunk by chunk and handed
    to nfh_upload_gl as device memory (no host copy).  Returns (dist_mb, true_F, true_alpha, true_freq of the
    block, seconds generating, seconds uploading, bytes)."""
    import torch
    chunk = 1 << 18
    while chunk > 4096 and chunk * n_total > 2.2e8:
        chunk >>= 1
    s0, s1 = ctx.site_begin, ctx.site_begin + ctx.sites_owned
    buf = torch.empty((chunk, n_total, 3), dtype=torch.float64, device=dev)
    t_gen = t_up = 0.0
    meta, freq_parts = None, []
    c0 = (s0 // chunk) * chunk
    while c0 < s1:
        lo, hi = max(c0, s0), min(c0 + chunk, s1)
        t0 = time.perf_counter()
        gen = sim.simulate_torch(n_total, n_sites, device=dev, seed=seed, site_chunk=chunk, site_begin=lo, site_end=hi,
                                 out=buf[:hi - lo])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ctx.upload_gl(buf[:hi - lo], first_site=lo)
        t_gen += t1 - t0; t_up += time.perf_counter() - t1
        freq_parts.append(gen["true_freq"])
        meta = gen
        c0 += chunk
    if meta is None:                            # a rank without sites still needs the distances
        meta = sim.simulate_torch(n_total, n_sites, device=dev, seed=seed, site_chunk=chunk, site_begin=0, site_end=0)
    del buf
    torch.cuda.empty_cache()
    freq = np.concatenate(freq_parts) if freq_parts else np.empty(0)
    return meta["dist_mb"], meta["true_F"], meta["true_alpha"], freq, t_gen, t_up, ctx.sites_owned * n_total * 24


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args)), flush=True)
        return 0

    import torch
    import ngsf_hmm_b200 as nfh
    from ngsf_hmm_b200 import selfcheck, sim

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    n_ranks = world
    N_total, S = args.n_ind * n_ranks, args.n_sites
    # exchange between the recursion side and the frequency side.  Measured on 8 B200 (profiles/r02):
    #   all-to-alls both ways ("nccl", the posterior one behind the BFGS rounds): 20.5 ms per EM iteration of configs[1]
    #     at 8 GPUs, 314 ms of configs[2];
    #   kernels storing into peer windows ("direct"): 21.8 / 351 ms - the posterior tiles of all ranks hit NVLink in
    #     the same millisecond and the E-step + BFGS phase of most ranks stretches from 4.1 to 6.3 ms;
    #   "mixed" (posteriors by all-to-all, emission ratios by peer stores from the frequency kernel): best at 2 and 4
    #     GPUs (15.6 / 16.8 ms against 16.4 / - ) and the default there; not yet run at 8 (its first 8-rank run hung in
    #     the self-check on the ranks that own no individual of the 11; fixed in bfgs_update_lockstep, covered by a
    #     CPU test and by an emulated 8-rank run since - tests/simt/run_multi_rank_emulated.py - but there was no GPU
    #     time left to TIME it), so 8 ranks keep the all-to-alls until it has been;
    #   fixed frequencies: nothing to exchange per iteration ("direct" = emission refresh by peer stores, once).
    # NFH_EXCHANGE=direct|mixed|nccl overrides (NFH_PEER_DIRECT=0/1 = nccl/direct is still honoured).
    mode = os.environ.get("NFH_EXCHANGE", "")
    if not mode and os.environ.get("NFH_PEER_DIRECT", "") in ("0", "1"):
        mode = "direct" if os.environ["NFH_PEER_DIRECT"] == "1" else "nccl"
    if not mode:
        mode = "direct" if args.fixed else ("mixed" if world <= 4 else "nccl")
    if world == 1:
        mode = "none"
    direct = {"direct": True, "mixed": "mixed", "nccl": False, "none": False}[mode]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- multi-rank self-check (untimed): the sharded path against one rank, before anything is timed
    parity = None
    if world > 1 and not args.no_parity_check:
        parity = selfcheck.multi_rank_check(local_rank, direct=direct)
        flag = torch.tensor([1 if (rank != 0 or parity["ok"]) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if rank == 0:
                print(json.dumps({"error": "multi-rank parity check failed", "parity_check": parity}), flush=True)
            dist.destroy_process_group()
            return 1

    # ---- setup (untimed for `value`, counted in e2e_full_run): synthetic GL -> device -> nfh_upload_gl
    t_run0 = time.perf_counter()
    ctx = nfh.Context(N_total, S, device=local_rank, n_ranks=n_ranks, rank=rank)
    dist_mb, true_F, true_alpha, true_freq, t_gen, t_upload, gl_bytes = upload_synthetic(ctx, sim, dev, N_total, S)
    t0 = time.perf_counter()
    ctx.upload_pos_dist(dist_mb)
    t_upload += time.perf_counter() - t0

    freq_est = 0 if args.fixed else 1
    runner = nfh.EmRank(ctx, freq_est=freq_est, indF_fixed=args.fixed, alpha_fixed=args.fixed)
    exchange = "none (1 rank)"
    if world > 1:
        if direct:
            runner.enable_peer_direct(posteriors=direct != "mixed")
            exchange = ("fused: kernels store into peer windows over NVLink (CUDA IPC)" if direct is True else
                        "mixed: posteriors by NCCL all-to-all behind the BFGS rounds, emission ratios stored into peer "
                        "windows by the frequency kernel")
        else:
            exchange = "NCCL all-to-all"
    n_own = ctx.n_ind_owned
    own = slice(ctx.ind_begin, ctx.ind_begin + n_own)

    def reset_state():
        if args.fixed:
            ctx.set_freq(np.clip(true_freq, 0.01, 0.49))
            F = np.clip(true_F[own], 1e-6, 1 - 1e-6).copy(); a = true_alpha[own].copy()
        else:
            ctx.set_freq(np.full(ctx.sites_owned, START_FREQ))
            F = np.full(n_own, START_F); a = np.full(n_own, START_ALPHA)
        ctx.set_ind_params(F, a)
        runner.refresh_emissions()
        return F, a

    ext = torch.cuda.ExternalStream(ctx.stream)
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)

    # ---- settling pass (untimed, discarded): the W + K iterations the legs below repeat, run once beforehand, so that
    #      one-off costs of a fresh process - lazily loaded kernel variants (the objective kernel has one per point
    #      layout), NCCL connections, the page-locked frequency buffer - do not land in the first timed leg (on 2 GPUs
    #      the first leg was 1.3 ms per step slower than the identical later ones)
    F, a = reset_state()
    for _ in range(args.warmup + args.steps):
        runner.iteration(F, a, want_freq=bool(freq_est))

    # ---- device-timed leg: reset, W warm-up iterations, then K EM iterations, state resident in HBM
    F, a = reset_state()
    for _ in range(args.warmup):
        runner.iteration(F, a, want_freq=bool(freq_est))
    ctx.timing(os.environ.get("NFH_BENCH_NO_FAMILY_TIMING", "") != "1")
    ctx.timing_read(reset=True)
    ctx.freq_passes(reset=True)
    launches0 = ctx.kernel_launches
    evals0, rounds0 = runner.total_evals, runner.total_rounds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    e0.record(ext)
    lk = None
    for _ in range(args.steps):
        lk, _ = runner.iteration(F, a, want_freq=False)
    e1.record(ext)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    fam = ctx.timing_read(reset=True)
    site_passes = ctx.freq_passes(reset=True)      # rank 0's sites, all K steps
    ctx.timing(False)
    launches = ctx.kernel_launches - launches0
    evals, rounds = runner.total_evals - evals0, runner.total_rounds - rounds0

    # ---- end-to-end leg: the SAME iterations (reset + the same warm-up first) through the host-buffer API, wall
    #      clock, parameters uploaded and lkl / F / alpha / freq downloaded every step
    F, a = reset_state()
    for _ in range(args.warmup):
        runner.iteration(F, a, want_freq=bool(freq_est))
    barrier()
    t0 = time.perf_counter()
    fr = None
    for _ in range(args.steps):
        lk, fr = runner.iteration(F, a, want_freq=bool(freq_est))
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    h2d = 2 * n_own * 8 + (evals / max(args.steps, 1)) * (4 + 8 + 8)
    d2h = n_own * 8 + (ctx.sites_owned * 8 if freq_est else 0) + (evals / max(args.steps, 1)) * 8

    # ---- optional: where a multi-rank step goes, rank by rank (separate untimed leg, CUDA events between the phases)
    rank_trace = None
    if args.trace and world > 1 and freq_est:
        F, a = reset_state()
        for _ in range(args.warmup):
            runner.iteration(F, a, want_freq=False)
        runner.trace = []
        r0 = runner.total_rounds
        barrier()
        tt0 = time.perf_counter()
        for _ in range(args.steps):
            runner.iteration(F, a, want_freq=False)
        barrier()
        trace_ms = (time.perf_counter() - tt0) * 1e3 / args.steps
        tsum = runner.trace_summary()
        runner.trace = None
        mine = torch.tensor([tsum["estep_bfgs"], tsum["wait_ranks"], tsum["freq"], tsum["exchange_back"],
                             (runner.total_rounds - r0) / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_trace = {"wall_ms_per_step_of_this_leg": trace_ms,
                      "columns": ["estep_bfgs_ms", "wait_ranks_ms", "freq_ms", "exchange_back_ms", "bfgs_rounds"],
                      "per_rank": [[round(float(x), 3) for x in t] for t in allr]}

    # ---- the E-step on its own (products + carries + apply): inside a free-parameter EM iteration its forward
    #      products ride on the first objective round, so the per-step "estep" time covers carries + apply only
    ctx.timing(True)
    ctx.timing_read(reset=True)
    for _ in range(5):
        ctx.estep_async()
    ctx.sync()
    estep_alone_ms = ctx.timing_read(reset=True)["estep"][0] / 5.0

    # ---- end of a run: emission refresh with e0, Viterbi decoding, downloads (one-off per run)
    t1 = time.perf_counter()
    runner.refresh_emissions(with_e0=True)
    ctx.set_ind_params(F, a)
    ctx.sync()
    t2 = time.perf_counter()
    path = ctx.viterbi()
    t3 = time.perf_counter()
    vit = ctx.timing_read(reset=True)
    ctx.timing(False)
    post = ctx.get_posterior()
    t4 = time.perf_counter()
    t_full = t4 - t_run0
    fp64_probe = ctx.probe_fp64()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count

    times = torch.tensor([dev_ms, e2e_s * 1e3, estep_alone_ms, (t4 - t1) * 1e3, vit["viterbi"][0]], dtype=torch.float64,
                         device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, estep_alone_ms, t_final_ms, vit_ms = (float(x) for x in times)

    if rank == 0:
        units = float(N_total) * S * args.steps
        value = units / (dev_ms * 1e-3)
        hbm_peak, peak_src = measured_peaks()
        per_step = {k: v[0] / args.steps for k, v in fam.items()}
        rank_units = float(n_own) * S                       # recursion-side units of this rank per step
        freq_units = float(N_total) * ctx.sites_owned       # frequency-side units of this rank per step
        estep_gbs = ESTEP_BYTES_PER_IND_SITE * rank_units / (estep_alone_ms * 1e-3) / 1e9
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp64_arith = sm_count * FP64_LANES_PER_SM * 2 * sm_mhz * 1e6
        fp64_note = {"probe_tflops": fp64_probe / 1e12, "arithmetic_tflops": fp64_arith / 1e12,
                     "arithmetic": f"{sm_count} SMs x {FP64_LANES_PER_SM} FP64 lanes x 2 flop x {sm_mhz:.0f} MHz (median "
                                   f"SM clock under load)",
                     "frac_uses": "the live DFMA probe (nfh_probe_fp64); frac_of_arithmetic is given beside it"}
        traffic = ncu_traffic() or {}
        t_estep = traffic.get("estep") and traffic["estep"] * rank_units
        roof_estep = {"kernel": "estep (chunk_products + tile_carries + chunk_apply)", "bound": "hbm",
                      "achieved": estep_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": estep_gbs / hbm_peak,
                      "traffic": t_estep, "peak_source": peak_src, "ms_standalone": estep_alone_ms,
                      "ms_per_step": per_step["estep"], "n_sites": S, "n_ind": n_own,
                      "note": "achieved/frac are of the stand-alone E-step (products + carries + apply) at 24 B per "
                              "individual-site; inside a free-parameter EM iteration the forward products are shared "
                              "with the first objective round (nfh_estep_with_batch) and ms_per_step covers carries + "
                              "apply"}
        roofs = {"estep": roof_estep}
        if not args.fixed:
            passes_per_site = site_passes / max(ctx.sites_owned * args.steps, 1)
            freq_tf = (FREQ_FLOPS_PER_IND_PASS * passes_per_site * freq_units / (per_step["freq"] * 1e-3) / 1e12
                       if per_step["freq"] > 0 else 0.0)
            evals_step = evals / args.steps
            lkl_flops = (LKL_FLOPS_PER_IND_SITE_POINT + 0.6 * EXP_FLOPS) * evals_step * S
            lkl_tf = lkl_flops / (per_step["lkl_batch"] * 1e-3) / 1e12 if per_step["lkl_batch"] > 0 else 0.0
            t_freq = traffic.get("freq") and traffic["freq"] * freq_units
            t_lkl = traffic.get("lkl_batch_per_group_site") and traffic["lkl_batch_per_group_site"] * rank_units
            roofs["freq"] = {
                "kernel": "freq_emission_*", "bound": "fp64", "achieved": freq_tf, "peak": fp64_probe / 1e12,
                "unit": "TFLOP/s", "frac": freq_tf / (fp64_probe / 1e12), "frac_of_arithmetic": freq_tf / (fp64_arith / 1e12),
                "traffic": t_freq, "passes_per_site": passes_per_site,
                "bound_note": "FP64 CUDA-core pipe (DFMA), not tensor cores: the work is per-individual rational "
                              "functions with no dense contraction",
                "instruction_mix_ceiling": FREQ_FLOPS_PER_IND_PASS / (2 * FREQ_INSTR_PER_IND_PASS),
                "operand_fetch_note": "DFMA with 3 register operands issues every 3.06 cycles, not 2 "
                                      "(profiles/microbench/fp64_operands.cu)",
                "peak_source": "measured live: DFMA probe kernel (nfh_probe_fp64), 2 flop/DFMA", "ms_per_step": per_step["freq"]}
            roofs["lkl_batch"] = {
                "kernel": "lkl_tile_products", "bound": "fp64", "achieved": lkl_tf, "peak": fp64_probe / 1e12,
                "unit": "TFLOP/s", "frac": lkl_tf / (fp64_probe / 1e12), "frac_of_arithmetic": lkl_tf / (fp64_arith / 1e12),
                "traffic": t_lkl, "peak_source": "measured live: DFMA probe kernel", "ms_per_step": per_step["lkl_batch"],
                "flop_note": "flops counted at the table tier of kappa (13 instructions); tiles in the polynomial tier "
                             "execute fewer, so this over-states their work"}
        dominant = max(roofs, key=lambda k: roofs[k]["ms_per_step"])
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args, N_total),
                       "step": ("one fixed-parameter EM iteration = forward-backward E-step with posteriors"
                                if args.fixed else
                                "one EM iteration = E-step + lockstep BFGS(F,alpha) + freq EM + emission refresh"),
                       "l2": f"inputs per step ({(40 if not args.fixed else 16) * rank_units / 1e9:.1f} GB per GPU) "
                             f"exceed the 126 MB L2",
                       "parallelism": f"individuals sharded x{world}, sites sharded x{world} for the freq stage",
                       "exchange": exchange,
                       "legs": "device-timed and end-to-end legs both start from reset state + the same warm-up, after one "
                               "discarded pass over the same iterations"},
            "roofline": roofs[dominant], "roofline_estep": roof_estep,
            "kernel_ms_per_step": per_step,
            "fp64_peak": fp64_note,
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                    "note": "GL upload and final posterior/path download are one-off per run, see e2e_full_run"},
            "e2e_full_run": {"seconds": t_full, "iterations": 3 * (args.steps + args.warmup),
                             "covers": "context creation, synthetic GL generation on the GPU, nfh_upload_gl, the settling "
                                       "pass and both timed legs with their warm-ups, emission refresh with e0, Viterbi, "
                                       "path and posterior download", "generate_s": t_gen, "gl_upload_s": t_upload,
                             "gl_bytes": gl_bytes, "final_refresh_viterbi_download_s": t_final_ms / 1e3},
            "viterbi": {"kernel_ms": vit_ms, "ind_sites_per_s": float(N_total) * S / (vit_ms * 1e-3) if vit_ms > 0 else None,
                        "refresh_with_e0_s": t2 - t1, "decode_and_path_download_s": t3 - t2,
                        "posterior_download_s": t4 - t3},
            "gpu_launches": int(launches), "clocks": clocks,
            "final_loglkl_rank0": float(np.sum(lk)), "viterbi_ibd_fraction_rank0": float(path.mean()),
        }
        if not args.fixed:
            out["roofline_freq"] = roofs["freq"]
            out["roofline_lkl_batch"] = roofs["lkl_batch"]
            out["bfgs"] = {"objective_evals_per_step": evals / args.steps, "rounds_per_step": rounds / args.steps}
        if parity is not None:
            out["parity_check"] = parity
        if rank_trace is not None:
            out["rank_trace"] = rank_trace
        if not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = reference_arm(args, as_cpu_baseline=True)
            except Exception as ex:  # noqa: BLE001
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                       "sample": f"unavailable: {ex}"}
        print(json.dumps(out), flush=True)
    del post, path
    barrier()                                   # peers may still gather from this rank's windows
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
