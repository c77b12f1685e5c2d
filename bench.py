#!/usr/bin/env python
"""bench.py - individual-sites per second per EM iteration on B200.

One "step" = one full EM iteration of the hot path on a fixed synthetic input
(the reference's iter_EM, EM.cpp:139-289): fused forward-backward E-step with
posteriors, lockstep L-BFGS-B update of every individual's (F, alpha) around
batched objective launches, per-site allele-frequency EM fused with the
emission refresh.  Workload at N GPUs: BASELINE.json configs[1] per GPU
(100 individuals x 1,000,000 sites, depth-2 GLs, --freq_est 1, start values
--freq 0.1 --indF 0.1,0.2) => weak scaling: N*100 individuals.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every
field is obtained.
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
IND_PER_GPU = 100
N_SITES = 1_000_000
START_FREQ, START_F, START_ALPHA = 0.1, 0.1, 0.2
# algorithmic work per unit, stated in DESIGN.md "Measurement"
ESTEP_BYTES_PER_IND_SITE = 24.0          # read 2 emissions, write 1 posterior
FREQ_FLOPS_PER_IND_PASS = 13.75          # odds-form est_maf contribution: 8 FP64 instructions (5.75 of them FMA)
FREQ_INSTR_PER_IND_PASS = 8.0            # -> at most 13.75/16 of the DFMA peak; passes per site are counted by the kernel
LKL_FLOPS_PER_IND_SITE_POINT = 14.0      # factored 2x2 update: 2 ADD + 4 FMA + 4 MUL per objective point and site
EXP_FLOPS = 21.0                         # kappa = expm1(alpha d): 13 FP64 instructions, 8 of them FMA; 3 per 5 points


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n_ind", type=int, default=IND_PER_GPU, help="individuals per GPU")
    ap.add_argument("--n_sites", type=int, default=N_SITES)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--cpu_sites", type=int, default=4000, help="sites of the bounded CPU-baseline sample")
    return ap.parse_args()


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
    path = os.path.join(ROOT, "profiles", "ncu_traffic_r01f.json")
    try:
        with open(path) as fh:
            return json.load(fh)["bytes_per_ind_site"]
    except Exception:  # noqa: BLE001
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
def reference_arm(args, as_cpu_baseline=False):
    """The UNMODIFIED reference (oracle/_ref, built from /root/reference) timed on host cores: one step = one
    iter_EM() on a bounded sample of the workload (same generator, same start values)."""
    from _oracle import Ref
    import ngsf_hmm_b200  # noqa: F401
    from ngsf_hmm_b200 import sim
    cores = os.cpu_count() or 1
    N, S = args.n_ind, args.cpu_sites
    d = sim.simulate_torch(N, S, device="cpu", seed=1002, site_chunk=1 << 14)
    threads = min(cores, N)
    ref = Ref()
    st = ref.state(d["log_gl"].numpy(), d["dist_mb"], START_FREQ, START_F, START_ALPHA, freq_est=1, n_threads=threads)
    steps, warm = (1, 0) if as_cpu_baseline else (args.steps, min(args.warmup, 1))
    for _ in range(warm):
        st.iter_EM()
    t0 = time.perf_counter()
    for _ in range(steps):
        st.iter_EM()
    dt = (time.perf_counter() - t0) / steps
    st.close()
    value = N * S / dt
    cb = {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
          "sample": f"{N} individuals x {S} sites of the same synthetic workload, {steps} iter_EM() call(s) of the "
                    f"unmodified reference in-process (oracle/_ref), --n_threads {threads} of {cores} host cores; "
                    f"its frequency loop is serial"}
    if as_cpu_baseline:
        return cb
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[1] sample: {N} ind x {S} sites, --freq_est 1, --freq 0.1 --indF 0.1,0.2"},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ---------------------------------------------------------------------------
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
    from ngsf_hmm_b200 import sim

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed): synthetic GL for this rank's SITE block, all individuals -> pinned host -> upload
    t_setup = time.perf_counter()
    ctx = nfh.Context(N_total, S, device=local_rank, n_ranks=n_ranks, rank=rank)
    gen = sim.simulate_torch(N_total, S, device=dev, seed=1002, site_begin=ctx.site_begin,
                             site_end=ctx.site_begin + ctx.sites_owned)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_setup
    t_up0 = time.perf_counter()
    ctx.upload_gl(gen["log_gl"])
    ctx.upload_pos_dist(gen["dist_mb"])
    t_upload = time.perf_counter() - t_up0
    gl_bytes = gen["log_gl"].numel() * 8
    del gen
    torch.cuda.empty_cache()

    runner = nfh.EmRank(ctx, freq_est=1)
    exchange = "none (1 rank)"
    if world > 1:
        if os.environ.get("NFH_PEER_DIRECT", "1") != "0":
            runner.enable_peer_direct()
            exchange = "fused: kernels store into peer windows over NVLink (CUDA IPC)"
        else:
            exchange = "NCCL all-to-all"
    n_own = ctx.n_ind_owned

    def reset_state():
        ctx.set_freq(np.full(ctx.sites_owned, START_FREQ))
        F = np.full(n_own, START_F); a = np.full(n_own, START_ALPHA)
        ctx.set_ind_params(F, a)
        runner.refresh_emissions()
        return F, a

    F, a = reset_state()
    for _ in range(args.warmup):
        runner.iteration(F, a, want_freq=True)       # also page-locks the frequency download buffer (one-off)

    # ---- device-timed region: K successive EM iterations, state resident in HBM
    ctx.timing(True)
    ctx.timing_read(reset=True)
    ctx.freq_passes(reset=True)
    launches0 = ctx.kernel_launches
    evals0, rounds0 = runner.total_evals, runner.total_rounds
    ext = torch.cuda.ExternalStream(ctx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)
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

    # ---- end-to-end region: the same K iterations through the host-buffer API, wall clock,
    #      parameters uploaded and lkl / F / alpha / freq downloaded every step
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lk, fr = runner.iteration(F, a, want_freq=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    h2d = 2 * n_own * 8 + (evals / max(args.steps, 1)) * (4 + 8 + 8)
    d2h = n_own * 8 + ctx.sites_owned * 8 + (evals / max(args.steps, 1)) * 8

    # ---- the E-step on its own (products + carries + apply): inside an EM iteration its forward products
    #      ride on the first objective round, so the per-step "estep" time covers carries + apply only
    ctx.timing(True)
    ctx.timing_read(reset=True)
    for _ in range(5):
        ctx.estep()
    estep_alone_ms = ctx.timing_read(reset=True)["estep"][0] / 5.0
    ctx.timing(False)

    # ---- one-off costs of a whole run, reported beside the per-iteration numbers
    t1 = time.perf_counter()
    runner.refresh_emissions(with_e0=True)
    path = ctx.viterbi()
    post = ctx.get_posterior()
    t_final = time.perf_counter() - t1
    fp64_peak = ctx.probe_fp64()

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(times[0]), float(times[1])

    if rank == 0:
        units = float(N_total) * S * args.steps
        value = units / (dev_ms * 1e-3)
        hbm_peak, peak_src = measured_peaks()
        per_step = {k: v[0] / args.steps for k, v in fam.items()}
        rank_units = float(n_own) * S                       # recursion-side units of this rank per step
        freq_units = float(N_total) * ctx.sites_owned       # frequency-side units of this rank per step
        estep_gbs = ESTEP_BYTES_PER_IND_SITE * rank_units / (estep_alone_ms * 1e-3) / 1e9
        passes_per_site = site_passes / max(ctx.sites_owned * args.steps, 1)
        freq_tf = FREQ_FLOPS_PER_IND_PASS * passes_per_site * freq_units / (per_step["freq"] * 1e-3) / 1e12
        evals_step = evals / args.steps
        lkl_flops = (LKL_FLOPS_PER_IND_SITE_POINT + 0.6 * EXP_FLOPS) * evals_step * S
        lkl_tf = lkl_flops / (per_step["lkl_batch"] * 1e-3) / 1e12 if per_step["lkl_batch"] > 0 else 0.0
        dominant = max(("estep", "lkl_batch", "freq"), key=lambda k: per_step[k])
        traffic = ncu_traffic() or {}
        t_estep = traffic.get("estep") and traffic["estep"] * rank_units
        t_freq = traffic.get("freq") and traffic["freq"] * freq_units
        t_lkl = traffic.get("lkl_batch_per_group_site") and traffic["lkl_batch_per_group_site"] * rank_units  # a round with every individual active
        roof_estep = {"kernel": "estep (tile_products + carries + apply)", "bound": "hbm", "achieved": estep_gbs,
                      "peak": hbm_peak, "unit": "GB/s", "frac": estep_gbs / hbm_peak, "traffic": t_estep,
                      "peak_source": peak_src, "ms_standalone": estep_alone_ms, "ms_per_step": per_step["estep"],
                      "note": "achieved/frac are of the stand-alone E-step (products + carries + apply); inside an EM "
                              "iteration the forward products are shared with the first objective round "
                              "(nfh_estep_with_batch) and ms_per_step covers carries + apply"}
        roof_freq = {"kernel": "freq_emission_warp", "bound": "fp64", "achieved": freq_tf,
                     "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": freq_tf / (fp64_peak / 1e12),
                     "traffic": t_freq, "passes_per_site": passes_per_site,
                     "bound_note": "FP64 CUDA-core pipe (DFMA), not tensor cores: the work is per-individual rational "
                                   "functions with no dense contraction",
                     "instruction_mix_ceiling": FREQ_FLOPS_PER_IND_PASS / (2 * FREQ_INSTR_PER_IND_PASS),
                     "operand_fetch_note": "DFMA with 3 register operands issues every 3.06 cycles, not 2 "
                                           "(profiles/microbench/fp64_operands.cu)", "peak_source": "measured live: DFMA probe kernel (nfh_probe_fp64), 2 flop/DFMA",
                     "ms_per_step": per_step["freq"]}
        roof_lkl = {"kernel": "lkl_tile_products", "bound": "fp64", "achieved": lkl_tf, "peak": fp64_peak / 1e12,
                    "unit": "TFLOP/s", "frac": lkl_tf / (fp64_peak / 1e12), "traffic": t_lkl,
                    "peak_source": "measured live: DFMA probe kernel", "ms_per_step": per_step["lkl_batch"]}
        roofs = {"estep": roof_estep, "freq": roof_freq, "lkl_batch": roof_lkl}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[1] per GPU: {args.n_ind} individuals x {S} sites, depth-2 GL, "
                                   f"--freq_est 1, start --freq 0.1 --indF 0.1,0.2; {N_total} individuals total",
                       "step": "one EM iteration = E-step + lockstep BFGS(F,alpha) + freq EM + emission refresh",
                       "l2": "inputs per step (GL+emission+posterior = 4.0 GB per GPU) exceed the 126 MB L2",
                       "parallelism": f"individuals sharded x{world}, sites sharded x{world} for the freq stage",
                       "exchange": exchange},
            "roofline": roofs[dominant], "roofline_estep": roof_estep, "roofline_freq": roof_freq,
            "roofline_lkl_batch": roof_lkl,
            "kernel_ms_per_step": per_step,
            "bfgs": {"objective_evals_per_step": evals_step, "rounds_per_step": rounds / args.steps},
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                    "note": "GL upload and final posterior/path download are one-off per run, see run_overheads"},
            "run_overheads": {"gl_upload_s": t_upload, "gl_bytes": gl_bytes, "generate_s": t_gen,
                              "viterbi_posterior_download_s": t_final},
            "gpu_launches": int(launches), "clocks": clocks,
            "final_loglkl_rank0": float(np.sum(lk)), "viterbi_ibd_fraction_rank0": float(path.mean()),
        }
        if not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = reference_arm(args, as_cpu_baseline=True)
            except Exception as ex:  # noqa: BLE001
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                       "sample": f"unavailable: {ex}"}
        print(json.dumps(out), flush=True)
    del post, path
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
