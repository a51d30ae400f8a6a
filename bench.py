#!/usr/bin/env python
"""
bench.py — headline metric of BASELINE.json: log-posterior + gradient evaluations per second
on workload H (the S3 Gaussian location-scale GAMLSS, n = 1e6 rows, two cubic P-splines with 20
coefficients each + two intercepts, P = 42) with 1024 chains per GPU.

One "step" = one batched logp+grad call: 1024 chains evaluated against the one resident copy of
the data. One evaluation = (log-posterior, full gradient) of ONE chain over ALL n rows, so a step
counts 1024 evaluations per GPU. Multi-GPU: chains are sharded (each rank holds a replica of the
44 MB data set and its own 1024 chains), no data-path collective -> "scaling": "weak".

The same JSON line carries, beside the headline keys:
  "ess"        (N = 1) the second half of the metric: NUTS-within-Gibbs on H (reference NUTSKernel
               defaults, Stan-style warm-up), min bulk-ESS of the posterior draws / posterior
               seconds, split-R-hat, SM clock over the run, and a CPU arm (the same sampler on the
               oracle's CPU port, host cores stated);
  "secondary"  other BASELINE.json configs through the same C ABI: at N = 1 the per-config
               timings S2 / S3 (+ IWLS pass) / S4 shard / S5; at N > 1 H with the 1024 chains
               STRONG-scaled over the ranks and S4 with the data ROWS sharded (1.25e6 rows per
               rank) and the [C x (P+1)] partials summed inside the C call over NCCL
               (lslb_logp_grad_nccl_f32) and over NVLink peer memory (lslb_logp_grad_peer_f32).

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...   # CPU arm: the oracle's port of the reference algorithm
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "logp_grad_evals_per_sec"
UNIT = "evals/s"
F_ROW = 16 * 2 + 14   # 16 flop per cubic smooth (fwd+bwd) x 2 + 14 (Normal loc-scale)
D_ROW = 4 + 20 * 2    # bytes per row: y + (int32 start + 4 weights) per smooth


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU")
    ap.add_argument("--cpu-chains", type=int, default=64, help="chains in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the per-config / sharded entries")
    ap.add_argument("--no-ess", action="store_true", help="skip the NUTS ESS/sec run")
    # ESS run: the metric's 1024 chains on the asynchronous driver, ~2.5 min on the GPU (warm-up 73 s,
    # posterior 76 s); --ess-chains 512 halves it. Longer runs are committed under profiles/
    ap.add_argument("--ess-chains", type=int, default=1024)
    ap.add_argument("--ess-driver", default="async", choices=["async", "lockstep"])
    ap.add_argument("--ess-warmup", type=int, default=650, help="warm-up duration (Stan-style epochs, goose/warmup.py)")
    ap.add_argument("--ess-term", type=int, default=500,
                    help="length of the final fast-adaptation epoch (reference default 50; its own kernel tests "
                         "use long ones so that the averaged dual-averaging step size settles near the target)")
    ap.add_argument("--ess-samples", type=int, default=800,
                    help="posterior draws per chain: split-R-hat < 1.01 needs ~50 effective draws per half chain")
    ap.add_argument("--ess-cpu-chains", type=int, default=8)
    ap.add_argument("--ess-cpu-seconds", type=float, default=25.0)
    return ap.parse_args()


def workload_name(n, chains):
    return (f"H: Gaussian location-scale GAMLSS (S3 model), n={n}, 2 cubic P-splines k=20 + "
            f"2 intercepts (P=42), {chains} chains per GPU")


# ----------------------------------------------------------------------------------------------
# CPU arm
# ----------------------------------------------------------------------------------------------
def cpu_sample(spec, centre, chains, seconds=12.0, min_repeats=3):
    """Times the oracle's port of the reference algorithm on a bounded sample: `chains` chains on
    the full n rows, all host threads."""
    import torch

    from liesel_b200 import synth
    from oracle.cpu_port import CpuPort

    torch.set_num_threads(os.cpu_count() or 1)
    port = CpuPort(spec)
    th, aux = synth.evaluation_points(spec, centre, chains, seed=778)
    t_one = port.time(th, aux, repeats=1, warmup=1)
    repeats = max(min_repeats, int(seconds / max(t_one, 1e-6)))
    repeats = min(repeats, 200)
    t0 = time.perf_counter()
    for _ in range(repeats):
        port.logp_grad(th, aux)
    dt = (time.perf_counter() - t0) / repeats
    return {
        "value": chains / dt, "unit": UNIT, "cores": port.threads, "kind": "port",
        "sample": (f"{chains} chains x n={spec.n} rows per call (dense-basis GEMM formulation of the "
                   f"reference, torch CPU fp32, {port.threads} threads), {repeats} calls, "
                   f"{dt * 1e3:.1f} ms/call"),
    }, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from liesel_b200 import synth

    spec, centre = synth.s3_gamlss(n=args.n)
    import torch

    from oracle.cpu_port import CpuPort

    torch.set_num_threads(os.cpu_count() or 1)
    port = CpuPort(spec)
    th, aux = synth.evaluation_points(spec, centre, args.cpu_chains, seed=778)
    for _ in range(max(args.warmup, 1)):
        port.logp_grad(th, aux)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.logp_grad(th, aux)
    dt = (time.perf_counter() - t0) / args.steps
    value = args.cpu_chains / dt
    sample = (f"each step = {args.cpu_chains} chains x n={args.n} rows (bounded sample of the "
              f"{args.chains}-chain workload), torch CPU fp32, {port.threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.chains), "n": args.n,
                   "chains_per_gpu": args.chains, "cpu_sample_chains": args.cpu_chains},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": port.threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("liesel 0.5.1 needs Python>=3.13 + JAX/TFP/BlackJAX, none installable here; this arm "
                 "times oracle/cpu_port.py, the CPU port of the reference algorithm"),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def wait_started(self, timeout=5.0):
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.02)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            # samples taken while the GPU was actually loaded: power above the midpoint of the range
            thr = 0.5 * (min(pw) + max(pw))
            load = [c for c, w in zip(sm, pw) if w >= thr] or sm
            out = {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)),
                   "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load),
                   "power_w_max": float(max(pw))}
        return out


def measure_pipe_peaks(torch, _lib, sm_count):
    """FP32 FMA (scalar and packed f32x2), MUFU and LDS peaks measured live on this GPU."""
    import ctypes as C

    sink = torch.zeros(4, device="cuda")
    res = {}
    for kind, name, iters in [(0, "ffma", 4000), (1, "ffma2", 4000), (7, "ffma_3reg", 4000), (2, "mufu_ex2", 2000)]:
        work = C.c_double(0.0)
        best = None
        def go():
            _lib.check(_lib.bench_lib().lslb_debug_pipe_peak(kind, sm_count * 8, iters, _lib.ptr(sink),
                                                             C.byref(work), _lib.stream_ptr()), "pipe_peak")
        for rep in range(10):  # best of nine after one discarded repetition
            # the timed launch is queued BEHIND an untimed one, so the start event is stamped when the
            # stream is busy and the timed kernel follows it without a host-side gap (under torchrun the
            # ranks' host threads compete and a late launch used to be counted as kernel time)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            go()
            s.record()
            go()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e)
            if rep > 0:
                best = ms if best is None else min(best, ms)
        res[name] = work.value / (best * 1e-3)
    return {
        "fp32_fma_tflops": 2 * res["ffma"] / 1e12,
        "fp32_ffma2_tflops": 2 * res["ffma2"] / 1e12,
        "mufu_ex2_tops": res["mufu_ex2"] / 1e12,
        "fp32_fma_3_distinct_registers_tflops": 2 * res["ffma_3reg"] / 1e12,
    }



# ----------------------------------------------------------------------------------------------
# secondary entries and the ESS run
# ----------------------------------------------------------------------------------------------
def timed_calls(torch, dist, fn, reps, flush, world, dev, warm=3):
    """Mean ms per call over `reps` calls (L2 flushed before each, CUDA events on the current
    stream, barrier + synchronize on both sides), max over ranks."""
    for _ in range(warm):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# algorithmic flops per batched call (C evaluations) of the secondary configs
def _flops(name, n, C, p=0):
    per_row = {"S1": 4 * 10 + 14, "S2": 16 * 5 + 12, "S3": 16 * 2 + 14, "S3_iwls": 8 + 8 + 20 + 8 + 14,
               "S4": 4 * p + 12, "S5": 4 * 8 + 10}[name]
    return float(n) * C * per_row


def s3_iwls_sampler(torch, spec, plan, params, aux, warmup=200, samples=300):
    """BASELINE config 3 as a SAMPLER: an IWLS kernel per coefficient block + Gibbs tau2 (what
    ``dist_reg_mcmc`` configures, model/distreg.py:300-349), spline blocks in the sum-to-zero coordinates."""
    from liesel_b200 import diagnostics, reparam
    from liesel_b200.iwls_sampler import IWLSGibbsSampler

    cons = {b.name: reparam.sum_to_zero_basis(plan.column_means(b.name)) for b in spec.blocks if b.kind == 1}
    smp = IWLSGibbsSampler(plan, params, aux, seed=1, constraints=cons)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    smp.warmup(warmup)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    draws, _, infos = smp.sample(samples)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ess = diagnostics.ess_bulk_device(draws)
    acc = [float(np.mean([float(i[b.name]["acceptance_prob"].mean()) for i in infos])) for b in smp.blocks]
    return {"chains": int(draws.shape[1]), "warmup_iters": warmup, "posterior_iters": samples,
            "iterations_per_s": samples / (t2 - t1), "posterior_seconds": t2 - t1, "warmup_seconds": t1 - t0,
            "ess_bulk_min": float(ess.min()), "split_rhat_max": float(np.max(diagnostics.split_rhat_device(draws))),
            "ess_per_sec_posterior": float(ess.min() / (t2 - t1)), "mean_accept_min_over_blocks": min(acc),
            "evals": dict(smp.evals),
            "what": "IWLS transition per block (2 score+information+Cholesky passes + 1 log-posterior) + Gibbs tau2"}


def secondary_single_gpu(torch, dist, dev, flush, fp32_peak, tf32x3_note, reps=10):
    """S1, S2, S3 (+ one IWLS pass), one S4 shard and S5 (+ IWLS passes of both blocks) at BASELINE.json's sizes on one GPU."""
    from liesel_b200 import synth
    from liesel_b200.plan import Plan

    out = []
    cfgs = [
        ("S1", lambda: synth.s1_blr(n=1000), 4, None, "Bayesian linear regression, n=1000, p=10, 4 chains (launch-latency bound: one launch per call)"),
        ("S2", lambda: synth.s2_pspline_gam(n=100_000), 64, "loc_np0", "P-spline GAM, 5 cubic smooths x 20, n=1e5, 64 chains"),
        ("S3", lambda: synth.s3_gamlss(n=1_000_000), 256, "loc_np0", "Gaussian location-scale GAMLSS, n=1e6, 256 chains"),
        ("S4", lambda: synth.s4_logistic(n=10_000_000, world_size=8), 1024, None,
         "logistic regression p=256: ONE rank's shard of n=1e7 (1.25e6 rows), 1024 chains, no collective at N=1"),
        ("S5", lambda: synth.s5_poisson_ri(n=5_000_000, G=100_000), 1024, "log_rate_p0",
         "Poisson random intercept, G=1e5, n=5e6, 1024 chains"),
    ]
    for name, make, C, iwls_block, what in cfgs:
        spec, centre = make()
        plan = Plan(spec, device=dev)
        th, aux = synth.evaluation_points(spec, centre, C)
        t = torch.as_tensor(th, device=dev)
        a = torch.as_tensor(aux, device=dev) if spec.num_aux else None
        ms = timed_calls(torch, dist, lambda: plan.logp_grad(t, a), reps, flush, 1, dev)
        ent = {"name": name, "workload": what, "n": spec.n, "chains": C, "P": plan.P,
               "kernel": plan.kernel_name(C), "variant": plan.variant, "ms_per_call": ms,
               "evals_per_s": C / (ms * 1e-3)}
        p_dense = plan.P if name == "S4" else 0
        fl = _flops(name, spec.n, C, p_dense)
        ent["algorithmic_flops_per_call"] = fl
        ent["achieved_tflops"] = fl / (ms * 1e-3) / 1e12
        if name == "S4":
            ent["roofline_note"] = tf32x3_note
        else:
            ent["frac_of_fp32_peak"] = ent["achieved_tflops"] / fp32_peak
        if iwls_block:
            ms_i = timed_calls(torch, dist, lambda: plan.iwls(iwls_block, t, a), reps, flush, 1, dev)
            ent["iwls"] = {"block": iwls_block, "ms_per_pass": ms_i, "passes_per_s": C / (ms_i * 1e-3),
                           "what": "score + observed information + ridge + Cholesky of one coefficient block"}
            if name == "S3":
                fi = _flops("S3_iwls", spec.n, C)
                ent["iwls"]["frac_of_fp32_peak"] = fi / (ms_i * 1e-3) / 1e12 / fp32_peak
                ent["iwls_sampler"] = s3_iwls_sampler(torch, spec, plan, t, a)
            if name == "S5":  # the random-intercept block: diagonal information (lslb_iwls_group)
                ms_g = timed_calls(torch, dist, lambda: plan.iwls_group(t, a), reps, flush, 1, dev)
                ent["iwls_group"] = {"block": "log_rate_ri0", "ms_per_pass": ms_g, "passes_per_s": C / (ms_g * 1e-3),
                                     "what": "score + diagonal of the observed information of the random-intercept block (G = 1e5)"}
        out.append(ent)
        del plan, t, a
        torch.cuda.empty_cache()
    return out


def secondary_multi_gpu(torch, dist, dev, flush, rank, world, K, h_plan, h_spec, h_centre, total_chains=1024):
    """N > 1: (1) H strong-scaled (1024 chains in total, data replicated, no collective);
    (2) S4 with the ROWS sharded, 1.25e6 rows per rank (n = 1e7 at N = 8, BASELINE config 4): every
    rank evaluates all 1024 chains on its rows and the [C x (P+1)] partials are summed over the ranks
    inside the C call, over NCCL and over NVLink peer memory."""
    from liesel_b200 import dist as ld
    from liesel_b200 import synth
    from liesel_b200.plan import Plan

    out = []
    # -- (1) strong scaling of the headline -----------------------------------------------------
    a0, a1 = ld.chain_shard(total_chains, rank, world)
    th, aux = synth.evaluation_points(h_spec, h_centre, total_chains, seed=777)
    t = torch.as_tensor(th[a0:a1], device=dev).contiguous()
    a = torch.as_tensor(aux[a0:a1], device=dev).contiguous()
    lp = torch.empty(a1 - a0, dtype=torch.float32, device=dev)
    g = torch.empty((a1 - a0, h_plan.P), dtype=torch.float32, device=dev)
    ms = timed_calls(torch, dist, lambda: h_plan.logp_grad(t, a, out=(lp, g)), K, flush, world, dev)
    out.append({"name": "H_strong_scaling", "scaling": "strong",
                "workload": f"H (n={h_spec.n}) with {total_chains} chains IN TOTAL, {a1 - a0} per GPU, data replicated",
                "chains_total": total_chains, "chains_per_gpu": a1 - a0, "kernel": h_plan.kernel_name(a1 - a0),
                "ms_per_call": ms, "evals_per_s": total_chains / (ms * 1e-3), "collective": None})
    del t, a, lp, g

    # -- (2) S4 row-sharded -------------------------------------------------------------------
    rows_per_rank, p, C = 1_250_000, 256, 1024
    n_total = rows_per_rank * world
    spec, centre = synth.s4_logistic(n=n_total, p=p, rank=rank, world_size=world)
    plan = Plan(spec, device=dev)
    th, _ = synth.evaluation_points(spec, centre, C)
    t = torch.as_tensor(th, device=dev)
    flags = 0 if rank == 0 else ld.SKIP_PRIOR
    base = {"workload": (f"S4 logistic regression, n={n_total} rows sharded over {world} GPUs "
                         f"({rows_per_rank} per rank), p={p}, {C} chains on every rank"),
            "scaling": "weak in rows (BASELINE config 4 at N=8)", "n_total": n_total, "chains": C,
            "kernel": plan.kernel_name(C), "variant": plan.variant,
            "exchange_bytes_per_call": C * (p + 1) * 4}
    # rank-local partial sums, then the reference sum over ranks through torch.distributed
    ms_local = timed_calls(torch, dist, lambda: plan.logp_grad(t, None, flags=flags), K, flush, world, dev)
    lp_l, g_l = plan.logp_grad(t, None, flags=flags)
    ref = torch.cat([lp_l[:, None], g_l], dim=1).double()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    out.append(dict(base, name="S4_rowshard_local_only", transport=None, ms_per_call=ms_local,
                    note="each rank's partial evaluation without the sum over ranks (max over ranks)"))
    for name, make in (("nccl", lambda: ld.NcclComm(dev)), ("peer", lambda: ld.PeerGroup(C * (p + 1), dev))):
        try:
            sh = make()
            ms_t = timed_calls(torch, dist, lambda: plan.logp_grad(t, None, flags=flags, shard=sh), K, flush, world, dev)
            lp_s, g_s = plan.logp_grad(t, None, flags=flags, shard=sh)
            got = torch.cat([lp_s[:, None], g_s], dim=1).double()
            err = (got - ref).abs().max()
            # relative to the largest entry of the same column (log-posterior / each gradient column):
            # element-wise ratios explode where a gradient entry cancels to ~0 across the ranks
            rel = ((got - ref).abs().amax(dim=0) / ref.abs().amax(dim=0).clamp_min(1e-30)).max()
            stat = torch.stack([err, rel])
            dist.all_reduce(stat, op=dist.ReduceOp.MAX)
            ent = dict(base, name=f"S4_rowshard_{name}", transport=name,
                       entry_point=f"lslb_logp_grad_{name}_f32", ms_per_call=ms_t,
                       evals_per_s=C / (ms_t * 1e-3), exchange_ms=max(ms_t - ms_local, 0.0),
                       max_abs_diff_vs_rank_local_sums=float(stat[0]), max_rel_diff_per_column=float(stat[1]))
            if name == "peer":
                ent["timed_out"] = bool(sh.timed_out())
            out.append(ent)
            torch.cuda.synchronize()
            dist.barrier()
            sh.close()
        except Exception as e:  # a transport that cannot be set up is reported, not fatal
            out.append(dict(base, name=f"S4_rowshard_{name}", transport=name, error=repr(e)[:300]))
    del plan, t
    torch.cuda.empty_cache()
    return out


def run_ess(args, torch, dev, plan, spec, centre, local):
    """NUTS-within-Gibbs on H: all coefficients in one NUTS block (reference NUTSKernel defaults:
    max_treedepth 10, target acceptance 0.8, diagonal mass matrix, Stan-style warm-up epochs), tau2
    by its Gibbs full conditional after every transition; both smooths under the sum-to-zero
    constraint (liesel_b200/reparam.py). ESS = min over coordinates of bulk-ESS of all posterior draws.
    Driver: the asynchronous one (liesel_b200/nuts_async.py: every chain walks its own trees, the batch
    shares one evaluation per step) unless --ess-driver lockstep."""
    from liesel_b200 import reparam, synth
    from liesel_b200.nuts_gibbs import AsyncNutsGibbs, NutsGibbs, PlanEvaluator, summarise, summarise_async

    C, W, S = args.ess_chains, args.ess_warmup, args.ess_samples
    cm = {b.name: plan.column_means(b.name) for b in spec.blocks if b.kind == 1}
    rep = reparam.sum_to_zero(spec, cm, device=dev)
    th0, aux0 = synth.evaluation_points(spec, centre, C, seed=1)
    sampler = ClockSampler(local)
    sampler.wait_started()
    common = ("one NUTS block over all coefficients + Gibbs draw of every tau2 per transition; reference "
              "defaults (max_treedepth 10, target accept 0.8, diagonal mass matrix, Stan-style warm-up epochs "
              f"with a final fast epoch of {args.ess_term}); ")
    if args.ess_driver == "async":
        run = AsyncNutsGibbs(spec, PlanEvaluator(plan), torch.as_tensor(th0, device=dev),
                             torch.as_tensor(aux0, device=dev), W, S, reparam=rep, seed=1,
                             epoch_kw=dict(term_duration=args.ess_term))
        timing = run.run()
        clocks = sampler.stop()
        out = summarise_async(run, timing)
        q_state, step_state, mass_state = run.nuts.st["q"], run.nuts.st["step"], run.nuts.st["inv_mass"]
        out.update({
            "sampler": common + "asynchronous driver: every chain walks its own trees, one batched evaluation "
                                "per step shared by all chains (lslb_nuts_async_step)",
            "mean_batch_fill": 1.0,
        })
    else:
        t0 = time.perf_counter()
        run = NutsGibbs(spec, PlanEvaluator(plan), torch.as_tensor(th0, device=dev),
                        torch.as_tensor(aux0, device=dev), reparam=rep, seed=1, compact=True)
        winfo = run.warmup(W, term_duration=args.ess_term)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        draws, infos = run.sample(S)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        clocks = sampler.stop()
        out = summarise(draws, infos, t2 - t1, t1 - t0)
        nuts = run.nuts
        q_state, step_state, mass_state = nuts.q, nuts.step_size, nuts.inv_mass
        out.update({
            "sampler": common + "lock-step driver, finished chains dropped from the batched evaluation",
            "warmup_mean_treedepth": float(np.mean([float(i.treedepth.float().mean()) for i in winfo])),
            "batched_evals_warmup": int(sum(i.evaluations for i in winfo)),
            "step_size_median": float(nuts.step_size.median()),
            "mean_batch_fill": float(nuts.chain_evals) / max(1.0, float(nuts.evals) * C),
        })
    out.update({
        "driver": args.ess_driver,
        "model": workload_name(spec.n, C).replace(" per GPU", ""),
        "constraint": "sum-to-zero on both smooths (beta = Z gamma, 2 x 19 + 2 = 40 sampled coordinates)",
        "warmup_iters": W, "warmup_term_duration": args.ess_term, "posterior_iters": S,
        "clocks": clocks,
    })
    # ---- for comparison on the same model and chain count: the IWLS-within-Gibbs scheme (what Liesel's
    #      dist_reg_mcmc configures by default, model/distreg.py:300-349) - not the NUTS metric, reported beside it
    try:
        out["iwls_sampler_same_model"] = s3_iwls_sampler(torch, spec, plan, torch.as_tensor(th0, device=dev),
                                                         torch.as_tensor(aux0, device=dev))
    except Exception as e:  # never let the side entry take the metric down
        out["iwls_sampler_same_model"] = {"error": repr(e)[:200]}
    # ---- CPU arm: the same sampler on the oracle's CPU port, host cores ----------------------
    cpu = None
    if not args.no_cpu_baseline:
        from oracle.cpu_port import CpuEvaluator

        Cc = min(args.ess_cpu_chains, C)
        torch.set_num_threads(os.cpu_count() or 1)
        ev = CpuEvaluator(spec)
        rep_c = reparam.sum_to_zero(spec, cm)
        q_full = run.full(q_state[:Cc]).cpu()
        run_c = NutsGibbs(spec, ev, q_full, run.aux[:Cc].cpu(), reparam=rep_c, seed=2,
                          initial_step_size=1.0, fused=False)
        # adapted state of the GPU run's first Cc chains: the CPU arm times POSTERIOR transitions
        run_c.nuts.step_size = step_state[:Cc].cpu().clone()
        run_c.nuts.inv_mass = mass_state[:Cc].cpu().clone()
        ev0 = run_c.nuts.evals
        tc0 = time.perf_counter()
        iters = 0
        while iters < 1 or (time.perf_counter() - tc0 < args.ess_cpu_seconds and iters < S):
            run_c.nuts.transition()
            run_c._gibbs(run_c.nuts)
            iters += 1
        dtc = time.perf_counter() - tc0
        ess_per_draw_chain = out["ess_bulk_min"] / (S * C)
        cpu = {
            "kind": "port", "cores": ev.port.threads, "chains": Cc, "transitions_timed": iters,
            "seconds": dtc, "seconds_per_transition": dtc / iters,
            "batched_evals": int(run_c.nuts.evals - ev0),
            "ess_per_sec_posterior": ess_per_draw_chain * Cc * iters / dtc,
            "how": ("real lock-step NUTS+Gibbs transitions on oracle/cpu_port.py (torch CPU fp32, all host threads) from "
                    "the GPU run's adapted state (positions, step sizes, mass matrices of its first "
                    f"{Cc} chains); ESS per draw and chain ({ess_per_draw_chain:.3f}) is the GPU run's — same "
                    "sampler, same model — because a CPU run long enough to estimate ESS itself does not "
                    f"fit the bench budget; bounded sample ({Cc} of the {C} chains)"),
        }
    out["cpu_arm"] = cpu
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from liesel_b200 import _lib, synth
    from liesel_b200.interface import B200Interface
    from liesel_b200.plan import Plan

    C_, n, K, W = args.chains, args.n, args.steps, max(args.warmup, 3)
    spec, centre = synth.s3_gamlss(n=n)
    plan = Plan(spec, device=dev)
    iface = B200Interface(plan)
    P, A = plan.P, plan.A
    th, aux = synth.evaluation_points(spec, centre, C_, seed=777 + rank)
    d_th, d_aux = torch.as_tensor(th, device=dev), torch.as_tensor(aux, device=dev)
    logp = torch.empty(C_, dtype=torch.float32, device=dev)
    grad = torch.empty((C_, P), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()

    def step():
        plan.logp_grad(d_th, d_aux, want_grad=True, out=(logp, grad))

    # clocks are sampled from before the warm-up until the end of the e2e loop; every buffer the
    # timed loops need is allocated first so the GPU stays loaded while the sampler runs
    h_th = torch.as_tensor(th).pin_memory()
    h_aux = torch.as_tensor(aux).pin_memory()
    h_logp = torch.empty(C_, dtype=torch.float32).pin_memory()
    h_grad = torch.empty((C_, P), dtype=torch.float32).pin_memory()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_started()
    t_w = time.perf_counter()
    nwarm = 0
    while nwarm < W or time.perf_counter() - t_w < 0.5:  # >= W steps and >= 0.5 s under load
        flush.zero_()
        step()
        nwarm += 1
        if nwarm % 16 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()

    # ---- timed region: value (inputs resident in HBM) ------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in kev:  # force creation of the underlying cudaEvent_t handles
        a.record(); b.record()
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()
        _lib.lib.lslb_debug_set_events(kev[k][0].cuda_event, kev[k][1].cuda_event)
        ev[k][0].record()
        step()
        ev[k][1].record()
    _lib.lib.lslb_debug_set_events(None, None)
    torch.cuda.synchronize()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = C_ * world * K / (total_ms_max * 1e-3)

    # ---- e2e: host buffers -> public API -> host buffers --------------------------------
    keys = spec.position_keys()

    def e2e_step():
        p_dev = h_th.to(dev, non_blocking=True)
        a_dev = h_aux.to(dev, non_blocking=True)
        state = iface.unpack(p_dev, a_dev)
        lp, g = iface.log_prob_and_grad(state)
        h_logp.copy_(lp, non_blocking=True)
        h_grad.copy_(torch.cat([g[k] for k in keys], dim=1), non_blocking=True)

    for _ in range(3):
        flush.zero_()
        e2e_step()
    torch.cuda.synchronize()
    eev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    torch.cuda.synchronize()
    for k in range(K):
        flush.zero_()
        eev[k][0].record()
        e2e_step()
        eev[k][1].record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if sampler else None
    e2e_ms = float(sum(a.elapsed_time(b) for a, b in eev))
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = C_ * world * K / (float(t.item()) * 1e-3)

    secondary = []
    if world > 1 and not args.no_secondary:
        secondary = secondary_multi_gpu(torch, dist, dev, flush, rank, world, K, plan, spec, centre)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0) ---------------------------------------
    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks_file.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks_file else "fallback"
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    pipe = measure_pipe_peaks(torch, _lib, sm_count)
    k_ms = float(np.mean(kern_ms))
    flops = float(C_) * n * F_ROW
    bytes_alg = float(n) * D_ROW + 4.0 * (2 * C_ * P + C_)
    fp32_peak = max(pipe["fp32_fma_tflops"], pipe["fp32_ffma2_tflops"])
    achieved_tf = flops / (k_ms * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("eval_kernel_dram_bytes")
    except (OSError, ValueError):
        pass
    roofline = {
        "kernel": plan.kernel_name(C_) + "<FAM_NORMAL_LS, NS=2, PM=2, GRAD, UNITY> (ncu: lslb::banded_x2_kernel<0,2,2,1,1>)"
        if plan.kernel_name(C_) == "lslb::banded_x2_kernel" else plan.kernel_name(C_),
        "bound": "fp32_simt",
        "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak,
        "peak_source": "FP32 FMA micro-benchmark run live in this process (lslb_debug_pipe_peak)",
        "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": bytes_alg,
        "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (total_ms / K),
        "traffic": traffic,
        "traffic_source": "ncu --set full capture of this command, committed as profiles/traffic.json (offline, not measured in this run)",
        "hbm": {"achieved": bytes_alg / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_alg / (k_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
        "pipe_peaks_measured": pipe,
        "why": ("data is read once per call for all 1024 chains: 1070 flop/B, far above the FP32 "
                "ridge point (~11 flop/B); the binding pipe is FP32 CUDA cores, not HBM or tensor"),
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_sample(spec, centre, args.cpu_chains)
    if world == 1 and not args.no_secondary:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        bf16 = float(mp.get("bf16_tflops_sustained", mp.get("bf16_tflops", 0.0)) or 0.0)
        note = ("tensor pipe: eta = X B and G = X^T R as tcgen05 kind::tf32 GEMMs in split precision "
                "(2 + 3 passes for rtol 1e-5); executed tensor flops = 2.5 x the algorithmic 4 n p C; TF32 "
                f"dense rate = half the bf16 rate (MEASURED_PEAKS.json sustained bf16 {bf16:.0f} TFLOP/s)")
        secondary = secondary_single_gpu(torch, dist, dev, flush, fp32_peak, note)
        for ent in secondary:
            if ent["name"] == "S4" and bf16 > 0:
                ent["frac_of_tf32_split_bound"] = 2.5 * ent["achieved_tflops"] / (0.5 * bf16)
    ess = None
    if world == 1 and not args.no_ess:
        ess = run_ess(args, torch, dev, plan, spec, centre, local)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(n, C_), "n": n, "chains_per_gpu": C_, "P": P,
                   "parallelism": f"chain-shard x{world} (data replicated, no collective)",
                   "l2": "flushed between timed iterations (256 MiB write)",
                   "kernel_variant": plan.variant},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(C_ * (P + A) * 4), "d2h_bytes_per_step": int(C_ * (P + 1) * 4),
                "api": "B200Interface.log_prob_and_grad on pinned host buffers"},
        # per step: gather_small_kernel, the row-streaming kernel, finalize_rows_kernel (the log-posterior is
        # added by its last CTA per chain group) - ncu launch list: profiles/r02_launches.md
        "gpu_launches": int(3 * K),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "ess": ess,
        "secondary": secondary,
        "wall_s_timed_region_incl_flush": wall,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version
    # banner on stdout under NCCL_DEBUG=VERSION), so everything but the result line goes to stderr:
    # file descriptor 1 is pointed at stderr for the run and the line is written to the saved one.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    global print
    builtin_print = print

    def capture(*a, **k):
        if k.get("file") is None:
            lines.append(" ".join(str(x) for x in a))
        else:
            builtin_print(*a, **k)

    print = capture
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        print = builtin_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        builtin_print(ln, flush=True)


if __name__ == "__main__":
    main()
