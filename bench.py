#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native parallel-tempering hot path.

Headline workload (BASELINE.json configs[2], "C3"): PenDigit-shape 16-30-10 classifier on synthetic data,
1,000,000 train rows (+65,536 test rows), 1024 chains PER GPU (weak scaling), random-walk parallel tempering, no
surrogate. A "step" is one full MCMC step of every chain: in-kernel Philox proposal + prior, batched likelihood on
train and test, Metropolis-Hastings accept with history bookkeeping (and the exchange sweep when the step index hits
swap_interval). metric = replica x row likelihood evaluations per second (whole job).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--path auto|f32|tc]
                    [--scaling weak|strong] [--no-secondary] [--no-cpu]

One JSON line on stdout (rank 0). Besides the contract's keys it carries
  pipes        live SM-pipe micro-benchmarks (sapt_measure_pipes): the denominators of roofline.governing
  secondary    short C4 (Covtype-shape Langevin) and C5 (Chess-shape, surrogate_prob 0.5) legs, run AFTER the headline's
               timed region: ms/step, evals/s, samples/s, dominant kernel and its own roofline fraction
  strong       (N > 1) the same C3 step with 1024 chains IN TOTAL sharded over the N GPUs (BASELINE configs[2] as written)
  sharding_invariant  (N > 1) 30 steps of a small problem on N ranks equal the same seed on one rank, decision for decision
See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3 = dict(name="C3 PenDigit-shape 16-30-10 synthetic", D=16, H=30, O=10, n_train=1_000_000, n_test=65_536,
          chains_per_gpu=1024, maxtemp=4.0, swap_interval=5, seed=1234)
C4 = dict(name="C4 Covtype-shape 54-64-7 synthetic, batch-Langevin p=0.5", D=54, H=64, O=7, n_train=581_012, n_test=65_536,
          chains_per_gpu=256, maxtemp=4.0, seed=4)
C5 = dict(name="C5 Chess-shape 6-25-18 synthetic, surrogate_prob 0.5 (fixed MLP P-200-50-1)", D=6, H=25, O=18, n_train=16_833,
          n_test=11_223, chains_per_gpu=8192, maxtemp=4.0, swap_interval=4, seed=4)
METRIC = "replica_x_row_likelihood_evals_per_sec"
UNIT = "evals/s"


def synth(D, O, n, seed):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((n, D)).astype(np.float32)
    y = rs.randint(0, O, size=n).astype(np.int32)
    return x, y


# ---------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (ptReplica.likelihood_func -> Network.evaluate_proposal, _revamp.py:391-402,
# 155-174) loaded from oracle/_ref (archive written by oracle/build_ref.py), one replica per host core like the
# reference's one-process-per-replica model. Falls back to the oracle's literal per-row port only when the archive
# is absent, and says which ("kind").
# ---------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(topo, kind):
    _W["topo"], _W["kind"] = topo, kind
    if kind == "reference":
        import contextlib, io, warnings
        from oracle import ref_loader
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            ref = ref_loader.load_reference("RV")
        _W["ref"], _W["loader"] = ref, ref_loader
    else:
        from oracle import sapt_oracle as so
        _W["so"] = so


def _cpu_worker(args):
    n_rows, seed = args
    topo = _W["topo"]
    D, H, O = topo
    x, y = synth(D, O, n_rows, seed)
    data = np.concatenate([x.astype(np.float64), y[:, None].astype(np.float64)], axis=1)
    P = D * H + H * O + H + O
    w = np.random.RandomState(seed + 1).randn(P)
    if _W["kind"] == "reference":
        import contextlib, io
        ref = _W["ref"]
        with contextlib.redirect_stdout(io.StringIO()):
            net = ref.Network(topo, data, data[:1], 0.5)
            rep = _W["loader"].DuckReplica(ref, topo, 1.0)
            t0 = time.perf_counter()
            rep.likelihood_func(net, data, w)
            return time.perf_counter() - t0
    t0 = time.perf_counter()
    _W["so"].likelihood_func(data, w, topo, 1.0, rowloop=True)
    return time.perf_counter() - t0


class CpuArm:
    """Pool of one process per host core, created and warmed BEFORE anything is timed; throughput is computed from the
    workers' own in-loop clocks (slowest worker = the wall of one parallel pass), so fork / import / data synthesis
    are not in the number."""

    def __init__(self, topo):
        import multiprocessing as mp
        from oracle import ref_loader
        self.kind = "reference" if (ref_loader.reference_available() or ref_loader.archive_available()) else "port"
        self.source = ref_loader.source_kind() if self.kind == "reference" else "oracle.sapt_oracle.likelihood_func(rowloop=True)"
        self.cores = os.cpu_count() or 1
        self.topo = topo
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init, initargs=(topo, self.kind))
        self.pool.map(_cpu_worker, [(200, 50 + c) for c in range(self.cores)])   # warm-up: imports, first-call costs

    def measure(self, sample_rows, passes=1):
        worst, wall0 = 0.0, time.perf_counter()
        for p in range(passes):
            worst += max(self.pool.map(_cpu_worker, [(sample_rows, 100 + c + 1000 * p) for c in range(self.cores)]))
        wall = time.perf_counter() - wall0
        value = self.cores * sample_rows * passes / worst
        what = ("the UNMODIFIED reference's ptReplica.likelihood_func (_revamp.py:391-402) loaded from " + self.source
                if self.kind == "reference" else "the oracle's literal per-row port of the reference loop (no reference archive found)")
        return dict(value=value, unit=UNIT, cores=self.cores, kind=self.kind,
                    sample=f"{self.cores} replicas (one process per host core) x {sample_rows} synthetic rows x {passes} pass(es) of {what}; "
                           f"cost is linear in rows and replicas",
                    timing="rows / slowest worker's in-loop time per pass; pool created and warmed before timing",
                    slowest_worker_s=worst, wall_s=wall)

    def close(self):
        self.pool.close(); self.pool.join()


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    topo = [C3["D"], C3["H"], C3["O"]]
    arm = CpuArm(topo)
    sample_rows = 10000
    for _ in range(args.warmup):
        arm.measure(1000)
    per_step = []
    for _ in range(args.steps):
        per_step.append(arm.measure(sample_rows)["slowest_worker_s"])
    arm.close()
    total = arm.cores * sample_rows * args.steps
    value = total / sum(per_step)
    what = "unmodified reference ptReplica.likelihood_func from " + arm.source if arm.kind == "reference" else "oracle per-row port"
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * sum(per_step) / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=C3["name"], rows_per_step=sample_rows, replicas_per_step=arm.cores,
                            note="bounded sample of the same workload; cost is linear in rows and replicas"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=arm.cores, kind=arm.kind,
                                  sample=f"{arm.cores} processes x {sample_rows} rows per step, {what}",
                                  timing="rows / slowest worker's in-loop time per step"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.t = index, [], False, None

    def _loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.t = threading.Thread(target=self._loop, daemon=True); self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t: self.t.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------
# helpers of the GPU arm
# ---------------------------------------------------------------------------------------------
def _allmax(x, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _time_alone(fn, flush, reps):
    """Average device time of fn() with the L2 flushed before every call (CUDA events on the current stream)."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms))


def _timed_steps(eng, first, steps, flush, world):
    """K steps (+ rendezvous) timed one by one with CUDA events, L2 flushed (untimed) between them; max over ranks."""
    import torch
    import torch.distributed as dist
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    ev0 = eng.evals
    for k in range(steps):
        i = first + k
        flush.zero_()
        evs[k][0].record()
        eng.step(i); eng.rendezvous(i)
        evs[k][1].record()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    total_ms = _allmax(sum(a.elapsed_time(b) for a, b in evs), world)
    return total_ms, eng.evals - ev0


def _governing(work, pipes, sm_mhz, n_sm, evals, k_ms):
    """SM-pipe ceiling of the likelihood kernel from (a) the MUFU / FFMA operations it EXECUTES per evaluation
    (sapt_loglik_work_model: padding and shared reciprocals included) and (b) the pipe rates MEASURED on this GPU a moment ago
    (sapt_measure_pipes), rescaled to the SM clock observed during the timed region."""
    mufu_rate = min(pipes["ex2_per_clk_sm"], pipes["rcp_per_clk_sm"])
    ffma_rate = pipes["ffma_per_clk_sm"]
    clk_per_eval = max(work["mufu"] / mufu_rate, work["ffma"] / ffma_rate if work["ffma"] else 0.0)
    peak = n_sm * sm_mhz * 1e6 / clk_per_eval
    ach = evals / (k_ms * 1e-3)
    return dict(bound="sm_pipes(mufu|fp32)", achieved=ach, peak=peak, unit=UNIT, frac=ach / peak,
                mufu_per_eval_executed=work["mufu"], ffma_per_eval_executed=work["ffma"],
                mufu_per_clk_sm_measured=mufu_rate, ffma_per_clk_sm_measured=ffma_rate, sm_mhz_used=sm_mhz,
                model=f"{n_sm} SMs x {sm_mhz:.0f} MHz / max({work['mufu']:.0f} MUFU / {mufu_rate:.2f} per clk, "
                      f"{work['ffma']:.0f} FFMA / {ffma_rate:.1f} per clk)")


def _stored_traffic(kernel_key, chains):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture. A stored figure,
    not measured in this run: it is only reported while the kernel source still hashes to what was profiled."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = tj.get(kernel_key)
        if not e or e.get("chains") != chains:
            return None, "no stored ncu capture for this kernel / chain count"
        src = os.path.join(ROOT, e["source"])
        sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
        if sha != e["source_sha256"]:
            return None, f"stored ncu capture ({e['capture']}) predates the current {e['source']}: stale, not reported"
        return e["dram_bytes_per_launch"], f"stored ncu figure ({e['capture']}), kernel source unchanged since (sha256 {sha[:12]})"
    except Exception as ex:   # noqa: BLE001
        return None, f"no stored traffic figure ({type(ex).__name__})"


class FixedTrainer:
    """Stands in for surrogate.train in the throughput legs: hands out one fixed MLP at the first rendezvous and never looks
    at the batch (accepts_lazy_batch: the engine then does not read the batch length back to the host)."""
    accepts_lazy_batch = True

    def __init__(self, weights):
        self.w = weights

    def __call__(self, batch, signature):
        return self.w if signature == 1 else None


def _fixed_surrogate(P, h1, h2, seed):
    rs = np.random.RandomState(seed)
    u = lambda *s: rs.uniform(-0.05, 0.05, s).astype(np.float32)
    return dict(W1=u(P, h1), b1=u(h1), W2=u(h1, h2), b2=u(h2), W3=u(h2, 1), b3=u(1))


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = C3
    topo = [W["D"], W["H"], W["O"]]

    # ---- CPU baseline first (rank 0, N = 1): the worker pool is forked before this process touches CUDA ----
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        arm = CpuArm(topo)
        cb = arm.measure(30000)
        arm.close()

    import torch
    import torch.distributed as dist
    from sapt_b200 import _lib
    from sapt_b200.engine import EngineConfig, PTEngine
    from sapt_b200.ladder import assign_temperatures
    from sapt_b200.problem import Problem

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    pipes = _lib.measure_pipes(local)
    pipes = _lib.measure_pipes(local)          # second pass: clocks have ramped
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"

    xtr, ytr = synth(W["D"], W["O"], W["n_train"], W["seed"])
    xte, yte = synth(W["D"], W["O"], W["n_test"], W["seed"] + 1)
    path = {"auto": _lib.PATH_AUTO, "f32": _lib.PATH_F32, "tc": _lib.PATH_TC}[args.path]
    prob = Problem.from_f32(topo, xtr, ytr, xte, yte, device=local, path=path)
    strong = args.scaling == "strong"
    Rl = args.chains or (W["chains_per_gpu"] // world if strong else W["chains_per_gpu"])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    prime = W["swap_interval"] + 1   # untimed priming steps BEFORE the warm-up: first exchange (NCCL communicator, sweep scratch)

    def c3_engine(r_local):
        cfg = EngineConfig(variant="rw", samples=args.steps + args.warmup + prime + 2, swap_interval=W["swap_interval"],
                           surrogate_interval=W["swap_interval"], surrogate_prob=0.0, use_surrogate=False, use_langevin=False,
                           seed=W["seed"], record_history=True, record_pos=False, path=path, rank=rank, world=world)
        e = PTEngine(prob, assign_temperatures(r_local * world, W["maxtemp"]), cfg)
        e.initialize(np.random.RandomState(W["seed"] + 2).randn(prob.P))
        for i in range(prime + args.warmup):
            e.step(i); e.rendezvous(i)
        return e

    eng = c3_engine(Rl)
    Rg = Rl * world
    sampler = ClockSampler(local); sampler.start()
    L.sapt_launch_count_reset()
    t_wall0 = time.perf_counter()
    total_ms, _ = _timed_steps(eng, prime + args.warmup, args.steps, flush, world)
    t_wall = time.perf_counter() - t_wall0
    launches = int(L.sapt_launch_count())
    clocks = sampler.stop()
    evals_per_step = Rg * (W["n_train"] + W["n_test"])
    value = evals_per_step * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel timed alone (CUDA events on its stream) for the roofline ----
    work = _lib.work_model(prob.handle, path)
    tc_path = work["path"] == _lib.PATH_TC
    k_ms = _time_alone(lambda: prob.loglik(eng.prop, 0, out=(eng.L_train, eng.nc_train, eng.sq_train), path=path), flush, max(3, min(10, args.steps)))
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    ach_tf = work["flops_algorithmic"] * Rl * W["n_train"] / (k_ms * 1e-3) / 1e12
    sm_mhz = clocks["sm_mhz"] or pipes["sm_hz"] / 1e6
    kname = "loglik_tc2_kernel" if tc_path else "loglik_f32_kernel"
    traffic, traffic_note = _stored_traffic(kname, Rl)
    roofline = dict(bound="tensor", achieved=ach_tf, peak=peak_tf, unit="TFLOP/s", frac=ach_tf / peak_tf, traffic=traffic, traffic_source=traffic_note,
                    kernel=kname, kernel_ms=k_ms, peak_source=peak_src + " burst",
                    note="algorithmic GEMM FLOPs (2DH+2HO per replica x row) vs dense bf16 peak; the kernel is bound by the SFU / issue "
                         "side of its epilogue, see governing. kernel_ms includes the operand pack and the partial-sum fold (<1 %)",
                    tensor_flops_issued_per_eval=work["mma_flops"],
                    governing=_governing(work, pipes, sm_mhz, n_sm, Rl * W["n_train"], k_ms),
                    hbm=dict(achieved_gbs=(Rl * prob.P * 8 + W["n_train"] * (4 * W["D"] + 4)) / (k_ms * 1e-3) / 1e9,
                             peak_gbs=float(peaks.get("hbm_gbs", 6650.0))))

    # ---- e2e: host theta -> C-ABI (sapt_loglik_host: H2D, kernels, D2H) -> host outputs, copies inside the timed region ----
    def e2e_run(thn, steps):
        prob.loglik_host(thn, 0); prob.loglik_host(thn, 1)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            prob.loglik_host(thn, 0)
            prob.loglik_host(thn, 1)
        return evals_per_step * steps / _allmax(time.perf_counter() - t0, world)

    e2e_steps = max(2, min(args.steps, 5))
    th_pinned = torch.empty((Rl, prob.P), dtype=torch.float32).pin_memory()
    th_pinned.copy_(eng.theta[:, : prob.P].cpu())
    th_pageable = np.array(th_pinned.numpy(), copy=True)
    e2e = dict(value=e2e_run(th_pinned.numpy(), e2e_steps), unit=UNIT, h2d_bytes_per_step=2 * Rl * prob.P * 4,
               d2h_bytes_per_step=2 * Rl * (8 + 4 + 8), api="sapt_loglik_host (train + test), pinned host theta", steps=e2e_steps,
               pageable_value=e2e_run(th_pageable, e2e_steps),
               pageable_note="same call with a plain (pageable) numpy theta, what a caller of the reference owns")

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="strong" if strong else "weak", vs_baseline=None,
                dtype="f32" if not tc_path else "bf16x3->f32", data="synthetic",
                config=dict(workload=W["name"], topology=topo, n_train=W["n_train"], n_test=W["n_test"],
                            chains_per_gpu=Rl, chains_total=Rg, sampler="random-walk PT, in-kernel Philox, no surrogate",
                            swap_interval=W["swap_interval"], l2="256 MB memset between timed steps (untimed)",
                            priming=f"{prime} untimed steps before the {args.warmup} warm-up steps (first exchange: NCCL communicator, sweep scratch)",
                            path="tcgen05" if tc_path else "fp32"),
                samples_per_sec=Rg * args.steps / (total_ms * 1e-3), clocks=clocks, e2e=e2e, gpu_launches=launches,
                roofline=roofline, cpu_baseline=cb, wall_s=t_wall,
                pipes={k: pipes[k] for k in ("sms", "sm_hz", "sm_hz_mufu", "ex2_per_clk_sm", "rcp_per_clk_sm", "ffma_per_clk_sm",
                                             "fmul_per_clk_sm", "ffma2_per_clk_sm", "lop3_per_clk_sm", "mix_per_clk_sm")})
    line["exchanges_in_timed_region"] = sum(1 for k in range(args.steps) if (prime + args.warmup + k) % W["swap_interval"] == 0)
    if world == 1:
        res = eng.results()
        line["accept_rate"] = float(np.mean(res["naccept"]) / max(1, args.steps + args.warmup + prime))

    # ---- N > 1: BASELINE configs[2] as written — 1024 chains IN TOTAL sharded over the GPUs (strong scaling) ----
    if world > 1 and not strong and not args.chains and W["chains_per_gpu"] % world == 0:
        try:
            del eng
            es = c3_engine(W["chains_per_gpu"] // world)
            s_steps = min(args.steps, 10)
            ms_s, _ = _timed_steps(es, prime + args.warmup, s_steps, flush, world)
            line["strong"] = dict(chains_total=W["chains_per_gpu"], chains_per_gpu=W["chains_per_gpu"] // world, steps=s_steps,
                                  ms_per_step=ms_s / s_steps, value=W["chains_per_gpu"] * (W["n_train"] + W["n_test"]) * s_steps / (ms_s * 1e-3),
                                  unit=UNIT, note="same workload, total chain count fixed at 1024; the N=1 line of the default run is its 1-GPU point")
            del es
        except Exception as ex:   # noqa: BLE001
            line["strong"] = dict(error=f"{type(ex).__name__}: {ex}")
    if world > 1:
        try:
            line["sharding_invariant"] = sharding_check(rank, world, local)
        except Exception as ex:   # noqa: BLE001
            line["sharding_invariant"] = dict(ok=False, error=f"{type(ex).__name__}: {ex}")

    # ---- secondary configs (untimed for the headline) ----
    if not args.no_secondary:
        eng = None
        del prob
        torch.cuda.empty_cache()
        sec = {}
        legs = (("C4", secondary_c4), ("C5", secondary_c5)) + ((("C2", secondary_c2),) if world == 1 else ())
        for name, fn in legs:
            try:
                sec[name] = fn(rank, world, local, flush, pipes, n_sm, sm_mhz, peaks)
            except Exception as ex:   # noqa: BLE001
                sec[name] = dict(error=f"{type(ex).__name__}: {ex}")
            torch.cuda.empty_cache()
        line["secondary"] = sec
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def secondary_c2(rank, world, local, flush, pipes, n_sm, sm_mhz, peaks):
    """BASELINE configs[1]: Ionosphere-shape 34-50-2 (245 train / 109 test rows), 10 replicas, Langevin-gradient PT with the
    reference's exact per-row SGD epoch (LG:138-156), surrogate interval enabled, the PyTorch surrogate trainer in the loop —
    through the reference-facing class (ParallelTempering). Latency-bound (microseconds of work per step): steps/s, no roofline."""
    import torch
    from sapt_b200.parallel_tempering import ParallelTempering
    D, H, O, R, S = 34, 50, 2, 10, 300
    rs = np.random.RandomState(7)
    mk = lambda n: np.concatenate([rs.standard_normal((n, D)), rs.randint(0, O, n)[:, None].astype(np.float64)], axis=1)
    tr, te = mk(245), mk(109)
    out = {}
    for label, use_sur, samples in (("warm-up (untimed)", True, 120), ("surrogate_off", False, S), ("surrogate_on", True, S)):
        pt = ParallelTempering(use_sur, True, 0.01, False, tr, te, [D, H, O], R, 4.0, R * samples, 100000, 50, 0.5, "", "", 1, variant="lg", seed=3)
        pt.initialize_chains(0.25)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pt.run_chains()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if not label.startswith("warm"):
            out[label] = dict(wall_s=dt, steps_per_s=(samples - 1) / dt, samples_per_s=R * (samples - 1) / dt)
    return dict(workload="C2 Ionosphere-shape 34-50-2 synthetic (245 + 109 rows), 10 chains, exact-SGD Langevin p=0.5, surrogate interval 50 / prob 0.5, torch trainer",
                chains=R, samples_per_chain=S, api="ParallelTempering.run_chains (wall clock, host loop included)", **out)


def secondary_c4(rank, world, local, flush, pipes, n_sm, sm_mhz, peaks):
    """BASELINE configs[3]: Covtype-shape 54-64-7, 581,012 rows, 256 chains per GPU, Langevin PT with the batch gradient
     at l_prob 0.5: tcgen05 likelihood + tcgen05 batch-Langevin kernel."""
    import torch
    from sapt_b200 import _lib
    from sapt_b200.engine import EngineConfig, PTEngine
    from sapt_b200.ladder import assign_temperatures
    from sapt_b200.problem import Problem
    W = C4
    D, H, O, N, Rl = W["D"], W["H"], W["O"], W["n_train"], W["chains_per_gpu"]
    x, y = synth(D, O, N, 1); xt, yt = synth(D, O, W["n_test"], 2)
    prob = Problem.from_f32([D, H, O], x, y, xt, yt, device=local)
    warm, steps = 3, 6
    cfg = EngineConfig(variant="lg", samples=warm + steps + 2, swap_interval=100000, surrogate_interval=100000, surrogate_prob=0.0,
                       use_surrogate=False, use_langevin=True, learn_rate=0.5, langevin_mode="batch", langevin_batch_scale=1.0 / N,
                       seed=W["seed"], record_pos=False, rank=rank, world=world)
    eng = PTEngine(prob, assign_temperatures(Rl * world, W["maxtemp"]), cfg)
    eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
    for i in range(warm):
        eng.step(i); eng.rendezvous(i)
    ms, evals = _timed_steps(eng, warm, steps, flush, world)
    th = eng.theta
    ms_ll = _time_alone(lambda: prob.loglik(th, 0), flush, 3)
    ms_lg = _time_alone(lambda: prob.langevin(th, 0.5, mode=1, batch_scale=1.0 / N), flush, 3)
    work = _lib.work_model(prob.handle, _lib.PATH_AUTO)
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    lg_flops = (4.0 * D * H + 6.0 * H * O) * Rl * N      # forward 2DH+2HO, backward 2DH+4HO
    return dict(workload=W["name"], topology=[D, H, O], n_train=N, n_test=W["n_test"], chains_per_gpu=Rl, chains_total=Rl * world, scaling="weak",
                steps=steps, warmup=warm, ms_per_step=ms / steps, evals_per_s=evals * world / (ms * 1e-3), samples_per_s=Rl * world * steps / (ms * 1e-3),
                kernels_alone_ms=dict(loglik_tc2_kernel_train_all_chains=ms_ll, langevin_tc_kernel_all_chains=ms_lg),
                dominant_kernel="langevin_tc_kernel" if ms_lg * 0.5 > ms_ll else "loglik_tc2_kernel",
                note="a step = likelihood of every chain (train + test) + Langevin gradient and a second likelihood on the ~half of the chains whose coin says so",
                roofline_langevin=dict(bound="tensor", achieved=lg_flops / (ms_lg * 1e-3) / 1e12, peak=peak_tf, unit="TFLOP/s",
                                       frac=lg_flops / (ms_lg * 1e-3) / 1e12 / peak_tf,
                                       note="algorithmic FLOPs 4DH+6HO per replica x row (forward + both gradient contractions); the bf16-split products the kernel issues are not counted"),
                governing_loglik=_governing(work, pipes, sm_mhz, n_sm, Rl * N, ms_ll))


def secondary_c5(rank, world, local, flush, pipes, n_sm, sm_mhz, peaks):
    """BASELINE configs[4]: Chess-shape 6-25-18, 16,833 + 11,223 rows, 8192 chains per GPU (the weak-scaling point of the
    replica sweep), surrogate_prob 0.5 with a fixed random surrogate MLP (topology 3: P-200-50-1), exchange + surrogate
    rendezvous every 4 steps."""
    import torch
    from sapt_b200 import _lib
    from sapt_b200.engine import EngineConfig, PTEngine
    from sapt_b200.ladder import assign_temperatures
    from sapt_b200.problem import Problem
    W = C5
    D, H, O, N, Rl = W["D"], W["H"], W["O"], W["n_train"], W["chains_per_gpu"]
    x, y = synth(D, O, N, 5); xt, yt = synth(D, O, W["n_test"], 6)
    prob = Problem.from_f32([D, H, O], x, y, xt, yt, device=local)
    w = _fixed_surrogate(prob.P, 200, 50, 7)
    warm, steps = 12, 16
    cfg = EngineConfig(variant="rw", samples=warm + steps + 2, swap_interval=W["swap_interval"], surrogate_interval=W["swap_interval"],
                       surrogate_prob=0.5, use_surrogate=True, surrogate_topology=3, seed=W["seed"], record_pos=False, train_cap=W["swap_interval"],
                       rank=rank, world=world)
    eng = PTEngine(prob, assign_temperatures(Rl * world, W["maxtemp"]), cfg, surrogate_trainer=FixedTrainer(w))
    eng.keep_batches = False
    eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
    for i in range(warm):
        eng.step(i); eng.rendezvous(i)
    ms, evals = _timed_steps(eng, warm, steps, flush, world)
    # the pieces alone: likelihood of every chain on the train set; surrogate predict of every chain
    ms_ll = _time_alone(lambda: prob.loglik(eng.prop, 0), flush, 3)
    sw = eng.surrogate.weights
    rows = torch.arange(Rl, dtype=torch.int32, device="cuda"); n_rows = torch.tensor([Rl], dtype=torch.int32, device="cuda")
    import ctypes as C
    vp = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ms_sp = _time_alone(lambda: _lib.check(_lib.lib().sapt_surrogate_predict(vp(eng.prop), eng.ld, Rl, prob.P, vp(rows), vp(n_rows), vp(sw["W1"]), vp(sw["b1"]), 200,
                                                                              vp(sw["W2"]), vp(sw["b2"]), 50, vp(sw["W3"]), vp(sw["b3"]), vp(eng.sur_s), st)), flush, 3)
    work = _lib.work_model(prob.handle, _lib.PATH_AUTO)
    sp_flops = 2.0 * (prob.P * 200 + 200 * 50 + 50) * Rl
    return dict(workload=W["name"], topology=[D, H, O], n_train=N, n_test=W["n_test"], chains_per_gpu=Rl, chains_total=Rl * world, scaling="weak",
                steps=steps, warmup=warm, ms_per_step=ms / steps, evals_per_s=evals * world / (ms * 1e-3), samples_per_s=Rl * world * steps / (ms * 1e-3),
                rendezvous_in_timed_region=sum(1 for k in range(steps) if (warm + k) % W["swap_interval"] == 0),
                kernels_alone_ms=dict(loglik_train_all_chains=ms_ll, surrogate_predict_all_chains=ms_sp),
                dominant_kernel="loglik_tc2_kernel" if work["path"] == _lib.PATH_TC else "loglik_f32_kernel",
                governing_loglik=_governing(work, pipes, sm_mhz, n_sm, Rl * N, ms_ll),
                surrogate_predict=dict(tflops=sp_flops / (ms_sp * 1e-3) / 1e12, share_of_step=0.5 * ms_sp / (ms / steps),
                                       note="2(P*200 + 200*50 + 50) FLOPs per chain; about half the chains take the surrogate branch per step"))


def sharding_check(rank, world, local):
    """30 steps of a small problem (random-walk PT, exchanges every 3 steps, global surrogate with fixed weights from step 7)
    on all N ranks, and the same seed on rank 0 alone: every accept / exchange decision and the final binding must agree."""
    import torch
    import torch.distributed as dist
    from sapt_b200.engine import EngineConfig, PTEngine
    from sapt_b200.ladder import assign_temperatures
    from sapt_b200.problem import Problem
    D, H, O, n = 16, 30, 10, 4096
    x, y = synth(D, O, n, 77); xt, yt = synth(D, O, 1024, 78)
    prob = Problem.from_f32([D, H, O], x, y, xt, yt, device=local)
    Rg, S = 8 * world, 31
    w = _fixed_surrogate(prob.P, 64, 16, 9)

    def run(r, wd):
        cfg = EngineConfig(variant="rw", samples=S, swap_interval=3, surrogate_interval=3, surrogate_prob=0.4, use_surrogate=True,
                           surrogate_topology=1, seed=21, record_pos=False, rank=r, world=wd)
        e = PTEngine(prob, assign_temperatures(Rg, 4.0), cfg, surrogate_trainer=lambda batch, sig: (w if sig == 1 else None))
        e.initialize(np.random.RandomState(5).randn(prob.P) * 0.3)
        for i in range(S - 1):
            e.step(i); e.rendezvous(i)
        torch.cuda.synchronize()
        return e.results()

    multi = run(rank, world)
    dist.barrier()
    out = None
    if rank == 0:
        solo = run(0, 1)
        same = all(np.array_equal(multi[k], solo[k]) for k in ("accepted", "is_surrogate_step", "naccept", "row_of_slot")) \
            and multi["num_swap"] == solo["num_swap"] and all(np.array_equal(a, b) for a, b in zip(multi["swap_decisions"], solo["swap_decisions"]))
        lh = np.nan_to_num(multi["lhood_list"], posinf=0.0, neginf=0.0), np.nan_to_num(solo["lhood_list"], posinf=0.0, neginf=0.0)
        vals = bool(np.allclose(lh[0], lh[1], rtol=1e-12, atol=0))
        out = dict(ok=bool(same and vals), decisions_identical=bool(same), likelihood_histories_identical=vals, chains=Rg, steps=S - 1,
                   exchanges=len(solo["swap_decisions"]), swaps=int(solo["num_swap"]), surrogate_steps=int(solo["is_surrogate_step"].sum()),
                   accepted=int(solo["accepted"].sum()))
    dist.barrier()
    return out


_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner on the
    first communicator), so fd 1 is pointed at stderr for the whole run and the result goes to the saved fd."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default="auto", choices=["auto", "f32", "tc"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 1024 chains per GPU (default); strong: 1024 chains in total, sharded over the GPUs")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the C3 value 1024)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C4 / C5 legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
