#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native parallel-tempering hot path.

Workload (BASELINE.json configs[2], "C3"): PenDigit-shape 16-30-10 classifier on synthetic data,
1,000,000 train rows (+65,536 test rows), 1024 chains PER GPU (weak scaling), random-walk
parallel tempering, no surrogate. A "step" is one full MCMC step of every chain: in-kernel
Philox proposal + prior, batched likelihood on train and test, Metropolis-Hastings accept with
history bookkeeping (and the exchange sweep when the step index hits swap_interval).

metric = replica x row likelihood evaluations per second (whole job).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--path auto|f32|tc]

One JSON line on stdout (rank 0). See DESIGN.md "Measurement" for every field.
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
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="C3 PenDigit-shape 16-30-10 synthetic", D=16, H=30, O=10, n_train=1_000_000, n_test=65_536,
                chains_per_gpu=1024, maxtemp=4.0, swap_interval=5, seed=1234)
METRIC = "replica_x_row_likelihood_evals_per_sec"
UNIT = "evals/s"


def synth(D, O, n, seed):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((n, D)).astype(np.float32)
    y = rs.randint(0, O, size=n).astype(np.int32)
    return x, y


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's per-row numpy loop (oracle "port"), one replica per host core
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    topo, n_rows, seed = args
    from oracle import sapt_oracle as so
    D, H, O = topo
    x, y = synth(D, O, n_rows, seed)
    data = np.concatenate([x.astype(np.float64), y[:, None].astype(np.float64)], axis=1)
    w = np.random.RandomState(seed + 1).randn(so.num_param(topo))
    t0 = time.perf_counter()
    so.likelihood_func(data, w, topo, 1.0, rowloop=True)
    return time.perf_counter() - t0


def cpu_baseline(topo, sample_rows=20000, repeats=1):
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        for _ in range(repeats):
            pool.map(_cpu_worker, [(topo, sample_rows, 100 + c) for c in range(cores)])
    wall = time.perf_counter() - t0
    value = cores * sample_rows * repeats / wall
    return dict(value=value, unit=UNIT, cores=cores, kind="port",
                sample=f"{cores} replicas x {sample_rows} synthetic rows x {repeats} pass(es) of the reference's per-row "
                       f"numpy loop (oracle.likelihood_func(rowloop=True)), one process per core, linear in rows"), wall


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    /root/reference (Python, needs keras/matplotlib stubs) cannot travel to the GPU box, so the
    arm runs the oracle's literal per-row port of Network.evaluate_proposal + likelihood_func."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    topo = [WORKLOAD["D"], WORKLOAD["H"], WORKLOAD["O"]]
    sample_rows = 10000
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(topo, sample_rows=2000)
    times, cores = [], os.cpu_count() or 1
    for _ in range(args.steps):
        cb, wall = cpu_baseline(topo, sample_rows=sample_rows)
        times.append(wall)
    total = cores * sample_rows * args.steps
    value = total / sum(times)
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * sum(times) / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD["name"], rows_per_step=sample_rows, replicas_per_step=cores,
                            note="bounded sample of the same workload; cost is linear in rows and replicas"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{cores} processes x {sample_rows} rows per step, reference per-row numpy loop"),
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sapt_b200 import _lib
    from sapt_b200.engine import EngineConfig, PTEngine
    from sapt_b200.ladder import assign_temperatures
    from sapt_b200.problem import Problem

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = WORKLOAD
    topo = [W["D"], W["H"], W["O"]]
    xtr, ytr = synth(W["D"], W["O"], W["n_train"], W["seed"])
    xte, yte = synth(W["D"], W["O"], W["n_test"], W["seed"] + 1)
    path = {"auto": _lib.PATH_AUTO, "f32": _lib.PATH_F32, "tc": _lib.PATH_TC}[args.path]
    prob = Problem.from_f32(topo, xtr, ytr, xte, yte, device=local, path=path)
    Rl = args.chains or W["chains_per_gpu"]
    Rg = Rl * world
    temps = assign_temperatures(Rg, W["maxtemp"])
    cfg = EngineConfig(variant="rw", samples=args.steps + max(args.warmup, W["swap_interval"] + 1) + 2, swap_interval=W["swap_interval"],
                       surrogate_interval=W["swap_interval"], surrogate_prob=0.0, use_surrogate=False, use_langevin=False,
                       seed=W["seed"], record_history=True, record_pos=False, path=path, rank=rank, world=world)
    eng = PTEngine(prob, temps, cfg)
    w0 = np.random.RandomState(W["seed"] + 2).randn(prob.P)
    eng.initialize(w0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    # warm-up covers at least one exchange (first-use costs of the NCCL all-reduce and of the sweep kernel's scratch)
    n_warm = max(args.warmup, W["swap_interval"] + 1)
    for i in range(n_warm):
        eng.step(i); eng.rendezvous(i)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    sampler = ClockSampler(local); sampler.start()
    L = _lib.lib(); L.sapt_launch_count_reset()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        i = n_warm + k
        flush.zero_()                     # L2 flush between timed steps (not timed)
        evs[k][0].record()
        eng.step(i); eng.rendezvous(i)
        evs[k][1].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = int(L.sapt_launch_count())
    if world > 1: dist.barrier()
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    evals_per_step = Rg * (W["n_train"] + W["n_test"])
    value = evals_per_step * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel timed alone (CUDA events on its stream) for the roofline ----
    torch.cuda.synchronize()
    reps = max(3, min(10, args.steps))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(reps):
        flush.zero_()
        e0.record(); prob.loglik(eng.prop, 0, out=(eng.L_train, eng.nc_train, eng.sq_train), path=path); e1.record()
        torch.cuda.synchronize(); kms.append(e0.elapsed_time(e1))
    k_ms = float(np.mean(kms))
    flops_per_eval = 2 * W["D"] * W["H"] + 2 * W["H"] * W["O"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0)); peak_src = "measured (MEASURED_PEAKS.json burst)" if peaks else "fallback"
    ach_tf = flops_per_eval * Rl * W["n_train"] / (k_ms * 1e-3) / 1e12
    sm_mhz = clocks["sm_mhz"] or 1965.0
    mufu_per_eval = 2 * (W["H"] + W["O"]) + W["O"] + 1
    ffma_per_eval = W["D"] * W["H"] + W["H"] * W["O"]
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    tc_path = bool(prob.tc_supported and path != _lib.PATH_F32)
    # SM-pipe ceiling at the clock observed during the run: MUFU 16/clk/SM; FP32 FMA 128/clk/SM
    # (2(H+O)+O+1 MUFU per eval with an ex2+rcp sigmoid; both GEMMs leave the FMA pipe on the
    # tcgen05 path, so only the SFU term remains there)
    ffma_eff = 0 if tc_path else ffma_per_eval
    clk_per_eval = max(mufu_per_eval / 16.0, ffma_eff / 128.0)
    gov_peak = n_sm * sm_mhz * 1e6 / clk_per_eval
    gov_ach = Rl * W["n_train"] / (k_ms * 1e-3)
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        key = "loglik_tc2_kernel" if tc_path else "loglik_f32_kernel"
        if tj.get(key) and tj[key].get("chains") == Rl:
            traffic = tj[key]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = dict(bound="tensor", achieved=ach_tf, peak=peak_tf, unit="TFLOP/s", frac=ach_tf / peak_tf, traffic=traffic,
                    kernel="loglik_tc2_kernel" if tc_path else "loglik_f32_kernel", kernel_ms=k_ms, peak_source=peak_src,
                    note="GEMM-only FLOPs (2DH+2HO per replica x row) vs dense bf16 peak; the kernel is bound by the "
                         "SFU/FP32 epilogue, see governing",
                    governing=dict(bound="sm_pipes(mufu|fp32)", achieved=gov_ach, peak=gov_peak, unit=UNIT, frac=gov_ach / gov_peak,
                                   model=f"{n_sm} SMs x {sm_mhz:.0f} MHz / max({mufu_per_eval} MUFU/16, {ffma_eff} FFMA/128) clk per eval"),
                    hbm=dict(achieved_gbs=(Rl * prob.P * 8 + W["n_train"] * (4 * W["D"] + 4)) / (k_ms * 1e-3) / 1e9,
                             peak_gbs=float(peaks.get("hbm_gbs", 6650.0))))

    # ---- e2e: host theta (pinned) -> C-ABI -> host outputs, copies inside the timed region ----
    th_host = torch.empty((Rl, prob.P), dtype=torch.float32).pin_memory()
    th_host.copy_(eng.theta[:, : prob.P].cpu())
    thn = th_host.numpy()
    e2e_steps = max(2, min(args.steps, 5))
    prob.loglik_host(thn, 0); prob.loglik_host(thn, 1)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prob.loglik_host(thn, 0)
        prob.loglik_host(thn, 1)
    te = time.perf_counter() - t0
    tt = torch.tensor([te], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_val = evals_per_step * e2e_steps / float(tt.item())
    e2e = dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=2 * Rl * prob.P * 4, d2h_bytes_per_step=2 * Rl * (8 + 4 + 8),
               api="sapt_loglik_host (train + test) with pinned host theta", steps=e2e_steps)

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb, _ = cpu_baseline(topo, sample_rows=20000)

    if rank == 0:
        res = eng.results() if world == 1 else None
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=n_warm,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32" if not tc_path else "bf16x3->f32", data="synthetic",
                    config=dict(workload=W["name"], topology=topo, n_train=W["n_train"], n_test=W["n_test"],
                                chains_per_gpu=Rl, chains_total=Rg, sampler="random-walk PT, in-kernel Philox, no surrogate",
                                swap_interval=W["swap_interval"], l2="256 MB memset between timed steps (untimed)",
                                path="tcgen05" if tc_path else "fp32"),
                    samples_per_sec=Rg * args.steps / (total_ms * 1e-3), clocks=clocks, e2e=e2e, gpu_launches=launches,
                    roofline=roofline, cpu_baseline=cb, wall_s=t_wall)
        if res is not None:
            line["accept_rate"] = float(np.mean(res["naccept"]) / max(1, args.steps + n_warm))
            line["exchanges_in_timed_region"] = sum(1 for k in range(args.steps) if (n_warm + k) % W["swap_interval"] == 0)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (default: the C3 value 1024)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
