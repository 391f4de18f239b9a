#!/usr/bin/env python3
"""Benchmark of the PAULE planning hot path (BASELINE.json metric: inner planning steps x words / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math fp32|bf16]

A "step" is one inner planning step (forward of both LSTM models, 5-term loss, BPTT to the cps, Adam + clamp) for one
batch of words.  Workloads (BASELINE.json `configs`; random-init H=720 models from torch.manual_seed(0),
synthetic iid-uniform cps / mel targets, objective acoustic_semvec):

  N = 1   configs[1]: 64 words, 0.5 s utterances (T = 200 cp frames, 100 mel frames)           -- the headline
  N > 1   configs[3]: 2048 words, 1 s utterances (T = 400), sharded over the N GPUs (STRONG scaling: 1024 / 512 / 256 words
          per GPU), no data-path collective, one final NCCL all_gather of the planned cps and the loss log

Prints ONE JSON line (rank 0).
  value   words x steps / s with all state resident in HBM: CUDA events around the timed steps (barrier + synchronize on both
          sides, max over ranks).  The K steps are repeated `timed_reps` times back to back so that the timed region lasts
          >= 1 s (clock sampling needs it); ms_per_step is the mean over all of them.
  e2e     the same metric through the reference-facing API: wall clock of `Paule.plan_resynth(target_acoustic=<host numpy
          [B,Tm,60]>, initial_cp=<host numpy [B,T,30]>, n_outer=1, n_inner=K)` returning a `PlanningResults` of host arrays
          (at N > 1 through `distributed.plan_resynth_sharded`, i.e. including the final NCCL gather).
  roofline / cpu_baseline / other: see DESIGN.md section 6.
`--impl reference` times the reference's CPU arithmetic for the same workload (oracle port: torch.nn.LSTM + autograd +
torch.optim.Adam, bit-identical to the reference's plan_resynth loop) on the host cores, thread count swept.
"""
import argparse
import json
import math as pymath
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

HIDDEN = 720
METRIC, UNIT = "inner planning steps x words per second", "steps*words/s"
CFG1 = dict(name="configs[1]", words=64, T=200,
            text="configs[1]: batch 64 words on 1 B200, 0.5 s utterances (T=200 cp frames, 100 mel frames), mel + semvec + "
                 "velocity/jerk/local-linear loss, ForwardModel(1x720) + EmbeddingModel(2x720), random-init (seed 0), iid-uniform cps")
CFG3 = dict(name="configs[3]", words=2048, T=400,
            text="configs[3]: batch 2048 words sharded across the GPUs, 1 s utterances (T=400 cp frames, 200 mel frames), same "
                 "models / loss, final NCCL gather of planned cps + loss log")


def workload(n_gpus):
    return CFG1 if n_gpus <= 1 else CFG3


def synthetic_inputs(B, T, seed=5):
    """SURVEY 8(d) synthetic workload (the same generator the oracle uses for its checks): cp0 ~ U(-0.5, 0.5) iid [B,T,30]
    (the well-conditioned regime), target_mel ~ U(0, 1) [B,T//2,60], both drawn in float64 from one seeded generator."""
    import torch
    g = torch.Generator().manual_seed(seed)
    cp0 = torch.rand(B, T, 30, generator=g, dtype=torch.float64) - 0.5
    tmel = torch.rand(B, T // 2, 60, generator=g, dtype=torch.float64)
    return cp0.float(), tmel.float()


def flops_per_word_step(T):
    """forward + input-gradient backward, weight-gradient FLOPs excluded."""
    return 21.6e6 * T + 0.864e6


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------------------------
# the reference's CPU arithmetic (oracle port), timed on the host cores
# --------------------------------------------------------------------------------------------------------------
class CpuLoop:
    """The reference's inner loop (paule/paule.py:910-1211 without logging / VocalTractLab) for `words` words batched with
    per-word losses: the oracle port, i.e. the same torch CPU operators the reference runs (nn.LSTM / oneDNN, autograd,
    torch.optim.Adam)."""

    def __init__(self, words, T):
        import torch
        from oracle import paule_oracle as O
        self.O, self.torch, self.words = O, torch, words
        self.pred, self.emb, _ = O.build_reference_models(0, HIDDEN, torch.float32, with_inverse=False)
        cp0, self.tmel = O.synthetic_inputs(words, T, seed=5)
        self.lens = tuple(torch.tensor(T // 2) for _ in range(words))
        with torch.no_grad():
            self.tsv = self.emb(self.tmel, self.lens)
        self.x = cp0.clone().requires_grad_()
        self.opt = torch.optim.Adam([self.x], lr=0.01)

    def step(self):
        O, torch = self.O, self.torch
        self.opt.zero_grad()
        mel = self.pred(self.x)
        sv = self.emb(mel, self.lens)
        total, _ = O.per_word_losses(mel, self.tmel, sv, self.tsv, self.x)
        total.sum().backward()
        self.opt.step()
        with torch.no_grad():
            self.x.data = self.x.data.clamp(-O.CLAMP, O.CLAMP)

    def timed(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return (time.perf_counter() - t0) / steps


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference(words, T, steps, warmup, sweep=True):
    """Best thread count first (one warm-up + one timed step per candidate -- torchrun exports OMP_NUM_THREADS=1, which
    torch.set_num_threads overrides), then `steps` timed steps at that count.  Returns (steps*words/s, s/step, threads, sweep)."""
    import torch
    loop = CpuLoop(words, T)
    avail = host_threads()
    cands = sorted({c for c in (1, 4, 8, 16, 32, 64, avail) if c <= avail}) if sweep else [avail]
    table = {}
    for c in cands:
        torch.set_num_threads(c)
        table[c] = loop.timed(1, 1)
    best = min(table, key=table.get)
    torch.set_num_threads(best)
    s = loop.timed(steps, warmup)
    return words / s, s, best, {str(k): round(words / v, 2) for k, v in table.items()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload(args.gpus)
    words = 64                                         # all of configs[1]; a 64-word sample of configs[3]'s 2048
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    val, s_per_step, threads, table = cpu_reference(words, cfg["T"], steps, warmup)
    sample = (f"{words} words" + ("" if cfg["words"] == words else f" of the {cfg['words']}") + f" x {steps} inner steps "
              f"({warmup} warm-up), T={cfg['T']}, torch CPU fp32 (nn.LSTM/oneDNN + autograd + optim.Adam), batched with per-word "
              f"losses; thread sweep (steps*words/s per thread count, 1 step each): {table}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
            "scaling": "weak" if args.gpus <= 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["text"], "words": cfg["words"], "T": cfg["T"], "hidden": HIDDEN,
                       "sampled_words_per_step": words},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "host_threads_available": host_threads()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------------------------
def rnn_passes(B):
    """launches of one persistent recurrent layer (forward, backward) for B words: a launch holds at most 6 / 5 word groups of
    up to 64 words (csrc/tc_lstm.cuh: kMaxQ, kMaxQBwd)."""
    return -(-B // (6 * 64)), -(-B // (5 * 64))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import paule_b200 as P
    from paule_b200 import _lib, ops
    from paule_b200 import distributed as D
    # oracle/ is touched by the cpu_baseline leg only (cpu_reference); the GPU arm generates its own inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()
    math_name = args.math or ("bf16" if ops.tc_available() else "fp32")
    math = {"fp32": 0, "bf16": 1}[math_name]

    cfg = workload(world)
    Btot, T = cfg["words"], cfg["T"]
    lo, hi = D.shard_bounds(Btot, world, rank)
    B = hi - lo
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=HIDDEN).to(dev)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=HIDDEN).to(dev)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=HIDDEN).to(dev)
    cp_all, tmel_all = synthetic_inputs(Btot, T, seed=5)          # every rank draws the whole job, then takes its shard
    cp0, tmel = cp_all[lo:hi], tmel_all[lo:hi]
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up steps, then `reps` x K timed steps (>= 1 s of timed region)
    probe = P.BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=W + 4, math=math)
    probe.step(W)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); probe.step(2); b.record(); torch.cuda.synchronize()
    est_ms = a.elapsed_time(b) / 2
    probe.close(); del probe
    reps = max(1, int(pymath.ceil(args.min_timed_s * 1e3 / (K * est_ms))))
    if world > 1:
        r_t = torch.tensor([reps], device=dev)
        dist.all_reduce(r_t, op=dist.ReduceOp.MAX)
        reps = int(r_t.item())
    planner = P.BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=W + reps * K + 8, math=math)
    planner.step(W)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    planner.step(reps * K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps                               # ms per K steps
    clocks = sampler.stop() if rank == 0 else None
    planner.check()
    loss_curve = planner.losses()["total"].mean(1).cpu().tolist()
    ws_mb = planner.workspace.numel() / 1e6
    n_launch = planner.launches_per_step()
    roof = dominant_kernel_roofline(planner, math, dev) if rank == 0 else None
    kernels = kernel_rooflines(planner, math, dev) if (rank == 0 and math != 0) else None
    planner.close(); del planner
    torch.cuda.empty_cache()

    # ---- end to end through the reference-facing API: host numpy in, PlanningResults (host numpy) out
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, math=math)
    cp_np, tmel_np = cp_all.numpy(), tmel_all.numpy()
    kw = dict(initialize_from=None, objective="acoustic_semvec", n_outer=1, n_inner=K, continue_learning=False, verbose=False)

    def e2e_call():
        if world > 1:      # shard -> plan_resynth on the local words -> NCCL all_gather of planned cps + loss log
            return D.plan_resynth_sharded(pm, target_acoustic=tmel_np, initial_cp=cp_np, **kw)
        return pm.plan_resynth(target_acoustic=tmel_np, initial_cp=cp_np, **kw)

    res = e2e_call()                                              # warm-up: builds the planner, captures the graph, NCCL setup
    res = e2e_call()
    e2e_reps = max(2, int(pymath.ceil(min(args.min_timed_s, 1.0) * 1e3 / (K * est_ms))))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_reps):
        res = e2e_call()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_reps                 # s per K-step call
    local_res = res.local if world > 1 else res
    h2d = (cp0.numel() + tmel.numel()) * 4
    d2h = sum(np.asarray(x).nbytes for x in (local_res.planned_cp, local_res.initial_cp, local_res.initial_pred_mel,
                                             local_res.target_mel, local_res.pred_mel, local_res.initial_pred_semvec,
                                             local_res.pred_semvec)) + K * B * 6 * 4
    gather = None
    if world > 1:
        # the N-GPU job must return what one GPU would: rank 0 re-plans 16 words of the LAST rank's shard alone
        full_cp, full_loss = res.planned_cp, res.planned_loss_steps
        assert full_cp.shape == (Btot, T, 30) and full_loss.shape == (K, Btot)
        if rank == 0:
            w0 = Btot - 16
            solo = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, math=math).plan_resynth(
                target_acoustic=tmel_np[w0:], initial_cp=cp_np[w0:], **kw)
            d_cp = float(np.abs(solo.planned_cp - full_cp[w0:]).max())
            d_loss = float(np.abs(np.stack(solo.planned_loss_steps) / full_loss[:, w0:] - 1).max())
            assert d_cp < 1e-3 and d_loss < 1e-3, (d_cp, d_loss)
            gather = {"checked_words": [w0, Btot], "owner_rank": world - 1, "max_abs_cp_diff_vs_one_gpu": d_cp,
                      "max_rel_loss_diff_vs_one_gpu": d_loss, "gathered_bytes": int(full_cp.nbytes + full_loss.nbytes)}
    pm.last_planner.close()
    del pm
    torch.cuda.empty_cache()

    # ---- the other half of BASELINE.json's metric (ms per inner step at batch 1) and the other BASELINE configs
    extra = {}

    def timed(Bx, Tx, steps, seed=77, warm=3):
        cpx, tmx = synthetic_inputs(Bx, Tx, seed=seed)
        pl = P.BatchPlanner(pred, emb, cpx.to(dev), tmx.to(dev), None, max_log_steps=steps + warm + 1, math=math)
        pl.step(warm)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pl.step(steps)
        b.record()
        torch.cuda.synchronize()
        pl.check()
        gib = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        pl.close(); del pl
        torch.cuda.empty_cache()
        return a.elapsed_time(b) / steps, gib

    def entry(Bx, Tx, msx, gib=None, n=1):
        d = {"words": Bx * n, "T": Tx, "ms_per_inner_step": msx, "steps_words_per_s": Bx * n * 1e3 / msx,
             "tflops_algorithmic": flops_per_word_step(Tx) * Bx * n * 1e3 / msx / 1e12}
        if gib is not None:
            d["peak_mem_gib"] = round(gib, 1)
        return d

    if not args.quick:
        if world == 1:
            ms1, _ = timed(1, 200, 20)
            extra["batch1"] = {"ms_per_inner_step": ms1, "steps_words_per_s": 1e3 / ms1, "T": 200}
            # configs[2]: InverseModel initialisation + planning, 256 words, 1 s utterances, through Paule.plan_resynth
            _, tm2 = synthetic_inputs(256, 400, seed=41)
            pm2 = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, math=math)
            k2 = dict(target_acoustic=tm2.numpy(), initialize_from="acoustic", objective="acoustic_semvec", n_outer=1,
                      n_inner=10, continue_learning=False, verbose=False)
            pm2.plan_resynth(**k2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pm2.plan_resynth(**k2)
            t2 = time.perf_counter() - t0
            pm2.last_planner.close(); del pm2
            torch.cuda.empty_cache()
            ms2, g2 = timed(256, 400, 5)
            extra["configs[2]"] = dict(entry(256, 400, ms2, g2), plan_resynth_10_steps_incl_inverse_init_s=t2,
                                       e2e_steps_words_per_s=2560 / t2)
            torch.cuda.reset_peak_memory_stats(dev)
            ms3, g3 = timed(2048, 400, 3, warm=2)
            extra["configs[3]_on_one_gpu"] = entry(2048, 400, ms3, g3)
            torch.cuda.reset_peak_memory_stats(dev)
            ms4, g4 = timed(512, 1200, 3, warm=2)
            extra["configs[4]"] = entry(512, 1200, ms4, g4)
        else:
            # configs[4] sharded: 512 words x 3 s utterances over the N GPUs (strong scaling of the long-utterance stress)
            l4, h4 = D.shard_bounds(512, world, rank)
            ms4, g4 = timed(h4 - l4, 1200, 3, seed=77 + rank, warm=2)
            t4 = torch.tensor([ms4], device=dev, dtype=torch.float64)
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
            extra["configs[4]_sharded"] = dict(entry(512, 1200, float(t4.item()), g4), words_per_gpu=h4 - l4)

    t_ms = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t_ms.tolist()
    if rank == 0:
        value = Btot * K / (ms_max * 1e-3)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_val, cpu_s, cores, table = cpu_reference(64, T, 5, 1)
            one_val, one_s, one_cores, _ = cpu_reference(1, T, 5, 1)   # how the reference is used: one word per call
            cpu = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"all 64 words x 5 inner steps (1 warm-up), T={T}, torch CPU fp32 oracle port (nn.LSTM/oneDNN + autograd "
                             f"+ optim.Adam), batched with per-word losses (the stronger baseline); thread sweep {table}",
                   "host_threads_available": host_threads(),
                   "batch1": {"value": one_val, "ms_per_inner_step": one_s * 1e3, "cores": one_cores,
                              "sample": "1 word x 5 inner steps: the reference plans one word per call (paule/paule.py:539)"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None,
            "dtype": {0: "f32", 1: "bf16 operands / f32 accumulate+state"}[math], "data": "synthetic",
            "config": {"workload": cfg["text"], "words": Btot, "words_per_gpu": B, "T": T, "hidden": HIDDEN, "math": math_name,
                       "cuda_graph": True, "timed_reps": reps,
                       "l2": "inputs larger than L2: the per-step activation stash is %.0f MB (> 126 MB L2)" % ws_mb,
                       "parallelism": f"words sharded over {world} GPU(s), no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": Btot * K / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K, "steps": K, "calls_timed": e2e_reps,
                    "ms_per_call": e2e_ms_max,
                    "how": "wall clock of Paule.plan_resynth(target_acoustic=host numpy [B,Tm,60], initial_cp=host numpy [B,T,30], "
                           "n_outer=1, n_inner=K) -> PlanningResults of host arrays"
                           + ("" if world == 1 else " on every rank's shard + final NCCL all_gather of planned cps and loss log "
                                                    "(distributed.plan_resynth_sharded)")},
            "gpu_launches": n_launch * K * reps,
            "final_gather": gather,
            "roofline": roof,
            "kernel_rooflines": kernels,
            "cpu_baseline": cpu,
            "tflops_algorithmic": flops_per_word_step(T) * Btot * K / (ms_max * 1e-3) / 1e12,
            "loss_first_last": [loss_curve[0], loss_curve[-1]],
            "other": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _time_fn(fn, reps=5, flush=None):
    import torch
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def dominant_kernel_roofline(planner, math, dev):
    """Time the recurrent kernels of one layer alone with CUDA events (same shapes as the step) and report the
    dominant one against the tensor roofline (it is a [B,720]x[720,2880] GEMM per time step)."""
    import torch
    from paule_b200 import _lib
    lib = _lib.load()
    B, T, H = planner.B, planner.T, planner.H
    pk = peaks()
    gates = torch.randn((T, B, 4 * H), device=dev) * 0.1
    h = torch.empty((T, B, H), device=dev)
    c = torch.empty((T, B, H), device=dev)
    scratch = torch.empty((B, H), device=dev)
    dh = torch.randn((T // 2, B, H), device=dev) * 1e-3
    L = planner.w_fwd
    st = torch.cuda.current_stream().cuda_stream
    xchg = None
    if math != 0:
        xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)   # status word starts at 0

    def fwd():
        if math == 0:
            _lib.check(lib.paule_lstm_seq_fwd_f32(gates.data_ptr(), L.w_hh.data_ptr(), h.data_ptr(), c.data_ptr(), T, B, H, st))
        else:
            _lib.check(lib.paule_tc_lstm_seq_fwd(gates.data_ptr(), L.packed.data_ptr(), h.data_ptr(), c.data_ptr(),
                                                 xchg.data_ptr(), None, T, B, math, st))

    def bwd():
        if math == 0:
            _lib.check(lib.paule_lstm_seq_bwd_f32(gates.data_ptr(), c.data_ptr(), L.w_hh_t.data_ptr(), dh.data_ptr(), 2, None,
                                                  scratch.data_ptr(), T, B, H, st))
        else:
            _lib.check(lib.paule_tc_lstm_seq_bwd(gates.data_ptr(), c.data_ptr(), L.packed.data_ptr(), dh.data_ptr(), 2, None,
                                                 xchg.data_ptr(), None, T, B, math, st))

    out = {name: _time_fn(fn, reps=3) for name, fn in (("fwd", fwd), ("bwd", bwd))}     # ms per T-step sequence
    name = "bwd" if out["bwd"] >= out["fwd"] else "fwd"
    flops_seq = 2.0 * B * 4 * H * H * T                 # T cell steps of a [B,H]x[H,4H] GEMM
    f_pass, b_pass = rnn_passes(B)
    launches = T if math == 0 else (f_pass if name == "fwd" else b_pass)
    achieved = flops_seq / (out[name] * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if math != 0 and os.path.exists(tpath):
        tj = json.load(open(tpath)).get(name)
        if tj:
            traffic = tj["dram_bytes_per_word_step"] * B * T / launches   # per launch, like `achieved`
    kname = {"fwd": "tc_lstm_fwd2_kernel", "bwd": "tc_lstm_bwd2_kernel"}[name]
    return {"bound": "tensor", "kernel": ("lstm_step_%s_f32" % name) if math == 0 else kname,
            "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
            "traffic": traffic, "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long sequence)",
            "us_per_cell_step": {k: v * 1e3 / T for k, v in out.items()},
            "launches_timed": launches, "flops_per_launch": flops_seq / launches,
            "note": "a chain of T dependent cell steps: latency-bound at this batch size, neither roofline binds (DESIGN.md 4)"}


def kernel_rooflines(planner, math, dev):
    """The two other kernel classes BASELINE.json's north_star asks about, each timed alone (CUDA events, L2 flushed between
    launches): the gate GEMM over all time steps on tcgen05 against the burst tensor peak, and the fused Adam + clamp kernel
    against the measured HBM bandwidth."""
    import torch
    from paule_b200 import _lib
    lib = _lib.load()
    pk = peaks()
    B, T, H, Tm = planner.B, planner.T, planner.H, planner.Tm
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    # gate GEMM of embedder layer 1: [Tm*Bw, 720] x [720, 2880] over the bf16 image sequence layer 0 leaves behind (wide outputs run
    # on CTA pairs: tc_gemm_img2_kernel, tcgen05.mma.cta_group::2).  Timed at the workload's batch and at 256 words per GPU (the
    # per-GPU batch of configs[3] on 8 GPUs): the small instance is 7 waves of tiles behind a fixed launch / ramp cost.
    L1 = planner.w_e1
    for key, Bw in (("gate_gemm", B), ("gate_gemm_256_words", 256)):
        img = torch.zeros(lib.paule_tc_img_seq_bytes(Tm, Bw, 1), dtype=torch.uint8, device=dev)
        cbuf = torch.empty((Tm, Bw, 4 * H), device=dev)

        def gemm():
            _lib.check(lib.paule_tc_gemm_img(img.data_ptr(), L1.packed_ih.data_ptr(), L1.bias.data_ptr(), cbuf.data_ptr(), Tm, Bw,
                                             4 * H, 1, 0, st))
        ms = _time_fn(gemm, flush=flush)
        fl = 2.0 * Tm * Bw * 4 * H * H       # algorithmic: K = 720 (the kernel multiplies the zero padding up to 768 as well)
        tf = fl / (ms * 1e-3) / 1e12
        out[key] = {"kernel": "tc_gemm_img2_kernel", "shape": [Tm * Bw, 4 * H, H], "bound": "tensor", "us": ms * 1e3,
                    "achieved": tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops_sustained"],
                    "peak_kind": "sustained cuBLAS bf16", "frac_of_burst_peak": tf / pk["bf16_tflops"],
                    "timing": "alone, L2 flushed before every launch"}
        del img, cbuf
    # Adam + clamp: 5 streams in (x, g_lstm, g_smooth, m, v), 3 out (x, m, v) of T*B*30 fp32
    n = T * B * 30
    x, g1, g2, m, v = (torch.rand(n, device=dev) for _ in range(5))
    step = torch.ones(1, dtype=torch.int32, device=dev)

    def adam():
        _lib.check(lib.paule_adam_clamp_f32(x.data_ptr(), g1.data_ptr(), g2.data_ptr(), m.data_ptr(), v.data_ptr(),
                                            step.data_ptr(), 0.01, 0.9, 0.999, 1e-8, 1.05, 0, None, 0, T, B, 30, st))
    ms = _time_fn(adam, flush=flush)
    by = 8.0 * n * 4
    out["adam_clamp"] = {"kernel": "adam_clamp_kernel", "bound": "hbm", "us": ms * 1e3, "achieved": by / (ms * 1e-3) / 1e9,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                         "bytes_per_launch": by}
    return out


_REAL_STDOUT = None


def emit(line) -> None:
    """the one JSON line, on the process's original stdout"""
    sys.stdout.flush()
    text = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(text); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--math", default=None, choices=[None, "fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the supplementary configs under `other`")
    ap.add_argument("--min-timed-s", type=float, default=1.0, help="repeat the K timed steps until the region lasts this long")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    # stdout carries exactly ONE JSON line: while the run is in progress file descriptor 1 points at stderr, so that anything
    # a library writes there (e.g. NCCL's "NCCL version ..." banner, which ignores NCCL_DEBUG_FILE) cannot get in front of it
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
