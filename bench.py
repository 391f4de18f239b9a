#!/usr/bin/env python3
"""Benchmark of the PAULE planning hot path (BASELINE.json metric: inner planning steps x words / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--math fp32|bf16]

A "step" is one inner planning step (forward of both LSTM models, 5-term loss, BPTT to the cps, Adam + clamp)
for one batch of words.  Workload at any N: BASELINE.json configs[1] per GPU -- 64 words, 0.5 s utterances
(T = 200 cp frames, 100 mel frames), objective acoustic_semvec, random-init H=720 models (torch.manual_seed(0)),
synthetic inputs; words are sharded over ranks with no data-path collective (weak scaling).

Prints ONE JSON line (rank 0).  `value` = words x steps / s with everything resident in HBM (CUDA events, max over
ranks); `e2e` = the same metric through the public API with the cp trajectories living in pinned HOST memory: every
step copies the cps host->device, runs the step, and reads the step's loss terms and updated cps back.
`--impl reference` times the reference's CPU arithmetic (oracle port: torch.nn.LSTM + autograd + torch.optim.Adam,
bit-identical to the reference's plan_resynth loop) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

B_PER_GPU, T_FRAMES, HIDDEN = 64, 200, 720
METRIC, UNIT = "inner planning steps x words per second", "steps*words/s"


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
def cpu_reference_run(words, steps, warmup, threads=None):
    """The reference's CPU arithmetic for the path (oracle port), timed with time.perf_counter."""
    import torch
    from oracle import paule_oracle as O
    if threads:
        torch.set_num_threads(threads)
    pred, emb, _ = O.build_reference_models(0, HIDDEN, torch.float32, with_inverse=False)
    cp0, tmel = O.synthetic_inputs(words, T_FRAMES, seed=5)
    lens = tuple(torch.tensor(T_FRAMES // 2) for _ in range(words))
    with torch.no_grad():
        tsv = emb(tmel, lens)
    x = cp0.clone().requires_grad_()
    opt = torch.optim.Adam([x], lr=0.01)

    def one():
        opt.zero_grad()
        mel = pred(x)
        sv = emb(mel, lens)
        total, _ = O.per_word_losses(mel, tmel, sv, tsv, x)
        total.sum().backward()
        opt.step()
        with torch.no_grad():
            x.data = x.data.clamp(-O.CLAMP, O.CLAMP)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return words * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    words = 16
    steps, warmup = max(1, min(args.steps, 4)), max(1, min(args.warmup, 1))
    val, s_per_step, threads = cpu_reference_run(words, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 0.5 s utterances (T=200), acoustic_semvec, H=720 random-init; "
                                   f"bounded sample of {words} of the 64 words per step on the host CPU"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{words} words x {steps} inner steps, T=200, torch CPU fp32 "
                                       "(nn.LSTM/oneDNN + autograd + optim.Adam), batched with per-word losses"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import paule_b200 as P
    from paule_b200 import _lib, ops
    # oracle/ is touched by the cpu_baseline leg only (cpu_reference_run); the GPU arm generates its own inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()
    tc_ok = _lib.load().paule_tc_packed_lstm_bytes(HIDDEN, 30) > 0
    math_name = args.math or ("bf16" if tc_ok else "fp32")
    math = {"fp32": 0, "bf16": 1, "bf16x3": 2}[math_name]

    B, T = B_PER_GPU, T_FRAMES
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=HIDDEN).to(dev)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=HIDDEN).to(dev)
    cp0, tmel = synthetic_inputs(B, T, seed=5 + rank)
    K, W = args.steps, args.warmup
    planner = P.BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=W + K + 8, math=math)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    planner.step(W)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    planner.step(K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    loss_curve = planner.losses()["total"].mean(1).cpu().tolist()

    # ---- end to end: cps live in pinned host memory, copied in and out every step
    host_cp = torch.empty((B, T, 30), dtype=torch.float32).pin_memory()
    host_cp.copy_(planner.planned_cp().cpu())
    host_terms = torch.empty((B, 6), dtype=torch.float32).pin_memory()
    Ke = max(3, min(K, 20))
    planner2 = P.BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=Ke + 4, math=math)
    staging = torch.empty((B, T, 30), device=dev)

    def e2e_step(i):
        staging.copy_(host_cp, non_blocking=True)             # H2D of this step's input (the cps)
        planner2.set_cp(staging)
        planner2.step(1)
        host_cp.copy_(planner2.planned_cp(), non_blocking=True)          # D2H of the updated cps
        host_terms.copy_(planner2.loss_log[i], non_blocking=True)        # D2H of the step's loss terms
        torch.cuda.synchronize()

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(1, Ke + 1):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- the only cross-GPU traffic of a planning job: one final gather of the planned cps + loss log (NCCL)
    gather_ms = None
    if world > 1:
        from paule_b200 import distributed as D
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.gather_words(planner.planned_cp(), world * B)          # warm-up (NCCL communicator setup)
        barrier()
        g0.record()
        full_cp = D.gather_words(planner.planned_cp(), world * B)
        full_loss = D.gather_words(planner.losses()["total"].transpose(0, 1).contiguous(), world * B)
        g1.record()
        barrier()
        assert full_cp.shape == (world * B, T, 30) and full_loss.shape[0] == world * B
        gather_ms = g0.elapsed_time(g1)

    # ---- the dominant kernel alone (live CUDA-event timing of the recurrent step kernels)
    roof = dominant_kernel_roofline(planner, math, dev)

    # ---- the other half of BASELINE.json's metric: ms per inner step at batch 1 (rank 0), and the throughput regime
    # (one rank's shard of configs[3]: 256 words, 1 s utterances) as supplementary figures
    extra = {}
    if rank == 0:
        def timed(Bx, Tx, steps):
            cpx, tmx = synthetic_inputs(Bx, Tx, seed=77)
            pl = P.BatchPlanner(pred, emb, cpx.to(dev), tmx.to(dev), None, max_log_steps=steps + 4, math=math)
            pl.step(3)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pl.step(steps)
            b.record()
            torch.cuda.synchronize()
            pl.close()
            return a.elapsed_time(b) / steps
        ms1 = timed(1, T, 10)
        extra["batch1"] = {"ms_per_inner_step": ms1, "steps_words_per_s": 1e3 / ms1, "T": T}
        if not args.no_cpu_baseline:
            msL = timed(256, 400, 3)
            extra["configs[3]_shard"] = {"words": 256, "T": 400, "ms_per_inner_step": msL, "steps_words_per_s": 256e3 / msL,
                                         "tflops_algorithmic": flops_per_word_step(400) * 256e3 / msL / 1e12}

    t_ms = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = t_ms.tolist()
    if rank == 0:
        value = world * B * K / (ms_max * 1e-3)
        cpu_val, _, cores = (None, None, None)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_val, cpu_s, cores = cpu_reference_run(16, 3, 1)
            one_val, one_s, _ = cpu_reference_run(1, 3, 1)      # how the reference is used: one word per call
            cpu = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "16 of the 64 words x 3 inner steps (1 warm-up), T=200, torch CPU fp32 oracle port "
                             "(nn.LSTM/oneDNN + autograd + optim.Adam), batched with per-word losses (the stronger baseline)",
                   "batch1": {"value": one_val, "ms_per_inner_step": one_s * 1e3,
                              "sample": "1 word x 3 inner steps: the reference plans one word per call (paule/paule.py:539)"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "bf16 operands / f32 accumulate+state", 2: "bf16x3 / f32"}[math], "data": "synthetic",
            "config": {"workload": "configs[1]: batch 64 words per GPU, 0.5 s utterances (T=200 cp frames, 100 mel frames), "
                                   "mel + semvec + velocity/jerk/local-linear loss, ForwardModel(1x720) + EmbeddingModel(2x720), "
                                   "random-init (seed 0), iid-uniform cps",
                       "words_per_gpu": B, "T": T, "hidden": HIDDEN, "math": math_name, "cuda_graph": True,
                       "l2": "inputs larger than L2: the per-step activation stash is %.0f MB (> 126 MB L2)"
                             % (planner.workspace.numel() / 1e6),
                       "parallelism": f"words sharded over {world} GPU(s), no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": world * B * Ke / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": B * T * 30 * 4, "d2h_bytes_per_step": B * T * 30 * 4 + B * 6 * 4,
                    "steps": Ke, "how": "cps in pinned host memory: H2D cps -> one inner step -> D2H cps + loss terms, every step"},
            "gpu_launches": launches_per_step(T, math) * K,
            "final_gather_ms": gather_ms,
            "roofline": roof,
            "cpu_baseline": cpu,
            "tflops_algorithmic": flops_per_word_step(T) * world * B * K / (ms_max * 1e-3) / 1e12,
            "loss_first_last": [loss_curve[0], loss_curve[-1]],
            "other": extra,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def launches_per_step(T, math):
    """Kernels of libpaule_b200.so launched by one paule_plan_step."""
    Tm = T // 2
    gemms = 1 + 1 + 1 + 1 + 1 + 1 + 1 + 1 + 1 + 1      # 3 input projections, post_linear, head, head^T, 3 dX, post_linear^T
    if math == 0:
        rec = 2 * (T + Tm + Tm)                       # one launch per time step, forward + backward
    else:
        rec = 2 * 3                                   # one persistent launch per layer and direction (batches <= 80 words)
    return 1 + gemms + rec + 2 + 1                    # tick, GEMMs, recurrences, loss (2), Adam  (= 20 for the tcgen05 path)


def dominant_kernel_roofline(planner, math, dev):
    """Time the recurrent kernels of one layer alone with CUDA events (same shapes as the step) and report the
    dominant one against the tensor roofline (it is a [B,720]x[720,2880] GEMM per time step)."""
    import torch
    from paule_b200 import _lib, ops
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

    out = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / reps          # ms per T-step sequence
    name = "bwd" if out["bwd"] >= out["fwd"] else "fwd"
    flops_seq = 2.0 * B * 4 * H * H * T                 # T cell steps of a [B,H]x[H,4H] GEMM
    achieved = flops_seq / (out[name] * 1e-3) / 1e12
    launches = T if math == 0 else 1
    traffic = None
    tpath = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if math != 0 and os.path.exists(tpath):
        tj = json.load(open(tpath)).get(name)
        if tj:
            traffic = tj["dram_bytes_per_word_step"] * B * T   # per launch, like `achieved`
    kname = {"fwd": "tc_lstm_fwd2_kernel<1>", "bwd": "tc_lstm_bwd2_kernel<1>"}[name]
    return {"bound": "tensor", "kernel": ("lstm_step_%s_f32" % name) if math == 0 else kname,
            "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
            "traffic": traffic, "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long sequence)",
            "us_per_cell_step": {k: v * 1e3 / T for k, v in out.items()},
            "launches_timed": launches, "flops_per_launch": flops_seq / launches}


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
    ap.add_argument("--math", default=None, choices=[None, "fp32", "bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
