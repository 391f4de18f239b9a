"""BASELINE configs 3-5 at their full single-GPU sizes: steps x words / s, peak memory, finiteness + monotone loss."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paule_b200 as P
from oracle import paule_oracle as O
dev = torch.device("cuda:0"); torch.manual_seed(0)
pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720).to(dev)
emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720).to(dev)
cfgs = {"cfg3 B=256 T=400": (256, 400), "cfg4 shard B=1024 T=400 (2 GPUs)": (1024, 400), "cfg4 shard B=512 T=400 (4 GPUs)": (512, 400),
        "cfg5 B=512 T=1200": (512, 1200), "cfg5 shard B=64 T=1200 (8 GPUs)": (64, 1200)}
for name in (sys.argv[1:] or list(cfgs)):
    B, T = cfgs[name]
    torch.cuda.reset_peak_memory_stats()
    cp0, tmel = O.synthetic_inputs(B, T, seed=9)
    pl = P.BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=8, math=1)
    pl.step(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pl.step(3); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    L = pl.losses()["total"].mean(1).cpu()
    ok = bool(torch.isfinite(L).all() and (L[1:] < L[:-1]).all())
    print(f"{name:36s} {ms:9.2f} ms/step  {B * 1e3 / ms:9.0f} steps*words/s  {21.6e6 * T * B / ms / 1e9:6.1f} TFLOP/s(alg)  "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:5.1f} GiB  loss {L[0]:.1f}->{L[-1]:.1f} monotone+finite={ok}", flush=True)
    pl.close(); del pl; torch.cuda.empty_cache()
