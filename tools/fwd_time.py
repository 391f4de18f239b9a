"""ms of the forward pass alone (paule_plan_forward: forward model -> pred_mel -> embedder -> semvec) and of a full inner step,
per batch size.  PAULE_NO_WAVEFRONT=1 selects the serial schedule.  Usage: python tools/fwd_time.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paule_b200 as P
dev = torch.device("cuda:0"); torch.manual_seed(0)
pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720).to(dev)
emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720).to(dev)
T = int(os.environ.get("T", "200"))
for B in [int(a) for a in sys.argv[1:]] or [1, 16, 32, 64, 128]:
    g = torch.Generator().manual_seed(5)
    cp0 = torch.rand(B, T, 30, generator=g) - 0.5
    tmel = torch.rand(B, T // 2, 60, generator=g)
    pl = P.BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=64, math=1)
    pl.step(3)
    def timed(fn, n):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    f = timed(lambda: P.ops.plan_forward(pl.cp, pl.pred_mel, pl.pred_sv, pl.workspace, pl._key), 10)
    s = timed(lambda: pl.step(1), 20)
    pl.check()
    print(f"B={B:4d} T={T}: forward {f:.3f} ms   inner step {s:.3f} ms   ({'serial' if os.environ.get('PAULE_NO_WAVEFRONT') else 'wavefront'})", flush=True)
    pl.close()
