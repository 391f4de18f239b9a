import sys, time, torch
sys.path.insert(0, "/root/repo")
import paule_b200 as P
from paule_b200 import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720).to(dev)
emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720).to(dev)
x = (torch.rand(64, 200, 30, device=dev) - 0.5)
for math in (0, 1):
    pred.math = emb.math = math
    for rep in range(3):
        xr = x.clone().requires_grad_()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sv = emb(pred(xr), [100] * 64)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        sv.square().sum().backward()
        torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"math={math}: pred+embedder forward {1e3*(t1-t0):.2f} ms, backward to cp {1e3*(t2-t1):.2f} ms (B=64, T=200)")
