"""Times paule_linear_f32 on the shapes of one planning step (B=64, T=200) and checks them against torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paule_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load(); st = ops._stream(); torch.manual_seed(0)
for name, M, N, K, acc in (("gates_f  K=30", 12800, 2880, 30, 0), ("gates_0  K=60", 6400, 2880, 60, 0), ("head     M=64", 64, 300, 720, 0),
                           ("dsv->dh1 M=64", 64, 720, 300, 0), ("dmel->dhp K=60", 6400, 720, 60, 0), ("acc K=60", 6400, 720, 60, 1),
                           ("ragged", 777, 250, 45, 0)):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    C = torch.randn(M, N, device=dev); C0 = C.clone()
    def run():
        return lib.paule_linear_f32(A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 1, K, 0, 0, 1, N, 0, acc, st)
    assert run() == 0
    ref = A.double() @ W.double().t() + b.double() + (C0.double() if acc else 0)
    err = (C.double() - ref).abs().max().item()
    C.copy_(C0)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    big = torch.empty(64 << 20, device=dev)
    ts = []
    for _ in range(5):
        big.zero_()     # flush L2
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        if acc: C.copy_(C0)
    print(f"{name:16s} M={M:6d} N={N:5d} K={K:4d}: {min(ts):8.1f} us   max err {err:.2e}   write {M*N*4/1e6:.0f} MB", flush=True)
