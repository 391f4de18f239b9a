"""Warm timing of paule_tc_gemm_img on the shapes of one planning step (B=64, T=200)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paule_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load(); st = ops._stream(); torch.manual_seed(0); H = 720
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for name, T, N, nseg, bias in (("emb1 in-proj  K=768  N=2880", 100, 2880, 1, True), ("post_linear   K=1536 N=60  ", 100, 60, 2, True),
                               ("dX emb1       K=3072 N=720 ", 100, 720, 4, False), ("dX emb0       K=3072 N=60  ", 100, 60, 4, False),
                               ("dX fwd        K=3072 N=30  ", 200, 30, 4, False)):
    img = (torch.randn(lib.paule_tc_img_seq_bytes(T, B, nseg) // 2, device=dev) * 0.1).to(torch.bfloat16).view(torch.uint8)
    W = torch.randn(N, nseg * H, device=dev) / H ** 0.5
    pk = torch.empty(lib.paule_tc_gemm_packed_bytes(N, nseg), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_gemm_pack(W.data_ptr(), pk.data_ptr(), N, nseg, st))
    b = torch.randn(N, device=dev) if bias else None
    out = torch.empty(T, B, N, device=dev)
    steps = T // 2 if nseg == 2 else T
    def run():
        return lib.paule_tc_gemm_img(img.data_ptr(), pk.data_ptr(), b.data_ptr() if bias else None, out.data_ptr(), steps, B, N, nseg, 0, st)
    assert run() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    M, K = steps * B, nseg * 768
    print(f"{name} M={M:6d}: {us:7.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s  out {M*N*4/1e6:6.1f} MB  A {M*K*2/1e6:6.1f} MB", flush=True)
