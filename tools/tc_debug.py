"""Bring-up check of the tcgen05 persistent-RNN kernels against the fp32 kernels (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paule_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
torch.manual_seed(0)
H = 720
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
lstm = torch.nn.LSTM(30, H, batch_first=True)
w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev), lstm.bias_hh_l0.to(dev), tc=True)
torch.cuda.synchronize()
print("packed bytes", w.packed.numel())
xp = (torch.randn(T, B, 4 * H, device=dev) * 0.5)
st = ops._stream()

def fp32_fwd():
    g = xp.clone(); h = torch.empty(T, B, H, device=dev); c = torch.empty(T, B, H, device=dev)
    _lib.check(lib.paule_lstm_seq_fwd_f32(g.data_ptr(), w.w_hh.data_ptr(), h.data_ptr(), c.data_ptr(), T, B, H, st))
    return g, h, c

def tc_fwd():
    g = xp.clone(); h = torch.zeros(T, B, H, device=dev); c = torch.zeros(T, B, H, device=dev)
    xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
    rc = lib.paule_tc_lstm_seq_fwd(g.data_ptr(), w.packed.data_ptr(), h.data_ptr(), c.data_ptr(), xchg.data_ptr(), None, T, B, 1, st)
    torch.cuda.synchronize()
    print("tc fwd rc", rc, lib.paule_last_cuda_error(), "err flag", xchg[2048:2052].view(torch.int32).item(), "counter", xchg[0:4].view(torch.int32).item())
    return g, h, c

g0, h0, c0 = fp32_fwd()
g1, h1, c1 = tc_fwd()
for t in range(T):
    print(f"t={t} max|dh| {(h1[t]-h0[t]).abs().max().item():.3e} max|dc| {(c1[t]-c0[t]).abs().max().item():.3e} max|dgates| {(g1[t]-g0[t]).abs().max().item():.3e}  |h| {h0[t].abs().max().item():.3f}")
if (h1 - h0).abs().max().item() > 0.05:
    d = (h1[1] - h0[1]).abs()
    print("step-1 error by row (first 16 rows):", d.max(1).values[:16].tolist())
    print("step-1 error by unit (first 16 units):", d.max(0).values[:16].tolist())
    print("h1[1,0,:8]", h1[1, 0, :8].tolist(), "ref", h0[1, 0, :8].tolist())

# backward
dh_seq = torch.randn(T, B, H, device=dev) * 1e-2
dh_last = torch.randn(B, H, device=dev) * 1e-2
def fp32_bwd():
    da = g0.clone(); scratch = torch.empty(B, H, device=dev)
    _lib.check(lib.paule_lstm_seq_bwd_f32(da.data_ptr(), c0.data_ptr(), w.w_hh_t.data_ptr(), dh_seq.data_ptr(), 1, dh_last.data_ptr(), scratch.data_ptr(), T, B, H, st))
    return da
def tc_bwd():
    da = g0.clone()
    xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
    rc = lib.paule_tc_lstm_seq_bwd(da.data_ptr(), c0.data_ptr(), w.packed.data_ptr(), dh_seq.data_ptr(), 1, dh_last.data_ptr(), xchg.data_ptr(), None, T, B, 1, st)
    torch.cuda.synchronize()
    print("tc bwd rc", rc, lib.paule_last_cuda_error(), "err flag", xchg[2048:2052].view(torch.int32).item(), "counter", xchg[0:4].view(torch.int32).item())
    return da
d0 = fp32_bwd(); d1 = tc_bwd()
for t in reversed(range(T)):
    print(f"t={t} max|dda| {(d1[t]-d0[t]).abs().max().item():.3e}  max|da| {d0[t].abs().max().item():.3e}")
# timing
import time
for name, fn in (("fp32_fwd", fp32_fwd), ("tc_fwd", None), ("tc_bwd", None)):
    pass
Tb = 200
xpb = torch.randn(Tb, 64, 4 * H, device=dev) * 0.5
hb = torch.empty(Tb, 64, H, device=dev); cb = torch.empty(Tb, 64, H, device=dev)
xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(64), dtype=torch.uint8, device=dev)
dhb = torch.randn(Tb, 64, H, device=dev) * 1e-2
for name in ("fwd", "bwd"):
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if name == "fwd":
            lib.paule_tc_lstm_seq_fwd(xpb.data_ptr(), w.packed.data_ptr(), hb.data_ptr(), cb.data_ptr(), xchg.data_ptr(), None, Tb, 64, 1, st)
        else:
            lib.paule_tc_lstm_seq_bwd(xpb.data_ptr(), cb.data_ptr(), w.packed.data_ptr(), dhb.data_ptr(), 1, None, xchg.data_ptr(), None, Tb, 64, 1, st)
        e1.record(); torch.cuda.synchronize()
        print(f"tc {name} T={Tb} B=64: {e0.elapsed_time(e1)*1e3/Tb:.2f} us per step; err flag", xchg[2048:2052].view(torch.int32).item())
