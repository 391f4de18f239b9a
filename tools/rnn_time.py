"""us per LSTM cell step of the persistent tcgen05 kernels (forward and BPTT), timed alone with CUDA events.
PAULE_RNN_V1=1 selects the v1 kernels.  Usage: python tools/rnn_time.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paule_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load(); torch.manual_seed(0); H = 720
lstm = torch.nn.LSTM(30, H, batch_first=True)
w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev), lstm.bias_hh_l0.to(dev), tc=True)
st = ops._stream(); T = 200
which = "v1" if os.environ.get("PAULE_RNN_V1") else "v2"
for B in [int(a) for a in sys.argv[1:]] or [64, 1, 16, 96, 256]:
    xp = torch.randn(T, B, 4 * H, device=dev) * 0.5
    h = torch.empty(T, B, H, device=dev); c = torch.empty(T, B, H, device=dev)
    dh = torch.randn(T, B, H, device=dev) * 1e-2
    xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
    himg = torch.zeros(lib.paule_tc_img_seq_bytes(T, B, 1), dtype=torch.uint8, device=dev)
    daimg = torch.zeros(lib.paule_tc_img_seq_bytes(T, B, 4), dtype=torch.uint8, device=dev)
    res = {}
    for name, fn in (("fwd", lambda g: lib.paule_tc_lstm_seq_fwd(g.data_ptr(), w.packed.data_ptr(), h.data_ptr(), c.data_ptr(), xchg.data_ptr(), himg.data_ptr(), T, B, 1, st)),
                     ("bwd", lambda g: lib.paule_tc_lstm_seq_bwd(g.data_ptr(), c.data_ptr(), w.packed.data_ptr(), dh.data_ptr(), 1, None, xchg.data_ptr(), daimg.data_ptr(), T, B, 1, st))):
        best = 1e9
        for rep in range(4):
            g = xp.clone()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); rc = fn(g); e1.record(); torch.cuda.synchronize()
            assert rc == 0, rc
            best = min(best, e0.elapsed_time(e1) * 1e3 / T)
        res[name] = best
        err = xchg[2048:2052].view(torch.int32).item()
        assert err == 0, f"watchdog {err}"
    print(f"{which} B={B:4d} T={T}: fwd {res['fwd']:.2f} us/step  bwd {res['bwd']:.2f} us/step  "
          f"({B * T / (res['fwd'] * T) :.1f} / {B * T / (res['bwd'] * T):.1f} word-steps/us)", flush=True)
