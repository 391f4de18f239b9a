"""Timeline of one CTA of the persistent forward kernel (build with NVCC_EXTRA=-DPAULE_TC_TIMELINE).  Usage: tc_timeline.py B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paule_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load(); torch.manual_seed(0); H = 720
lstm = torch.nn.LSTM(30, H, batch_first=True)
w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev), lstm.bias_hh_l0.to(dev), tc=True)
st = ops._stream(); T = 100
NAMES = {1: "L visit", 2: "L fetched", 3: "L mma issued", 10: "E wait", 11: "E acc ready", 12: "E published q", 13: "E stash q",
         14: "E partials pushed", 15: "E partials arrived"}
for B in [int(a) for a in sys.argv[1:]] or [384]:
    x = (torch.rand(T, B, 30, device=dev) - 0.5)
    ximg = torch.zeros(lib.paule_tc_x_image_bytes(T, B), dtype=torch.uint8, device=dev)
    lib.paule_tc_x_image(x.data_ptr(), ximg.data_ptr(), T, B, 30, st)
    g = torch.empty(T, B, 4 * H, device=dev); h = torch.empty(T, B, H, device=dev); c = torch.empty(T, B, H, device=dev)
    for rep in range(2):
        xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.paule_tc_lstm_seq_fwd_x(g.data_ptr(), w.packed.data_ptr(), w.bias.data_ptr(), ximg.data_ptr(), h.data_ptr(), c.data_ptr(),
                                         xchg.data_ptr(), None, T, B, 1, st)
        e1.record(); torch.cuda.synchronize()
    ev = [r for r in xchg[3072:4096].view(torch.int64).cpu().reshape(-1, 2).tolist() if r[1] != 0]
    n = len(ev)
    ev.sort(key=lambda r: r[1])
    t0 = ev[0][1] if ev else 0
    print(f"B={B} fused fwd {e0.elapsed_time(e1)*1e3/T:.2f} us/step, rc={rc}, {n} events")
    for tag, ts in ev:
        e, slot, step = tag >> 32, (tag >> 16) & 0xffff, tag & 0xffff
        print(f"  {(ts - t0)/1e3:8.2f} us  t={step} {'slot' if e < 12 else 'quarter'} {slot}  {NAMES.get(e, e)}")

    # backward
    dh = torch.randn(T, B, H, device=dev) * 1e-2
    daimg = torch.zeros(lib.paule_tc_img_seq_bytes(T, B, 4), dtype=torch.uint8, device=dev)
    for rep in range(2):
        xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
        gg = g.clone()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.paule_tc_lstm_seq_bwd_img(gg.data_ptr(), c.data_ptr(), w.packed.data_ptr(), dh.data_ptr(), 1, None, xchg.data_ptr(), daimg.data_ptr(), T, B, 1, st)
        e1.record(); torch.cuda.synchronize()
    ev = [r for r in xchg[3072:4096].view(torch.int64).cpu().reshape(-1, 2).tolist() if r[1] != 0]
    n = len(ev)
    ev.sort(key=lambda r: r[1])
    t0 = ev[0][1] if ev else 0
    print(f"B={B} bwd {e0.elapsed_time(e1)*1e3/T:.2f} us/step, rc={rc}, {n} events")
    for tag, ts in ev:
        e, slot, step = tag >> 32, (tag >> 16) & 0xffff, tag & 0xffff
        print(f"  {(ts - t0)/1e3:8.2f} us  it={step} q={slot}  {NAMES.get(e, e)}")
