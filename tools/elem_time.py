"""Achieved HBM bandwidth of the fused elementwise kernels (loss + analytic gradients, Adam + clamp) against the measured
peak, at BASELINE shapes.  Algorithmic bytes: loss pass reads cp, mel, tmel and writes dmel, dcp_smooth
= (360 + 240) T B per word; Adam reads x, g_lstm, g_smooth, m, v and writes x, m, v = 960 T B per word."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paule_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load(); st = ops._stream()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(96 << 20, device=dev)   # 384 MB > L2
for B, T in ((64, 200), (256, 400), (1024, 400), (512, 1200)):
    Tm, C, Cm, S = T // 2, 30, 60, 300
    cp = torch.rand(T, B, C, device=dev) - 0.5
    mel, tmel = torch.rand(Tm, B, Cm, device=dev), torch.rand(Tm, B, Cm, device=dev)
    sv, tsv = torch.rand(B, S, device=dev), torch.rand(B, S, device=dev)
    terms = torch.empty(B, 6, device=dev); dmel = torch.empty_like(mel); dsv = torch.empty_like(sv); dcp = torch.empty_like(cp)
    scratch = torch.empty(lib.paule_plan_loss_scratch_floats(T, B), device=dev)
    g2 = torch.randn_like(cp) * 1e-3; m = torch.zeros_like(cp); v = torch.zeros_like(cp)
    step = torch.ones(1, dtype=torch.int32, device=dev)
    def loss():
        return lib.paule_plan_loss_f32(mel.data_ptr(), tmel.data_ptr(), sv.data_ptr(), tsv.data_ptr(), cp.data_ptr(), terms.data_ptr(),
                                       dmel.data_ptr(), dsv.data_ptr(), dcp.data_ptr(), scratch.data_ptr(), T, Tm, B, C, Cm, S, 0, st)
    def adam():
        return lib.paule_adam_clamp_f32(cp.data_ptr(), dcp.data_ptr(), g2.data_ptr(), m.data_ptr(), v.data_ptr(), step.data_ptr(),
                                        0.01, 0.9, 0.999, 1e-8, 1.05, 0, None, 0, T, B, C, st)
    for name, fn, nbytes in (("loss+grad (2 kernels)", loss, (360 + 240) * T * B + 3 * S * 4 * B), ("adam+clamp", adam, 960 * T * B)):
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); assert fn() == 0; e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        t = min(ts)
        print(f"B={B:5d} T={T:5d} {name:22s} {t*1e6:8.1f} us  {nbytes/1e6:8.2f} MB algorithmic  {nbytes/t/1e9:7.0f} GB/s = {100*nbytes/t/1e9/peak:5.1f}% of {peak:.0f}", flush=True)
