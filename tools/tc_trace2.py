import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paule_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load(); torch.manual_seed(0); H = 720
lstm = torch.nn.LSTM(30, H, batch_first=True)
w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev), lstm.bias_hh_l0.to(dev), tc=True)
st = ops._stream(); T = 200
for B in [int(a) for a in sys.argv[1:]] or [64, 1]:
    xpb = torch.randn(T, B, 4 * H, device=dev) * 0.5
    hb = torch.empty(T, B, H, device=dev); cb = torch.empty(T, B, H, device=dev)
    xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
    for rep in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        g = xpb.clone()
        e0.record()
        lib.paule_tc_lstm_seq_fwd(g.data_ptr(), w.packed.data_ptr(), hb.data_ptr(), cb.data_ptr(), xchg.data_ptr(), None, T, B, 1, st)
        e1.record(); torch.cuda.synchronize()
    print(f"B={B} fwd {e0.elapsed_time(e1)*1e3/T:.2f} us/step, err", xchg[2048:2052].view(torch.int32).item())
    tr = xchg[3072:3072 + 16 * 8].view(torch.int64).cpu().tolist()
    for who, base, names in (("loader kb3", 0, ["first elem visible", "bulk fetch", "fence+syncwarp", "acc_free wait", "4 mma + commit", "commit->mma_done", "(probe phase of fetch)", "(bulk passes)"]),
                             ("epilogue", 8, ["wait mma_done", "tmem ld", "zero+arrive", "transpose+cell+ll_store", "stash stores"])):
        print(" ", who, {n: round(v / T / 1e3, 3) for n, v in zip(names, tr[base:base + 8])})

    # backward
    dh = torch.randn(T, B, H, device=dev) * 1e-2
    for rep in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        g = xpb.clone()
        e0.record()
        lib.paule_tc_lstm_seq_bwd(g.data_ptr(), cb.data_ptr(), w.packed.data_ptr(), dh.data_ptr(), 1, None, xchg.data_ptr(), None, T, B, 1, st)
        e1.record(); torch.cuda.synchronize()
    print(f"B={B} bwd {e0.elapsed_time(e1)*1e3/T:.2f} us/step, err", xchg[2048:2052].view(torch.int32).item())
    tr = xchg[3072:3072 + 16 * 8].view(torch.int64).cpu().tolist()
    for who, base, names in (("loader kb3", 0, ["until probes pass", "bulk", "fence+acc_free", "4 mma + commit", "commit->mma_done"]),
                             ("epilogue", 8, ["prefetch loads issue", "wait mma_done", "tmem ld+zero", "dsmem push+arrive", "wait red_full", "sum+adjoint+xchg stores", "image+stash stores"])):
        print(" ", who, {n: round(v / T / 1e3, 3) for n, v in zip(names, tr[base:base + 8])})
