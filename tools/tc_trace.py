import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paule_b200 import _lib, ops
dev = torch.device("cuda:0"); lib = _lib.load(); torch.manual_seed(0); H = 720
lstm = torch.nn.LSTM(30, H, batch_first=True)
w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev), lstm.bias_hh_l0.to(dev), tc=True)
st = ops._stream(); Tb = 200
xpb = torch.randn(Tb, 64, 4 * H, device=dev) * 0.5
hb = torch.empty(Tb, 64, H, device=dev); cb = torch.empty(Tb, 64, H, device=dev)
xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(64), dtype=torch.uint8, device=dev)
for rep in range(2):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.paule_tc_lstm_seq_fwd(xpb.data_ptr(), w.packed.data_ptr(), hb.data_ptr(), cb.data_ptr(), xchg.data_ptr(), None, Tb, 64, 1, st)
    e1.record(); torch.cuda.synchronize()
    print(f"fwd {e0.elapsed_time(e1)*1e3/Tb:.2f} us/step, err", xchg[2048:2052].view(torch.int32).item())
    tr = xchg[3072:3072 + 24 * 8].view(torch.int64).cpu().tolist()
    for who, base, names in (("producer", 0, ["kblock wait", "copy issue"]), ("mma", 8, ["wait kblock", "issue 4 mma + commit"]),
                             ("epilogue", 16, ["wait mma_done", "tmem ld", "cell + h store", "bar + arrive", "stash stores"])):
        print(" ", who, {n: round(v / Tb / 1e3, 3) for n, v in zip(names, tr[base:base + 8])})
