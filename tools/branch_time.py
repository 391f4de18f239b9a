"""ms per inner step of the planner with the optional loss branches (B=64, T=200, bf16 math): plain, speech classifier
(fused into the criterion kernel), somatosensory feedback (three more models on the fp32 step kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paule_b200 as P
dev = torch.device("cuda:0"); torch.manual_seed(0)
pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720).to(dev)
emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720).to(dev)
cp_tube = P.ForwardModel(num_lstm_layers=1, hidden_size=360, output_size=10, input_size=30, apply_half_sequence=False).to(dev)
tube_mel = P.ForwardModel(num_lstm_layers=1, hidden_size=360, output_size=60, input_size=10, apply_half_sequence=True).to(dev)
tube_emb = P.EmbeddingModel(input_size=10, num_lstm_layers=2, hidden_size=720, dropout=0.0).to(dev)
cls = P.LinearClassifier(60, 1).to(dev)
B, T = 64, 200
cp0 = torch.rand(B, T, 30, device=dev) - 0.5
tmel = torch.rand(B, T // 2, 60, device=dev)
for name, kw in (("plain", {}), ("speech classifier", dict(speech_classifier=cls)),
                 ("somatosensory", dict(somatosensory=(cp_tube, tube_mel, tube_emb)))):
    pl = P.BatchPlanner(pred, emb, cp0, tmel, None, max_log_steps=64, math=1, **kw)
    pl.step(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record(); pl.step(n); e1.record(); torch.cuda.synchronize()
    print(f"{name:20s} {e0.elapsed_time(e1) / n:8.2f} ms / inner step", flush=True)
    pl.close()
