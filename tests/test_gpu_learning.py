"""Continue-learning of the predictive forward model on the GPU (reference: paule/paule.py:1361-1377): weight
gradients of the CUDA LSTM / Linear ops against torch autograd on the reference's modules, the learning loop against the
same loop in plain torch on the CPU, and the outer-loop hook of plan_resynth.  Run on the B200 box: python -m pytest -m gpu"""
import copy

import numpy as np
import pytest
import torch

from oracle import paule_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from paule_b200 import _lib
    _lib.require_device()
    return torch.device("cuda:0")


def _pair(hidden, layers, seed):
    """(CUDA ForwardModel, reference-arithmetic CPU twin) with identical weights."""
    import paule_b200 as P
    torch.manual_seed(seed)
    mine = P.ForwardModel(num_lstm_layers=layers, hidden_size=hidden)
    ref = _TorchForward(layers, hidden)
    ref.load_state_dict(mine.state_dict())
    return mine, ref


class _TorchForward(torch.nn.Module):
    """paule/models.py:326-356 restated with torch.nn (the oracle's arithmetic)."""

    def __init__(self, layers, hidden):
        super().__init__()
        self.half_sequence = torch.nn.AvgPool1d(2, stride=2)
        self.lstm = torch.nn.LSTM(30, hidden, num_layers=layers, batch_first=True)
        self.post_linear = torch.nn.Linear(hidden, 60)

    def forward(self, x, *args):
        out, _ = self.lstm(x)
        out = self.post_linear(out)
        return self.half_sequence(out.permute(0, 2, 1)).permute(0, 2, 1)


@pytest.mark.parametrize("hidden,layers,B,T", [(64, 2, 3, 20), (720, 1, 2, 14)])
def test_weight_gradients_match_torch_autograd(dev, hidden, layers, B, T):
    mine, ref = _pair(hidden, layers, 1)
    mine = mine.to(dev).train()
    g = torch.Generator().manual_seed(2)
    x = torch.rand(B, T, 30, generator=g) - 0.5
    y = torch.rand(B, T // 2, 60, generator=g)
    torch.sqrt(torch.mean((ref(x) - y) ** 2)).backward()
    out = mine(x.to(dev))
    loss = torch.sqrt(torch.mean((out - y.to(dev)) ** 2))
    loss.backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref(x).detach().numpy(), atol=2e-5)
    for (n, p), (_, q) in zip(mine.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        scale = q.grad.abs().max().item() + 1e-12
        np.testing.assert_allclose(p.grad.cpu().numpy(), q.grad.numpy(), atol=2e-4 * scale, rtol=2e-3, err_msg=n)


def test_continue_learning_pred_matches_the_same_loop_in_torch(dev):
    import paule_b200 as P
    mine, ref = _pair(96, 1, 3)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=96)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=96)
    pm = P.Paule(pred_model=mine, inv_model=inv, embedder=emb, device=dev)
    g = torch.Generator().manual_seed(4)
    lens = [20, 20, 20, 32, 32, 20, 32]
    cps = [(torch.rand(L, 30, generator=g) - 0.5).numpy() for L in lens]
    mels = [torch.rand(L // 2, 60, generator=g).numpy() for L in lens]
    losses = pm.continue_learning_pred(cps, mels, n_epochs=3, batch_size=3, shuffle=False)
    # the same loop on the CPU: lengths ascending, batches of <= 3 consecutive samples of one length, Adam(lr=1e-3)
    opt = torch.optim.Adam(ref.parameters(), lr=0.001)
    ref_losses = []
    for _ in range(3):
        ep = []
        for L in sorted(set(lens)):
            idx = [i for i, l in enumerate(lens) if l == L]
            for k in range(0, len(idx), 3):
                j = idx[k:k + 3]
                xb = torch.stack([torch.from_numpy(cps[i]) for i in j])
                yb = torch.stack([torch.from_numpy(mels[i]) for i in j])
                opt.zero_grad()
                l = torch.sqrt(torch.mean((ref(xb) - yb) ** 2))
                l.backward()
                opt.step()
                ep.append(float(l.detach()))
        ref_losses.append(float(np.mean(ep)))
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-4)
    assert losses[-1] < losses[0]
    for (n, p), (_, q) in zip(pm.pred_model.named_parameters(), ref.named_parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().numpy(), atol=2e-5, err_msg=n)


@pytest.mark.parametrize("math", [0, 1])
def test_plan_resynth_outer_loop_learns_and_repacks(dev, math):
    """continue_learning=True with a host-side synthesizer: the produced mels train pred_model between the outer iterations and
    the planner plans the next iteration with the updated weights."""
    import paule_b200 as P
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=720)
    rng = np.random.RandomState(0)
    proj = rng.randn(30, 60).astype(np.float32) * 0.2

    def synth(cp):   # deterministic stand-in for VocalTractLab + librosa: pooled frames through a fixed map
        pooled = 0.5 * (cp[0::2][: cp.shape[0] // 2] + cp[1::2][: cp.shape[0] // 2])
        return np.tanh(pooled @ proj) + 0.5

    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, math=math, synthesizer=synth)
    before = copy.deepcopy(pred.state_dict())
    _, tmel = O.synthetic_inputs(4, 40, seed=8)
    cp0, _ = O.synthetic_inputs(4, 40, seed=9)
    res = pm.plan_resynth(target_acoustic=tmel.numpy(), initial_cp=cp0.numpy(), initialize_from=None, objective="acoustic_semvec",
                          n_outer=2, n_inner=3, continue_learning=True, n_epochs=2, batch_size=2, verbose=False)
    assert len(res.pred_model_loss) == 4 and all(np.isfinite(res.pred_model_loss))
    assert res.pred_model_loss[-1] < res.pred_model_loss[0]
    after = pred.state_dict()
    assert any(not torch.equal(before[k].cpu(), after[k].cpu()) for k in before)
    assert np.isfinite(res.planned_cp).all() and len(res.planned_loss_steps) == 6
    # the planner now holds the UPDATED weights: its prediction for the planned cps == the updated module's own forward
    with torch.no_grad():
        direct = pred(torch.from_numpy(res.planned_cp).to(dev)).cpu().numpy()
    np.testing.assert_allclose(res.pred_mel, direct, atol=1e-5 if math == 0 else 5e-3)


def test_host_synthesis_pipeline_fills_the_produced_fields(dev):
    """with a host-side synthesizer the planned cps are synthesised at every outer-loop boundary on a thread
    pool (overlapped with the GPU's next outer iteration) and PlanningResults carries the produced side: prod_mel,
    prod_semvec, 5 rmse / 10 rmse losses per synthesis (paule.py:1109-1111, :1139-1146), best synthesis per word (:1158-1164)."""
    import threading
    import time
    import paule_b200 as P
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=720)
    proj = np.random.RandomState(1).randn(30, 60).astype(np.float32) * 0.2
    threads = set()

    def synth(cp):
        threads.add(threading.get_ident())
        time.sleep(0.02)                                   # a (very fast) VocalTractLab
        pooled = 0.5 * (cp[0::2][: cp.shape[0] // 2] + cp[1::2][: cp.shape[0] // 2])
        return (np.zeros(8, dtype=np.float32), 44100, np.tanh(pooled @ proj) + 0.5)     # (sig, sr, mel)

    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, synthesizer=synth, math=0)
    _, tmel = O.synthetic_inputs(5, 40, seed=8)
    cp0, _ = O.synthetic_inputs(5, 40, seed=9)
    res = pm.plan_resynth(target_acoustic=tmel.numpy(), initial_cp=cp0.numpy(), initialize_from=None, objective="acoustic_semvec",
                          n_outer=3, n_inner=2, continue_learning=False, log_signals=True, verbose=False)
    assert len(threads) > 1, "synthesis jobs must run on the pool, not serially on the caller's thread"
    assert len(res.prod_loss_steps) == 3 and len(res.prod_semvec_loss_steps) == 3 and len(res.sig_steps) == 3
    for b in range(5):
        np.testing.assert_allclose(res.prod_mel[b], synth(res.planned_cp[b])[2], atol=1e-6)
        np.testing.assert_allclose(res.initial_prod_mel[b], synth(cp0[b].numpy())[2], atol=1e-6)
        want = 5.0 * np.sqrt(np.mean((res.prod_mel[b] - tmel[b].numpy()) ** 2))
        np.testing.assert_allclose(res.prod_loss_steps[-1][b], want, rtol=1e-5)
        assert pm.best_synthesis_acoustic[b].mel_loss == pytest.approx(min(s[b] for s in res.prod_loss_steps), rel=1e-6)
        assert pm.best_synthesis_semantic[b].semvec_loss == pytest.approx(min(s[b] for s in res.prod_semvec_loss_steps), rel=1e-6)
    with torch.no_grad():
        sv = emb(torch.from_numpy(np.stack(res.prod_mel)).to(dev), [20] * 5).cpu().numpy()
    np.testing.assert_allclose(res.prod_semvec, sv, atol=1e-5)
    assert res.prod_sr == 44100 and res.pred_model_loss == []
    # single word: reference-shaped (unbatched) produced fields
    one = pm.plan_resynth(target_acoustic=tmel[0].numpy(), initial_cp=cp0[0].numpy(), initialize_from=None, objective="acoustic",
                          n_outer=2, n_inner=2, continue_learning=False, verbose=False)
    assert one.prod_mel.shape == (20, 60) and one.prod_semvec.shape == (300,) and np.ndim(one.prod_loss_steps[0]) == 0


# ---- continue-learning of the inverse model (paule/paule.py:1413-1436) ------------------------------------------------------
def test_own_weight_gradient_kernels_match_matmul(dev):
    """paule_gemm_tn_f32 / paule_colsum_f32 (dW = dA^T X, db = column sums over all (step, word) rows) against torch, incl. the
    split-reduction path (few output tiles) and ragged sizes."""
    from paule_b200 import ops
    g = torch.Generator().manual_seed(0)
    for R, M, N in ((1200, 2880, 30), (600, 2880, 720), (37, 70, 5), (4000, 60, 720), (1, 8, 8)):
        a = (torch.rand(R, M, generator=g) - 0.5).to(dev)
        b = (torch.rand(R, N, generator=g) - 0.5).to(dev)
        want = a.double().t() @ b.double()
        np.testing.assert_allclose(ops.gemm_tn(a, b).cpu().double().numpy(), want.cpu().numpy(), rtol=2e-5, atol=2e-5 * R ** 0.5)
        np.testing.assert_allclose(ops.colsum(a).cpu().double().numpy(), a.double().sum(0).cpu().numpy(), rtol=2e-5, atol=2e-5 * R ** 0.5)


def test_inverse_model_learning_step_matches_the_oracle(dev):
    """One epoch of Paule.continue_learning_inv (differentiable inverse-model forward + cp_trajectory_loss + Adam) == the same
    loop on the oracle's inverse model on the CPU: losses and every updated parameter."""
    import paule_b200 as P
    torch.manual_seed(3)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=64)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=64)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=64)
    ref = O.OracleInverseModel(num_lstm_layers=1, hidden_size=64)
    ref.load_state_dict(inv.state_dict())
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, math=0)
    pm.inv_optimizer = torch.optim.Adam(pm.inv_model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(4)
    mels = [torch.rand(n, 60, generator=g).numpy() for n in (12, 12, 12, 12)]
    cps = [(torch.rand(2 * m.shape[0], 30, generator=g) - 0.5).numpy() for m in mels]
    # the planning-time forward (hand-written stencil kernels, no autograd graph) and the differentiable one agree
    x = torch.from_numpy(np.stack(mels)).to(dev)
    with torch.no_grad():
        y_kernels = pm.inv_model(x)
    pm.inv_model.learnable = True
    y_train = pm.inv_model(x)
    pm.inv_model.learnable = False
    assert y_train.requires_grad and not y_kernels.requires_grad
    np.testing.assert_allclose(y_train.detach().cpu().numpy(), y_kernels.cpu().numpy(), atol=2e-5)
    # gradients of cp_trajectory_loss wrt every parameter: differentiable forward + own BPTT / weight-gradient kernels vs
    # autograd on the oracle's module (Adam-updated weights are not compared: Adam turns a sign flip of a near-zero gradient
    # into a +-lr step)
    for p in pm.inv_model.parameters():
        p.requires_grad_(True)
    pm.inv_model.learnable = True
    loss = P.Paule.cp_trajectory_loss(pm.inv_model(x), torch.from_numpy(np.stack(cps)).to(dev))[0]
    loss.backward()
    pm.inv_model.learnable = False
    for p in ref.parameters():
        p.requires_grad_(True)
    want = P.Paule.cp_trajectory_loss(ref(torch.from_numpy(np.stack(mels))), torch.from_numpy(np.stack(cps)))[0]
    want.backward()
    np.testing.assert_allclose(float(loss.detach()), float(want.detach()), rtol=1e-5)
    for (n, p), (_, q) in zip(pm.inv_model.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, n
        np.testing.assert_allclose(p.grad.cpu().numpy(), q.grad.numpy(), atol=2e-4 * float(q.grad.abs().max()) + 1e-7, err_msg=n)
    # the learning loop itself: the loss falls
    pm.inv_optimizer = torch.optim.Adam(pm.inv_model.parameters(), lr=1e-3)
    losses = pm.continue_learning_inv(mels, cps, n_epochs=6, batch_size=4, shuffle=False)
    assert len(losses) == 6 and losses[-1] < losses[0]
    np.testing.assert_allclose(losses[0], float(want.detach()), rtol=1e-4)


def test_audio_target_goes_through_the_host_mel_front_end(dev):
    """plan_resynth(target_acoustic=(signal, rate)) as in the reference (paule/paule.py:495-496, :523-529): the target mel is the
    normalised, min-shifted log-mel of the waveform (paule_b200/audio.py), target_sig / target_sr come back in the results."""
    import paule_b200 as P
    from paule_b200 import audio as A
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=64).to(dev)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=64).to(dev)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=64).to(dev)
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev)
    sr = 44100
    t = np.arange(int(0.1 * sr)) / sr
    sig = 0.2 * np.sin(2 * np.pi * 220.0 * t) * np.hanning(len(t))
    res = pm.plan_resynth(target_acoustic=(sig, sr), initialize_from="acoustic", objective="acoustic_semvec", n_outer=1, n_inner=3,
                          continue_learning=False, verbose=False)
    want = A.target_mel_from_audio(sig, sr)
    assert res.target_mel.shape == want.shape == (1 + len(sig) // 220, 60)
    np.testing.assert_allclose(res.target_mel, want, atol=1e-5)
    assert res.target_sr == sr and np.array_equal(res.target_sig, sig)
    assert res.planned_cp.shape == (2 * want.shape[0], 30) and len(res.planned_loss_steps) == 3
