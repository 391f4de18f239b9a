"""Optional loss branches of the planner on the GPU: the speech-classifier term fused into the criterion
kernel and the somatosensory cp -> tube -> (mel, semvec) branch, against

* vectors of the REAL reference's ``plan_resynth(use_speech_classifier=True / use_somatosensory_feedback=True)``
  (``tests/golden/make_branches_golden.py``, fp64, batch 1), through our ``Paule.plan_resynth``;
* the CPU oracle's batched loop with per-step gradients (``oracle.plan_inner_loop_branches``), teacher-free, B=3.

Tolerances as in test_gpu_planner.py: fp32 math 1e-4 relative loss / 1e-5 absolute cps; bf16 math 1e-3 / 1e-3.
"""
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


@pytest.fixture(scope="module")
def models(dev, golden):
    import paule_b200 as P
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=720)
    assert [O.state_dict_digest(m) for m in (pred, emb, inv)] == list(golden["digest32"])
    return pred.to(dev), emb.to(dev), inv.to(dev)


@pytest.fixture(scope="module")
def branch_models(dev, golden_branches):
    """same seeds / construction order as make_branches_golden.py::branch_models, from OUR module classes"""
    import paule_b200 as P
    torch.manual_seed(1)
    cp_tube = P.ForwardModel(num_lstm_layers=1, hidden_size=360, output_size=10, input_size=30, apply_half_sequence=False)
    tube_mel = P.ForwardModel(num_lstm_layers=1, hidden_size=360, output_size=60, input_size=10, apply_half_sequence=True)
    tube_emb = P.EmbeddingModel(input_size=10, num_lstm_layers=2, hidden_size=720, dropout=0.0, post_upsampling_size=0)
    torch.manual_seed(2)
    cls = P.LinearClassifier(input_dim=60, output_dim=1)
    with torch.no_grad():
        cls.linear.weight.mul_(30.0)
    np.testing.assert_array_equal(cls.linear.weight.detach().numpy(), golden_branches["cls_w"].astype(np.float32))
    return cp_tube.to(dev), tube_mel.to(dev), tube_emb.to(dev), cls.to(dev)


def _np(t):
    return t.detach().cpu().double().numpy()


TOL = {0: (1e-4, 2e-4, 1e-5), 1: (1e-3, 2e-2, 1e-3)}   # (loss rtol, grad rel-to-max, cp atol) per math mode


@pytest.mark.parametrize("objective", ["acoustic_semvec", "acoustic", "semvec"])
def test_classifier_plan_resynth_matches_the_real_reference(dev, models, branch_models, golden_branches, objective):
    import paule_b200 as P
    g = golden_branches
    pred, emb, inv = models
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, use_speech_classifier=True,
                 speech_classifier=branch_models[3], math=0)
    tag = f"cls_{objective}"
    n = len(g[f"{tag}_loss"])
    res = pm.plan_resynth(target_acoustic=g["tmel"][0].astype(np.float32), initial_cp=g["cp0"][0].astype(np.float32),
                          initialize_from=None, objective=objective, n_outer=1, n_inner=n, log_ii=1,
                          continue_learning=False, verbose=False, log_semantics=False)
    assert type(res).__name__ == "PlanningResultsWithSpeechClassifier"
    np.testing.assert_allclose(res.planned_loss_steps, g[f"{tag}_loss"], rtol=1e-4)
    np.testing.assert_allclose(res.pred_speech_classifier_loss_steps, g[f"{tag}_cls"], rtol=1e-4)
    np.testing.assert_allclose(res.planned_cp, g[f"{tag}_planned_cp"], atol=1e-5)
    assert res.prod_speech_classifier_loss_steps == []


def _ramp_cps(B, T):
    """cps linear in time: local-linear and jerk terms vanish and the velocity gradient lives on the edges only, so the
    gradient is dominated by the model paths and a branch's own contribution is not lost in fp32 cancellation."""
    t = torch.arange(T, dtype=torch.float32).view(1, T, 1)
    c = torch.arange(30, dtype=torch.float32).view(1, 1, 30)
    b = torch.arange(B, dtype=torch.float32).view(B, 1, 1)
    return (0.3 * torch.sin(c + b) + 0.002 * t * torch.cos(c + 0.5 * b)).contiguous()


@pytest.mark.parametrize("math", [0, 1])
def test_classifier_terms_and_gradient_match_the_oracle(dev, models, branch_models, golden, math):
    """B=3, free-running: logged terms and cps vs the CPU oracle; then the classifier's own contribution to d(loss)/d(cp)
    (gradient with minus gradient without the classifier, on ramp cps) vs the same difference of the oracle."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    cls = branch_models[3]
    cp0, tmel = torch.from_numpy(golden["b3_cp0"]), torch.from_numpy(golden["b3_tmel"])
    p32, e32, _ = O.build_reference_models(0, 720, torch.float32, with_inverse=False)
    ocls = O.build_branch_models(torch.float32)[3]
    lr, _, ca = TOL[math]
    for objective in ("acoustic_semvec", "semvec"):
        r1 = O.plan_inner_loop_branches(p32, e32, cp0, tmel, 3, objective=objective, speech_classifier=ocls)
        pl = BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=4, math=math, objective=objective,
                          speech_classifier=cls)
        pl.step(3)
        L = pl.losses()
        np.testing.assert_allclose(_np(L["total"]), r1["loss"].numpy(), rtol=lr)
        np.testing.assert_allclose(_np(L["speech_classifier"]), r1["aux"][:, :, 0].numpy(), rtol=max(lr, 2e-4))
        np.testing.assert_allclose(_np(pl.planned_cp()), r1["planned_cp"].numpy(), atol=ca)
    ramp = _ramp_cps(3, cp0.shape[1])
    w1 = O.plan_inner_loop_branches(p32, e32, ramp, tmel, 1, speech_classifier=ocls, log_grads=True)["grads"][0]
    w0 = O.plan_inner_loop(p32, e32, ramp, tmel, 1, log_grads=True)["grads"][0]
    want = (w1 - w0).double().numpy()
    got = []
    for c in (cls, None):
        pl = BatchPlanner(pred, emb, ramp.to(dev), tmel.to(dev), None, log_gradients=True, max_log_steps=2, math=math,
                          use_cuda_graph=False, speech_classifier=c)
        pl.step(1)
        got.append(_np(pl.last_grad()))
    scale = np.abs(want).max()
    assert scale > 1e-5, scale
    np.testing.assert_allclose(got[0] - got[1], want, atol=(0.02 if math == 0 else 0.15) * scale)


@pytest.mark.parametrize("per_step", [False, True])   # one graph replay per call / n_inner replays in one call
def test_somatosensory_plan_resynth_matches_the_real_reference(dev, models, branch_models, golden_branches, per_step):
    import paule_b200 as P
    g = golden_branches
    pred, emb, inv = models
    cp_tube, tube_mel, tube_emb, _ = branch_models
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, use_somatosensory_feedback=True,
                 cp_tube_model=cp_tube, tube_mel_model=tube_mel, tube_embedder=tube_emb, math=0)
    n = len(g["soma_loss"])
    res = pm.plan_resynth(target_acoustic=g["tmel"][0].astype(np.float32), initial_cp=g["cp0"][0].astype(np.float32),
                          initialize_from=None, objective="acoustic_semvec", n_outer=1, n_inner=n, log_ii=1,
                          continue_learning=False, verbose=False, log_semantics=False, log_cps=per_step)
    assert type(res).__name__ == "PlanningResultsWithSomatosensory"
    np.testing.assert_allclose(res.planned_loss_steps, g["soma_loss"], rtol=1e-4)
    np.testing.assert_allclose(res.planned_mel_loss_steps, g["soma_mel"], rtol=1e-4)
    np.testing.assert_allclose(res.pred_semvec_loss_steps, g["soma_sem"], rtol=1e-4)
    np.testing.assert_allclose(res.pred_tube_mel_loss_steps, g["soma_tube_mel"], rtol=1e-4)
    np.testing.assert_allclose(res.pred_tube_semvec_loss_steps, g["soma_tube_sem"], rtol=1e-4)
    np.testing.assert_allclose(res.planned_cp, g["soma_planned_cp"], atol=1e-5)
    np.testing.assert_allclose(res.pred_tube, g["soma_pred_tube"], atol=2e-5)
    np.testing.assert_allclose(res.pred_tube_mel, g["soma_pred_tube_mel"], atol=2e-5)
    np.testing.assert_allclose(res.pred_tube_semvec, g["soma_pred_tube_semvec"], atol=2e-5)


@pytest.mark.parametrize("math", [0, 1])
def test_somatosensory_gradient_and_batch_match_the_oracle(dev, models, branch_models, golden, math):
    """B=3: per-step total gradient (main path + tube branch), loss terms and cps vs the CPU oracle; the tube branch's own
    gradient (with minus without) is checked on its own scale."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    cp_tube, tube_mel, tube_emb, _ = branch_models
    cp0, tmel = torch.from_numpy(golden["b3_cp0"]), torch.from_numpy(golden["b3_tmel"])
    p32, e32, _ = O.build_reference_models(0, 720, torch.float32, with_inverse=False)
    oct_, otm, ote, _ = O.build_branch_models(torch.float32)
    n = 3
    r1 = O.plan_inner_loop_branches(p32, e32, cp0, tmel, n, cp_tube_model=oct_, tube_mel_model=otm, tube_embedder=ote,
                                    log_grads=True)
    lr, gr, ca = TOL[math]
    pl = BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, log_gradients=True, max_log_steps=n, math=math,
                      use_cuda_graph=False, somatosensory=(cp_tube, tube_mel, tube_emb))
    g_first = None
    for k in range(n):
        pl.step(1)
        gk = _np(pl.last_grad())
        if k == 0:
            g_first = gk
            np.testing.assert_allclose(gk, r1["grads"][0].numpy(), rtol=gr, atol=gr * r1["grads"][0].abs().max().item())
    L = pl.losses()
    np.testing.assert_allclose(_np(L["total"]), r1["loss"].numpy(), rtol=lr)
    # (bf16 mode: the three tube models run on the tcgen05 kernels, the 360-unit ones zero-padded to 720 units)
    assert pl.soma.tc == (math == 1)
    np.testing.assert_allclose(_np(L["tube_mel"]), r1["aux"][:, :, 1].numpy(), rtol=1e-4 if math == 0 else 2e-3)
    np.testing.assert_allclose(_np(L["tube_semvec"]), r1["aux"][:, :, 2].numpy(), rtol=1e-4 if math == 0 else 2e-3)
    np.testing.assert_allclose(_np(pl.planned_cp()), r1["planned_cp"].numpy(), atol=ca)
    # the branch's own gradient = extra_grad of the first step
    ramp = _ramp_cps(3, cp0.shape[1])
    w1 = O.plan_inner_loop_branches(p32, e32, ramp, tmel, 1, cp_tube_model=oct_, tube_mel_model=otm, tube_embedder=ote,
                                    log_grads=True)["grads"][0]
    w0 = O.plan_inner_loop(p32, e32, ramp, tmel, 1, log_grads=True)["grads"][0]
    want = (w1 - w0).double().numpy()
    pl2 = BatchPlanner(pred, emb, ramp.to(dev), tmel.to(dev), None, max_log_steps=2, math=math, use_cuda_graph=False,
                       somatosensory=(cp_tube, tube_mel, tube_emb))
    pl2.soma.run(pl2.cp, pl2.target_mel, pl2.target_sv)
    from paule_b200 import ops
    got = _np(ops.transpose_btc(pl2.soma.extra_grad))
    scale = np.abs(want).max()
    assert scale > 1e-6, scale
    np.testing.assert_allclose(got, want, atol=(0.02 if math == 0 else 0.05) * scale)
    with pytest.raises(AssertionError):       # negative control: a vanished branch gradient is rejected
        np.testing.assert_allclose(0.0 * got, want, atol=0.05 * scale)
    assert g_first is not None


def test_branches_are_exclusive_and_objective_checked(dev, models, branch_models, golden):
    import paule_b200 as P
    from paule_b200 import BatchPlanner
    pred, emb, inv = models
    cp_tube, tube_mel, tube_emb, cls = branch_models
    with pytest.raises(NotImplementedError):
        P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, use_somatosensory_feedback=True,
                use_speech_classifier=True, speech_classifier=cls, cp_tube_model=cp_tube, tube_mel_model=tube_mel,
                tube_embedder=tube_emb)
    cp0, tmel = torch.from_numpy(golden["b3_cp0"]).to(dev), torch.from_numpy(golden["b3_tmel"]).to(dev)
    with pytest.raises(NotImplementedError):
        BatchPlanner(pred, emb, cp0, tmel, None, objective="acoustic", somatosensory=(cp_tube, tube_mel, tube_emb), math=0)


def test_linear_classifier_module_forward(dev, branch_models):
    cls = branch_models[3]
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 11, 60, generator=g)
    w, b = cls.linear.weight.detach().cpu(), cls.linear.bias.detach().cpu()
    want = (x @ w.t() + b).squeeze(2)
    np.testing.assert_allclose(_np(cls(x.to(dev))), want.mean(1).double().numpy(), atol=1e-5)
    lens = [11, 7, 4]
    ref = torch.stack([want[i, :l].sum() / l for i, l in enumerate(lens)])
    np.testing.assert_allclose(_np(cls(x.to(dev), src_lens=lens)), ref.double().numpy(), atol=1e-5)


def test_dropout_tube_embedder_plans_without_graph(dev, models, branch_models, golden):
    """The shipped tube embedder has dropout 0.7 and the reference plans with it in training mode (paule/paule.py:261,928):
    stochastic, so only sanity is checked -- finite, decreasing loss, no CUDA graph."""
    import paule_b200 as P
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    cp_tube, tube_mel, _, _ = branch_models
    torch.manual_seed(7)
    te = P.EmbeddingModel(input_size=10, num_lstm_layers=2, hidden_size=720, dropout=0.7, post_upsampling_size=0).to(dev)
    cp0, tmel = torch.from_numpy(golden["b3_cp0"]).to(dev), torch.from_numpy(golden["b3_tmel"]).to(dev)
    pl = BatchPlanner(pred, emb, cp0, tmel, None, max_log_steps=4, somatosensory=(cp_tube, tube_mel, te), math=0)
    pl.step(4)
    assert pl._graph is None
    tot = _np(pl.losses()["total"])
    assert np.isfinite(tot).all() and (tot[-1] < tot[0]).all()


def test_generator_prologue_modes(dev, models, golden):
    """initialize_from='semvec' (cp GAN generator, paule/paule.py:558-565) and a missing acoustic target (mel GAN generator,
    :515-522): the prologue draws noise [B,1,100] on the device and runs the injected generators; planning then proceeds."""
    import paule_b200 as P
    pred, emb, inv = models
    torch.manual_seed(4)
    cp_gen, mel_gen = P.Generator(output_size=30), P.Generator(output_size=60)
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, cp_gen_model=cp_gen, mel_gen_model=mel_gen, device=dev, math=0)
    tmel = torch.from_numpy(golden["b3_tmel"]).to(dev)                  # [3,20,60]
    B, Tm = tmel.shape[0], tmel.shape[1]
    res = pm.plan_resynth(target_acoustic=tmel, initialize_from="semvec", objective="acoustic_semvec", n_outer=1, n_inner=3,
                          continue_learning=False, verbose=False, seed=123)
    with torch.no_grad():
        tsv = emb(tmel, [Tm] * B)
        torch.manual_seed(123)
        noise = torch.randn(B, 1, 100, device=dev)
        want = pm.cp_gen_model(noise, 2 * Tm, tsv)
    np.testing.assert_allclose(res.initial_cp, _np(want), atol=1e-6)
    assert res.planned_cp.shape == (B, 2 * Tm, 30)
    tot = np.asarray(res.planned_loss_steps)
    assert np.isfinite(tot).all() and (tot[-1] < tot[0]).all()
    # no acoustic target: the mel generator makes one from the semvec
    res2 = pm.plan_resynth(target_acoustic=None, target_semvec=tsv, target_seq_length=Tm, initialize_from="acoustic",
                           objective="acoustic_semvec", n_outer=1, n_inner=2, continue_learning=False, verbose=False,
                           seed=5)
    with torch.no_grad():
        torch.manual_seed(5)
        noise = torch.randn(B, 1, 100, device=dev)
        want_mel = pm.mel_gen_model(noise, Tm, tsv)
    np.testing.assert_allclose(res2.target_mel, _np(want_mel), atol=1e-6)
    assert np.isfinite(np.asarray(res2.planned_loss_steps)).all()


@pytest.mark.parametrize("math", [0, 1])
def test_classifier_with_ragged_batch_equals_words_planned_alone(dev, models, branch_models, math):
    """The fused classifier term averages over each word's OWN mel frames: every word of a ragged batch (padded, with
    non-zero padding content) equals the same word planned alone, classifier term and cps included."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    cls = branch_models[3]
    lengths = [40, 64, 51, 26]
    g = torch.Generator().manual_seed(17)
    T = max(lengths)
    cps = [torch.rand((L, 30), generator=g) - 0.5 for L in lengths]
    mels = [torch.rand((L // 2, 60), generator=g) for L in lengths]
    cp0, tmel = torch.full((len(lengths), T, 30), 0.21), torch.full((len(lengths), T // 2, 60), 3.0)
    for b, (c, m) in enumerate(zip(cps, mels)):
        cp0[b, :c.shape[0]] = c
        tmel[b, :m.shape[0]] = m
    n = 4
    pl = BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=n, math=math, lengths=lengths,
                      speech_classifier=cls)
    pl.step(n)
    cp_b, L_b = _np(pl.planned_cp()), {k: _np(v) for k, v in pl.losses().items()}
    assert np.isfinite(cp_b).all() and np.isfinite(L_b["total"]).all()
    cp_tol, loss_tol = (5e-6, 1e-5) if math == 0 else (3e-5, 2e-4)
    for b, L in enumerate(lengths):
        solo = BatchPlanner(pred, emb, cps[b][None].to(dev), mels[b][None].to(dev), None, max_log_steps=n, math=math,
                            speech_classifier=cls)
        solo.step(n)
        np.testing.assert_allclose(cp_b[b, :L], _np(solo.planned_cp())[0], atol=cp_tol, err_msg=f"word {b} (T={L})")
        for k in ("total", "mel", "semvec", "speech_classifier"):
            np.testing.assert_allclose(L_b[k][:, b], _np(solo.losses()[k])[:, 0], rtol=loss_tol, atol=1e-6, err_msg=f"{k} word {b}")
        solo.close()
