"""Step- and trajectory-level parity of the fused planner against the reference-generated golden vectors
and the CPU oracle, plus the Paule API surface.  Run on the B200 box: python -m pytest tests -m gpu

Tolerances (BASELINE.json north_star): per-step loss 1e-3 relative, final cp 1e-3 absolute in fp32 on the
well-conditioned (iid) input; fp32 math is held to much tighter bounds here (1e-4 / 1e-5).
"""
import numpy as np
import pytest
import torch

from oracle import paule_oracle as O

pytestmark = pytest.mark.gpu

MATHS = [0]   # PAULE_MATH_FP32; tensor-core modes are appended once paule_tc_* is available


def _tc_available():
    from paule_b200 import _lib
    return _lib.load().paule_tc_packed_lstm_bytes(720, 30) > 0


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


def _np(t):
    return t.detach().cpu().double().numpy()


def math_params():
    ps = [pytest.param(0, id="fp32")]
    ps.append(pytest.param(1, id="bf16", marks=pytest.mark.skipif(not _tc_available(), reason="tcgen05 path not built")))
    return ps


# tolerance per math mode: (loss rtol, grad rel-to-max, cp atol)
TOL = {0: (1e-4, 2e-4, 1e-5), 1: (1e-3, 2e-2, 1e-3)}


@pytest.mark.parametrize("math", math_params())
def test_teacher_forced_steps_match_golden(dev, models, golden, math):
    """Feed the oracle's cp of step k, compare loss terms, d(cp) and the post-Adam cp of that single step."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    cp0 = torch.from_numpy(golden["b3_cp0"]).to(dev)
    tmel = torch.from_numpy(golden["b3_tmel"]).to(dev)
    cps = golden["b3_cps"]            # [steps,B,T,30] cp BEFORE each update
    grads = golden["b3_grads"]
    n = cps.shape[0]
    pl = BatchPlanner(pred, emb, cp0, tmel, None, log_gradients=True, max_log_steps=n, math=math,
                      use_cuda_graph=False)
    # the target semvec comes from the planner's own embedder kernels, in the planner's arithmetic (bf16 operands: ~1e-4)
    np.testing.assert_allclose(_np(pl.target_sv), golden["b3_tsv"], atol=5e-6 if math == 0 else 2e-4)
    lr, gr, ca = TOL[math]
    for k in range(n):
        pl.set_cp(torch.from_numpy(cps[k]).to(dev))
        pl.step(1)
        g = _np(pl.last_grad())
        np.testing.assert_allclose(g, grads[k], rtol=gr, atol=gr * np.abs(grads[k]).max())
        nxt = cps[k + 1] if k + 1 < n else golden["b3_planned_cp"]
        np.testing.assert_allclose(_np(pl.planned_cp()), nxt, atol=ca)
    L = pl.losses()
    np.testing.assert_allclose(_np(L["total"]), golden["b3_loss"], rtol=lr)
    np.testing.assert_allclose(_np(torch.stack([L["mel"], L["semvec"], L["velocity"], L["jerk"], L["local_linear"]], -1)),
                               golden["b3_terms"], rtol=lr)


@pytest.mark.parametrize("math", math_params())
@pytest.mark.parametrize("graph", [False, True])
def test_free_running_matches_golden(dev, models, golden, math, graph):
    """Free-running 5 steps on the iid input: loss curve and final cps vs the reference-generated vectors."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    pl = BatchPlanner(pred, emb, torch.from_numpy(golden["b3_cp0"]).to(dev), torch.from_numpy(golden["b3_tmel"]).to(dev),
                      None, max_log_steps=8, math=math, use_cuda_graph=graph)
    pl.step(3)
    pl.step(2)
    lr, _, ca = TOL[math]
    np.testing.assert_allclose(_np(pl.losses()["total"]), golden["b3_loss"], rtol=lr)
    np.testing.assert_allclose(_np(pl.planned_cp()), golden["b3_planned_cp"], atol=ca)
    mel, sv = pl.forward()
    np.testing.assert_allclose(_np(mel), golden["b3_pred_mel"], atol=max(ca, 2e-5))
    np.testing.assert_allclose(_np(sv), golden["b3_pred_semvec"], atol=max(ca, 2e-5))


@pytest.mark.parametrize("math", math_params() + [pytest.param(None, id="default")])
@pytest.mark.parametrize("tag,objective,smiling", [("real64", "acoustic_semvec", False), ("real64_ac", "acoustic", False),
                                                   ("real64_sv", "semvec", True), ("real32", "acoustic_semvec", False)])
def test_paule_plan_resynth_matches_the_real_reference(dev, models, golden, tag, objective, smiling, math):
    """Paule.plan_resynth (our API) vs the REAL reference's plan_resynth outputs (fp64 / fp32 CPU), B=1 -- in fp32, in the
    bf16 tensor-core mode, and with the default arithmetic (the tensor-core path for Paule's 720-unit models)."""
    import paule_b200 as P
    from paule_b200 import ops
    pred, emb, inv = models
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, smiling=smiling, math=math)
    n = len(golden[f"{tag}_loss"])
    res = pm.plan_resynth(target_acoustic=golden[f"{tag}_tmel"][0].astype(np.float32),
                          initial_cp=golden[f"{tag}_cp0"][0].astype(np.float32), initialize_from=None,
                          objective=objective, n_outer=1, n_inner=n, log_ii=1, continue_learning=False,
                          verbose=False, log_semantics=False, log_cps=True)
    used = pm.last_planner.math
    assert used == (ops.default_math(720) if math is None else math)
    if math is None and _tc_available():
        assert used == ops.MATH_BF16, "the reference-facing API must take the tensor-core path by default"
    lr, _, ca = TOL[used]
    ma = 2e-5 if used == 0 else 2e-3          # predictions: bf16 operands
    assert len(res) == 33 and res._fields[0] == "planned_cp" and res._fields[-1] == "inv_model_loss"
    np.testing.assert_allclose(res.planned_loss_steps, golden[f"{tag}_loss"], rtol=lr)
    np.testing.assert_allclose(res.vel_loss_steps, golden[f"{tag}_vel"], rtol=lr)
    np.testing.assert_allclose(res.jerk_loss_steps, golden[f"{tag}_jerk"], rtol=lr)
    np.testing.assert_allclose(res.planned_mel_loss_steps, golden[f"{tag}_mel"], rtol=lr)
    if objective != "acoustic":
        np.testing.assert_allclose(res.pred_semvec_loss_steps, golden[f"{tag}_sem"], rtol=lr)
    assert res.planned_cp.shape == golden[f"{tag}_planned_cp"].shape
    np.testing.assert_allclose(res.planned_cp, golden[f"{tag}_planned_cp"], atol=ca)
    if tag.startswith("real64"):
        np.testing.assert_allclose(np.stack(res.cp_steps[0]), golden[f"{tag}_cp_steps"], atol=ca)
        np.testing.assert_allclose(res.pred_mel, golden[f"{tag}_pred_mel"], atol=ma)
        np.testing.assert_allclose(res.pred_semvec, golden[f"{tag}_pred_semvec"], atol=ma)
    if smiling:
        assert np.all(res.planned_cp[:, 4] == -1.0) and np.all(res.planned_cp[:, 1] == 1.0)


def test_plan_resynth_in_a_loop_reuses_one_planner(dev, models, golden):
    """plan_resynth is normally called in a loop over words (reference notebook, cell 28): the second call of the same
    shape re-arms the first call's planner (workspace, packed weights, CUDA graph) instead of allocating a new one, a call
    of another shape releases it, and a re-armed planner gives the results of a fresh one."""
    import paule_b200 as P
    pred, emb, inv = models
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev)
    kw = dict(initialize_from=None, objective="acoustic_semvec", n_outer=1, n_inner=4, continue_learning=False, verbose=False)
    cpa, mela = O.synthetic_inputs(2, 40, seed=61)
    cpb, melb = O.synthetic_inputs(2, 40, seed=62)
    ra = pm.plan_resynth(target_acoustic=mela.numpy(), initial_cp=cpa.numpy(), **kw)
    first = pm.last_planner
    ws_ptr = first.workspace.data_ptr()
    rb = pm.plan_resynth(target_acoustic=melb.numpy(), initial_cp=cpb.numpy(), **kw)
    assert pm.last_planner is first and first.workspace.data_ptr() == ws_ptr and first._graph is not None
    fresh = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev).plan_resynth(
        target_acoustic=melb.numpy(), initial_cp=cpb.numpy(), **kw)
    np.testing.assert_allclose(np.stack(rb.planned_loss_steps), np.stack(fresh.planned_loss_steps), rtol=2e-4)
    np.testing.assert_allclose(rb.planned_cp, fresh.planned_cp, atol=2e-5)
    assert not np.allclose(ra.planned_cp, rb.planned_cp)
    cpc, melc = O.synthetic_inputs(2, 60, seed=63)
    pm.plan_resynth(target_acoustic=melc.numpy(), initial_cp=cpc.numpy(), **kw)
    assert pm.last_planner is not first and first.workspace is None            # released, not leaked
    # planners that go out of scope leave the registry (it holds weak references)
    from paule_b200 import ops
    import gc
    gc.collect()
    n0 = len(ops._PLAN_REGISTRY)
    tmp = P.BatchPlanner(pred, emb, cpa.to(dev), mela.to(dev), None, max_log_steps=2)
    assert len(ops._PLAN_REGISTRY) == n0 + 1
    del tmp
    gc.collect()
    assert len(ops._PLAN_REGISTRY) == n0


def test_smooth_input_teacher_forced(dev, models, golden):
    """Chaotic regime (smooth cps): only step-0 quantities are pinned tightly."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    pl = BatchPlanner(pred, emb, torch.from_numpy(golden["real32_smooth_cp0"]).to(dev),
                      torch.from_numpy(golden["real32_smooth_tmel"]).to(dev), None, max_log_steps=8, math=0)
    pl.step(6)
    L = _np(pl.losses()["total"])[:, 0]
    np.testing.assert_allclose(L[:2], golden["real32_smooth_loss"][:2], rtol=1e-4)
    np.testing.assert_allclose(L, golden["real32_smooth_loss"], rtol=5e-2)


def test_batched_equals_solo_and_past_cp(dev, models):
    """B words batched == B solo runs (words are independent); past_cp prefix is frozen."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    cp0, tmel = O.synthetic_inputs(4, 40, seed=21)
    cp0, tmel = cp0.to(dev), tmel.to(dev)
    pl = BatchPlanner(pred, emb, cp0, tmel, None, max_log_steps=4, math=0)
    pl.step(4)
    full = pl.planned_cp()
    for i in (0, 3):
        solo = BatchPlanner(pred, emb, cp0[i:i + 1], tmel[i:i + 1], None, max_log_steps=4, math=0)
        solo.step(4)
        np.testing.assert_allclose(_np(solo.planned_cp()), _np(full[i:i + 1]), atol=2e-6)
    past = cp0[:, :6].clone()
    pl2 = BatchPlanner(pred, emb, cp0, tmel, None, max_log_steps=4, past_cp=past, math=0)
    pl2.step(3)
    assert torch.equal(pl2.planned_cp()[:, :6], past)
    # against the oracle
    pr, em, _ = O.build_reference_models(0, 720, torch.float32, with_inverse=False)
    r = O.plan_inner_loop(pr, em, cp0.cpu(), tmel.cpu(), 3, past_cp=past.cpu())
    np.testing.assert_allclose(_np(pl2.planned_cp()), r["planned_cp"].double().numpy(), atol=1e-5)


def test_value_errors_of_the_reference_api(dev, models):
    """The 8 ValueErrors of the reference's tests/test_paule.py:31-62 (with a mel array as acoustic target)."""
    import paule_b200 as P
    pred, emb, inv = models
    pm = P.PAULE(pred_model=pred, inv_model=inv, embedder=emb, device=dev)
    mel = np.random.RandomState(0).rand(20, 60).astype(np.float32)
    cp_11zeros = np.zeros((11, 30), dtype=np.float32)
    with pytest.raises(ValueError):
        pm.plan_resynth(target_acoustic=None, target_semvec=None)
    with pytest.raises(ValueError):
        pm.plan_resynth(target_acoustic=mel, target_semvec=None, n_inner=5, log_ii=10)
    with pytest.raises(ValueError):
        pm.plan_resynth(target_acoustic=None, target_semvec=np.zeros((300,)))
    with pytest.raises(ValueError):
        pm.plan_resynth(target_acoustic=mel, initialize_from='ERROR')
    with pytest.raises(ValueError):
        pm.plan_resynth(target_acoustic=mel, initial_cp=cp_11zeros, initialize_from='ERROR')
    with pytest.raises(ValueError):
        pm.plan_resynth(target_acoustic=mel, initial_cp=cp_11zeros)
    with pytest.raises(ValueError):
        pm.plan_resynth(target_acoustic=mel, past_cp=cp_11zeros)
    with pytest.raises(ValueError):
        pm.plan_resynth(target_acoustic=mel, objective="ERROR")


def test_inverse_init_batched_plan(dev, models):
    """BASELINE config 3 in miniature: inverse-model initialisation + planning, batched, results with a word axis."""
    import paule_b200 as P
    pred, emb, inv = models
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, math=0)
    _, tmel = O.synthetic_inputs(3, 40, seed=31)
    res = pm.plan_resynth(target_acoustic=tmel.numpy(), initialize_from="acoustic", objective="acoustic_semvec",
                          n_outer=2, n_inner=3, continue_learning=False, verbose=False)
    assert res.planned_cp.shape == (3, 40, 30) and res.initial_cp.shape == (3, 40, 30)
    assert np.abs(res.initial_cp).max() <= 1.0 and np.abs(res.planned_cp).max() <= 1.05
    assert len(res.planned_loss_steps) == 6 and res.planned_loss_steps[0].shape == (3,)
    # oracle: same init (its fp32-capable inverse model), same loop
    pr, em, iv = O.build_reference_models(0, 720, torch.float32)
    with torch.no_grad():
        init = iv(tmel).clamp(-1, 1)
    np.testing.assert_allclose(res.initial_cp, init.numpy(), atol=5e-5)
    r = O.plan_inner_loop(pr, em, torch.from_numpy(res.initial_cp), tmel, 6)
    np.testing.assert_allclose(np.stack(res.planned_loss_steps), r["loss"].numpy(), rtol=1e-3)


def _full_size_check(dev, pred, emb, cp0, tmel, n_steps, math, probe):
    """Size-independent properties at a BASELINE shape instead of an element-wise oracle run over the whole batch:
    the loss falls monotonically on the iid input and stays finite, the cps stay clamped, word `probe` of the batch
    equals the same word planned alone, and that word matches the CPU oracle within the BASELINE tolerance."""
    from paule_b200 import BatchPlanner
    cp0, tmel = cp0.to(dev), tmel.to(dev)
    pl = BatchPlanner(pred, emb, cp0, tmel, None, max_log_steps=n_steps, math=math)
    pl.step(n_steps)
    L = pl.losses()["total"]
    assert torch.isfinite(L).all()
    assert (L[1:] < L[:-1]).all(), "loss must fall monotonically on the well-conditioned input (SURVEY appendix B)"
    assert pl.planned_cp().abs().max().item() <= 1.05 + 1e-6
    batch_cp = pl.planned_cp()[probe:probe + 1].clone()
    batch_loss = L[:, probe:probe + 1].clone()
    pl.close()
    del pl
    solo = BatchPlanner(pred, emb, cp0[probe:probe + 1], tmel[probe:probe + 1], None, max_log_steps=n_steps, math=math)
    solo.step(n_steps)
    np.testing.assert_allclose(_np(solo.planned_cp()), _np(batch_cp), atol=5e-6 if math == 0 else 2e-5)
    np.testing.assert_allclose(_np(solo.losses()["total"]), _np(batch_loss), rtol=1e-5 if math == 0 else 1e-4)
    pr, em, _ = O.build_reference_models(0, 720, torch.float32, with_inverse=False)
    r = O.plan_inner_loop(pr, em, cp0[probe:probe + 1].cpu(), tmel[probe:probe + 1].cpu(), n_steps)
    np.testing.assert_allclose(_np(solo.losses()["total"]), r["loss"].double().numpy(), rtol=1e-3)
    np.testing.assert_allclose(_np(solo.planned_cp()), r["planned_cp"].double().numpy(), atol=1e-3)


@pytest.mark.parametrize("math", math_params())
def test_full_size_properties_cfg2(dev, models, math):
    """BASELINE config 2 exactly: B=64 words, 0.5 s utterances (T=200), 50 inner steps; word 7 against the oracle over all 50."""
    pred, emb, _ = models
    cp0, tmel = O.synthetic_inputs(64, 200, seed=5)
    _full_size_check(dev, pred, emb, cp0, tmel, 50, math, probe=7)


@pytest.mark.skipif(not _tc_available(), reason="tcgen05 path not built")
def test_full_size_properties_cfg3_inverse_init(dev, models):
    """BASELINE config 3: InverseModel initialisation + planning, B=256, 1 s utterances (T=400), through Paule.plan_resynth.
    The smooth inverse-model init is the chaotic regime: the initial cps are pinned against the oracle's
    inverse model, the planning itself through batch == solo and finiteness."""
    import paule_b200 as P
    pred, emb, inv = models
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, math=1)
    _, tmel = O.synthetic_inputs(256, 400, seed=41)
    res = pm.plan_resynth(target_acoustic=tmel.numpy(), initialize_from="acoustic", objective="acoustic_semvec",
                          n_outer=1, n_inner=3, continue_learning=False, verbose=False)
    assert res.planned_cp.shape == (256, 400, 30) and np.isfinite(res.planned_cp).all()
    assert np.abs(res.initial_cp).max() <= 1.0 and np.abs(res.planned_cp).max() <= 1.05 + 1e-6
    _, _, iv = O.build_reference_models(0, 720, torch.float32)
    with torch.no_grad():
        init = iv(tmel[100:103]).clamp(-1, 1)
    # the inverse model's recurrence runs on the bf16 tensor-core kernel in this mode (fp32 mode: 5e-5, test_inverse_init_batched_plan)
    np.testing.assert_allclose(res.initial_cp[100:103], init.numpy(), atol=5e-3)
    solo = pm.plan_resynth(target_acoustic=tmel[101].numpy(), initial_cp=res.initial_cp[101], initialize_from=None,
                           objective="acoustic_semvec", n_outer=1, n_inner=3, continue_learning=False, verbose=False)
    np.testing.assert_allclose(np.array(solo.planned_loss_steps), np.stack(res.planned_loss_steps)[:, 101], rtol=2e-3)


@pytest.mark.skipif(not _tc_available(), reason="tcgen05 path not built")
def test_full_size_properties_cfg4_shard(dev, models):
    """BASELINE config 4 as one rank of 8 sees it: 256 of the 2048 words, 1 s utterances (T=400)."""
    pred, emb, _ = models
    cp0, tmel = O.synthetic_inputs(256, 400, seed=43)
    _full_size_check(dev, pred, emb, cp0, tmel, 4, 1, probe=201)


@pytest.mark.skipif(not _tc_available(), reason="tcgen05 path not built")
def test_full_size_properties_cfg5_long_utterance_shard(dev, models):
    """BASELINE config 5 as one rank of 8 sees it: 64 of the 512 words, 3 s utterances (T=1200, deep BPTT)."""
    pred, emb, _ = models
    cp0, tmel = O.synthetic_inputs(64, 1200, seed=47)
    _full_size_check(dev, pred, emb, cp0, tmel, 3, 1, probe=33)


# ---- ragged batches: words of different lengths padded to the longest one --------------------------
RAGGED = [40, 64, 51, 80, 46, 80, 13]   # cp frames per word (one odd, one at the 13-frame minimum of the jerk stencil)


def _ragged_inputs(seed):
    g = torch.Generator().manual_seed(seed)
    T = max(RAGGED)
    cps = [torch.rand((L, 30), generator=g) - 0.5 for L in RAGGED]
    mels = [torch.rand((L // 2, 60), generator=g) for L in RAGGED]
    cp0, tmel = torch.zeros(len(RAGGED), T, 30), torch.zeros(len(RAGGED), T // 2, 60)
    for b, (c, m) in enumerate(zip(cps, mels)):
        cp0[b, :c.shape[0]] = c
        tmel[b, :m.shape[0]] = m
    # padding frames are NOT zeros: whatever is there must not leak into a word
    cp0[0, RAGGED[0]:] = 0.37
    tmel[0, RAGGED[0] // 2:] = 5.0
    return cps, mels, cp0, tmel


@pytest.mark.parametrize("math", math_params())
def test_ragged_batch_equals_words_planned_alone(dev, models, math):
    """Every word of a ragged batch == the same word planned alone (its own length, no padding), and the padding frames of
    the cps never move; word 2 (odd length) is also held against the CPU oracle at the BASELINE tolerance."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    cps, mels, cp0, tmel = _ragged_inputs(11)
    n = 6
    pl = BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=n, math=math, lengths=RAGGED)
    pl.step(n)
    cp_b, L_b = _np(pl.planned_cp()), {k: _np(v) for k, v in pl.losses().items()}
    mel_b, sv_b = pl.forward()
    assert np.isfinite(cp_b).all() and np.isfinite(L_b["total"]).all()
    np.testing.assert_array_equal(cp_b[0, RAGGED[0]:], np.float32(0.37))        # padding untouched
    cp_tol, loss_tol = (5e-6, 1e-5) if math == 0 else (3e-5, 2e-4)
    for b, L in enumerate(RAGGED):
        solo = BatchPlanner(pred, emb, cps[b][None].to(dev), mels[b][None].to(dev), None, max_log_steps=n, math=math)
        solo.step(n)
        np.testing.assert_allclose(cp_b[b, :L], _np(solo.planned_cp())[0], atol=cp_tol, err_msg=f"word {b} (T={L})")
        for k in ("total", "mel", "semvec", "velocity", "jerk", "local_linear"):
            np.testing.assert_allclose(L_b[k][:, b], _np(solo.losses()[k])[:, 0], rtol=loss_tol, atol=1e-6, err_msg=f"{k} word {b}")
        m1, s1 = solo.forward()
        np.testing.assert_allclose(_np(mel_b)[b, :L // 2], _np(m1)[0], atol=1e-5 if math == 0 else 2e-3)
        np.testing.assert_allclose(_np(sv_b)[b], _np(s1)[0], atol=1e-5 if math == 0 else 2e-3)
        solo.close()
    pr, em, _ = O.build_reference_models(0, 720, torch.float32, with_inverse=False)
    r = O.plan_inner_loop(pr, em, cps[2][None], mels[2][None], n)
    np.testing.assert_allclose(L_b["total"][:, 2], r["loss"].double().numpy()[:, 0], rtol=1e-3)
    np.testing.assert_allclose(cp_b[2, :RAGGED[2]], r["planned_cp"].double().numpy()[0], atol=1e-3)


def test_paule_plan_resynth_ragged_list(dev, models):
    """Paule.plan_resynth with a list of per-word mel arrays: per-word results without padding, each equal to the word
    planned through the single-word API; the inverse-model initialisation runs per word (it is not causal)."""
    import paule_b200 as P
    pred, emb, inv = models
    pm = P.Paule(pred_model=pred, inv_model=inv, embedder=emb, device=dev, math=0)
    g = torch.Generator().manual_seed(3)
    mels = [torch.rand((n, 60), generator=g).numpy() for n in (20, 32, 25)]
    res = pm.plan_resynth(target_acoustic=mels, initialize_from="acoustic", objective="acoustic_semvec", n_outer=1, n_inner=4,
                          continue_learning=False, verbose=False)
    assert [c.shape for c in res.planned_cp] == [(40, 30), (64, 30), (50, 30)]
    assert [m.shape for m in res.pred_mel] == [(20, 60), (32, 60), (25, 60)]
    for b, m in enumerate(mels):
        solo = pm.plan_resynth(target_acoustic=m, initialize_from="acoustic", objective="acoustic_semvec", n_outer=1, n_inner=4,
                               continue_learning=False, verbose=False)
        np.testing.assert_allclose(res.initial_cp[b], solo.initial_cp, atol=1e-6)
        np.testing.assert_allclose(np.stack(res.planned_loss_steps)[:, b], np.array(solo.planned_loss_steps), rtol=1e-4)
        np.testing.assert_allclose(res.planned_cp[b], solo.planned_cp, atol=1e-5)
    with pytest.raises(ValueError):
        P.BatchPlanner(pred, emb, torch.zeros(2, 40, 30, device=dev), torch.zeros(2, 20, 60, device=dev), None, lengths=[40, 12])


def test_watchdog_status_is_checked_when_results_are_read(dev, models):
    """The persistent kernels never hang the GPU: a stuck inter-CTA wait sets a sticky status word instead.  BatchPlanner reads
    it whenever results leave the planner and raises -- here the word is poked by hand."""
    from paule_b200 import BatchPlanner, _lib
    if not _tc_available():
        pytest.skip("tcgen05 path not built")
    pred, emb, _ = models
    cp0, tmel = O.synthetic_inputs(2, 40, seed=1)
    pl = BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=4, math=1)
    pl.step(2)
    pl.planned_cp(), pl.losses()          # healthy
    pl.workspace[pl._status_off:pl._status_off + 4].view(torch.int32).fill_(2)
    with pytest.raises(_lib.PauleB200Error):
        pl.planned_cp()
    with pytest.raises(_lib.PauleB200Error):
        pl.losses()


@pytest.mark.skipif(not _tc_available(), reason="tcgen05 path not built")
@pytest.mark.parametrize("B,T,steps", [(64, 200, 300), (100, 60, 150), (1, 200, 200), (200, 40, 100), (450, 20, 60)])
def test_soak_many_steps_no_watchdog_and_repeatable(dev, models, B, T, steps):
    """Soak of the exchange protocol of the persistent kernels (latency layout, 2- and 3-quarter CTAs, a batch beyond one launch
    cut into passes of different layouts) and of the CTA-pair GEMMs between them: hundreds
    of steps = 10^5 .. 10^6 inter-CTA exchanges per run; the watchdog word must stay clear, the loss finite and falling, and
    two runs of the same job must agree up to the bf16 rounding flips of the arrival-ordered accumulation."""
    from paule_b200 import BatchPlanner
    pred, emb, _ = models
    cp0, tmel = O.synthetic_inputs(B, T, seed=123)
    outs = []
    for _ in range(2):
        pl = BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=steps, math=1)
        pl.step(steps)
        pl.check()
        L = pl.losses()["total"]
        assert torch.isfinite(L).all()
        assert (L[-1] < L[0]).all() and (L[10:].mean(1)[1:] <= L[10:].mean(1)[:-1] * 1.001).all()
        outs.append((pl.planned_cp().clone(), L.clone()))
        pl.close()
    assert (outs[0][0] - outs[1][0]).abs().max().item() < 2e-3
    np.testing.assert_allclose(_np(outs[0][1][-1]), _np(outs[1][1][-1]), rtol=1e-3)


@pytest.mark.skipif(not _tc_available(), reason="tcgen05 path not built")
def test_narrow_models_plan_on_the_tensor_core_path_zero_padded(dev):
    """Models narrower than the 720 units the tcgen05 kernels are built for (here 360: the width of the reference's
    somatosensory models, paule/paule.py:231-249) plan in tensor-core mode with every layer zero-padded to 720 units, which is
    exact: losses / cps against the CPU oracle on the SAME 360-unit models, model-path gradient against fp64 autograd."""
    import paule_b200 as P
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=360).to(dev)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=360).to(dev)
    p32, e32, _ = O.build_reference_models(0, 360, torch.float32, with_inverse=False)
    assert O.state_dict_digest(p32) == O.state_dict_digest(pred.cpu()) and O.state_dict_digest(e32) == O.state_dict_digest(emb.cpu())
    pred, emb = pred.to(dev), emb.to(dev)
    cp0, tmel = O.synthetic_inputs(3, 40, seed=11)
    pl = P.BatchPlanner(pred, emb, cp0.to(dev), tmel.to(dev), None, max_log_steps=4, math=1)
    assert pl.H == 720 and pl._pad
    pl.step(1)
    p64, e64, _ = O.build_reference_models(0, 360, torch.float64, with_inverse=False)
    want = O.model_path_grad(p64, e64, cp0.double(), tmel.double()).numpy()
    got = _np(pl.last_grad_lstm())
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-2 * np.abs(want).max())
    with pytest.raises(AssertionError):
        np.testing.assert_allclose(0 * got, want, rtol=0, atol=2e-2 * np.abs(want).max())
    pl.step(3)
    r = O.plan_inner_loop(p32, e32, cp0, tmel, 4)
    np.testing.assert_allclose(_np(pl.losses()["total"]), r["loss"].numpy(), rtol=1e-3)
    np.testing.assert_allclose(_np(pl.planned_cp()), r["planned_cp"].numpy(), atol=1e-3)
