"""Parity of the MODEL-PATH gradient of the fused inner step (discrepancy.backward(), paule/paule.py:1052).

``paule_plan_step`` leaves d(mel + semvec terms)/d(cp) -- semvec -> head^T -> BPTT(embedder l1) -> dX -> BPTT(l0) -> dX (+ mel-loss
gradient) -> post_linear^T -> un-pool (x0.5) -> BPTT(forward model) -> dX -- in its own buffer (``BatchPlanner.last_grad_lstm``)
before the Adam kernel adds the smoothness gradients.  On the synthetic inputs that part is 10^4..10^6 times SMALLER than the
velocity / jerk / local-linear gradient, so it is compared here on its own, with a tolerance relative to ITS OWN maximum:

    fp32 kernels  <= 2e-3 x max|g_model|        bf16 tensor-core path (the benchmarked mode)  <= 2e-2 x max|g_model|

against (i) the REAL reference's ``xx_new.grad`` in fp64 minus the analytic smoothness gradient
(tests/golden/make_grad_golden.py, B = 1) and (ii) fp64 autograd of the oracle (pinned to (i) at 2e-9 by tests/test_oracle.py)
for batches, ragged batches and BASELINE configs[1]'s full shape.  Every comparison carries NEGATIVE CONTROLS: the same
assertion must reject a zeroed gradient, a gradient with the 0.5 un-pool factor dropped (= 2 x, every path goes through the
un-pooling once) and, for the combined objective, a gradient that lacks the embedder path.
"""
import numpy as np
import pytest
import torch

from oracle import paule_oracle as O

pytestmark = pytest.mark.gpu

REL = {0: 2e-3, 1: 2e-2}      # tolerance / max|model-path gradient| per math mode
OBJECTIVES = ["acoustic_semvec", "acoustic", "semvec"]


def _tc_available():
    from paule_b200 import _lib
    return _lib.load().paule_tc_packed_lstm_bytes(720, 30) > 0


def math_params():
    return [pytest.param(0, id="fp32"),
            pytest.param(1, id="bf16", marks=pytest.mark.skipif(not _tc_available(), reason="tcgen05 path not built"))]


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
    assert [O.state_dict_digest(m) for m in (pred, emb)] == list(golden["digest32"])[:2]
    return pred.to(dev), emb.to(dev)


@pytest.fixture(scope="module")
def models64():
    pred, emb, _ = O.build_reference_models(0, 720, torch.float64, with_inverse=False)
    return pred, emb


def _np(t):
    return t.detach().cpu().double().numpy()


def _close(got, want, rel):
    scale = np.abs(want).max()
    assert scale > 0
    np.testing.assert_allclose(got, want, rtol=0, atol=rel * scale)
    return float(np.abs(got - want).max() / scale)


def assert_model_gradient(got, want, rel, without_embedder=None):
    """got ~ want within rel x max|want| -- and the same assertion rejects the broken gradients it has to catch."""
    err = _close(got, want, rel)
    with pytest.raises(AssertionError):
        _close(np.zeros_like(got), want, rel)            # dcp_lstm zeroed
    with pytest.raises(AssertionError):
        _close(2.0 * got, want, rel)                     # 0.5 un-pool factor dropped
    if without_embedder is not None:                     # semvec -> embedder -> dmel path missing
        with pytest.raises(AssertionError):
            _close(without_embedder, want, rel)
    return err


@pytest.mark.parametrize("math", math_params())
@pytest.mark.parametrize("init", ["iid", "smooth"])
@pytest.mark.parametrize("objective", OBJECTIVES)
def test_model_gradient_teacher_forced_vs_the_real_reference(dev, models, golden_grad, math, init, objective):
    """B = 1, T = 40: the cps of every step of the real reference's run are fed in, the step's model-path gradient must match
    the reference's (xx_new.grad - smoothness gradient, fp64)."""
    from paule_b200 import BatchPlanner
    pred, emb = models
    g = golden_grad
    tag = f"{init}_{objective}"
    cps, want = g[f"{tag}_cps"], g[f"{tag}_grad_model"]
    tmel = torch.from_numpy(g[f"{init}_tmel"]).float().to(dev)
    n = cps.shape[0]
    pl = BatchPlanner(pred, emb, torch.from_numpy(cps[0][None]).float().to(dev), tmel, None, objective=objective,
                      log_gradients=True, max_log_steps=n, math=math, use_cuda_graph=False)
    for k in range(n):
        pl.set_cp(torch.from_numpy(cps[k][None]).float().to(dev))
        pl.step(1)
        lacking = g[f"{init}_acoustic_grad_model"][k][None] if objective == "acoustic_semvec" else None
        assert_model_gradient(_np(pl.last_grad_lstm()), want[k][None], REL[math], lacking)
        # the total gradient (what Adam consumes) = model part + smoothness part; the fp32 local-linear stencil carries a
        # cancellation floor of ~eps x |cp| x 2e5 / N = 2e-5 absolute -- which is why the model part is
        # pinned on its own above
        total = _np(pl.last_grad())
        np.testing.assert_allclose(total, g[f"{tag}_grad"][k][None], rtol=2e-4, atol=REL[math] * np.abs(want[k]).max() + 5e-5)
    np.testing.assert_allclose(_np(pl.losses()["total"])[:, 0], g[f"{tag}_loss"], rtol=1e-4 if math == 0 else 1e-3)
    pl.close()


@pytest.mark.parametrize("math", math_params())
@pytest.mark.parametrize("objective", OBJECTIVES)
@pytest.mark.parametrize("ragged", [False, True], ids=["full", "ragged"])
def test_model_gradient_batched_and_ragged(dev, models, models64, math, objective, ragged):
    """B = 3 words (T = 40), also as a ragged batch (40 / 26 / 33 frames): every word's model-path gradient vs fp64 autograd of
    the oracle on that word alone; padding frames receive exactly zero."""
    from paule_b200 import BatchPlanner
    pred, emb = models
    p64, e64 = models64
    cp0, tmel = O.synthetic_inputs(3, 40, seed=11, dtype=torch.float64)
    lens = [40, 26, 33] if ragged else None
    pl = BatchPlanner(pred, emb, cp0.float().to(dev), tmel.float().to(dev), None, objective=objective, max_log_steps=2,
                      math=math, use_cuda_graph=False, lengths=lens)
    pl.step(1)
    got = _np(pl.last_grad_lstm())
    want = O.model_path_grad(p64, e64, cp0, tmel, objective=objective, lens=lens).numpy()
    lack = O.model_path_grad(p64, e64, cp0, tmel, objective="acoustic", lens=lens).numpy() if objective == "acoustic_semvec" else None
    for b in range(3):
        assert_model_gradient(got[b], want[b], REL[math], None if lack is None else lack[b])
    if ragged:
        assert np.all(got[1, 26:] == 0) and np.all(got[2, 33:] == 0)
    pl.close()


@pytest.mark.parametrize("math", math_params())
def test_model_gradient_at_configs1_shape(dev, models, models64, math):
    """BASELINE configs[1] exactly (64 words x 200 cp frames, acoustic_semvec) in the benchmarked layout (CUDA graph, 64-word
    latency layout of the persistent kernels): four probe words against fp64 autograd of the oracle."""
    from paule_b200 import BatchPlanner
    pred, emb = models
    p64, e64 = models64
    cp0, tmel = O.synthetic_inputs(64, 200, seed=5, dtype=torch.float64)
    pl = BatchPlanner(pred, emb, cp0.float().to(dev), tmel.float().to(dev), None, max_log_steps=4, math=math)
    pl.step(1)
    got = _np(pl.last_grad_lstm())
    for b in (0, 7, 33, 63):
        want = O.model_path_grad(p64, e64, cp0[b:b + 1], tmel[b:b + 1])[0].numpy()
        lack = O.model_path_grad(p64, e64, cp0[b:b + 1], tmel[b:b + 1], objective="acoustic")[0].numpy()
        assert_model_gradient(got[b], want, REL[math], lack)
    pl.close()


@pytest.mark.parametrize("math", math_params())
def test_total_gradient_on_ramp_cps(dev, models, models64, math):
    """Ramp cps (every channel a straight line in time): local-linear and jerk terms vanish, the velocity gradient lives on
    the four frames at either end, so the TOTAL gradient Adam consumes is the model-path gradient almost everywhere -- the
    regime in which the planned cps themselves depend on the LSTM BPTT (on iid cps Adam just follows the smoothness sign)."""
    from paule_b200 import BatchPlanner
    pred, emb = models
    p64, e64 = models64
    g = torch.Generator().manual_seed(17)
    B, T = 2, 40
    # offsets / slopes on a 2^-10 / 2^-12 grid: every cp value and every stencil difference is exact in fp32, so the
    # smoothness gradient carries no cancellation noise and the comparison isolates the model path
    a = torch.round((torch.rand(B, 1, 30, generator=g, dtype=torch.float64) - 0.5) * 1024) / 1024
    s = torch.round((torch.rand(B, 1, 30, generator=g, dtype=torch.float64) - 0.5) * 0.02 * 4096) / 4096
    cp0 = a + s * torch.arange(T, dtype=torch.float64)[None, :, None]
    assert torch.equal(cp0.float().double(), cp0)
    tmel = torch.rand(B, T // 2, 60, generator=g, dtype=torch.float64)
    pl = BatchPlanner(pred, emb, cp0.float().to(dev), tmel.float().to(dev), None, log_gradients=True, max_log_steps=2,
                      math=math, use_cuda_graph=False)
    pl.step(1)
    model = O.model_path_grad(p64, e64, cp0, tmel).numpy()
    total = model + O.manual_smooth_grad(cp0).numpy()
    inner = slice(8, T - 8)
    assert np.abs(total[:, inner] - model[:, inner]).max() < 0.05 * np.abs(model).max()   # the regime claimed above
    assert_model_gradient(_np(pl.last_grad_lstm()), model, REL[math])
    got_total = _np(pl.last_grad())
    scale = np.abs(model).max()
    np.testing.assert_allclose(got_total[:, inner], total[:, inner], rtol=0, atol=REL[math] * scale + 1e-3 * scale)
    pl.close()
