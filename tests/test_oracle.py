"""The oracle against vectors produced by the reference itself (tests/golden/make_golden.py).

CPU only.  Pins oracle/paule_oracle.py: module restatements, the batched inner loop, and the
independent manual restatement (explicit LSTM + hand BPTT + hand Adam) whose formulas the CUDA
kernels implement.
"""
import numpy as np
import pytest
import torch

from oracle import paule_oracle as O


@pytest.fixture(scope="module")
def models32():
    return O.build_reference_models(0, 720, torch.float32)


@pytest.fixture(scope="module")
def models64():
    return O.build_reference_models(0, 720, torch.float64)


def test_weights_regenerate_bit_identically(golden, models32, models64):
    for mods, key in ((models32, "digest32"), (models64, "digest64")):
        got = [O.state_dict_digest(m) for m in mods]
        assert got == list(golden[key]), "seeded random-init weights differ from the reference's"


@pytest.mark.parametrize("tag,objective,smiling", [("real64", "acoustic_semvec", False),
                                                   ("real64_ac", "acoustic", False),
                                                   ("real64_sv", "semvec", True)])
def test_inner_loop_matches_real_plan_resynth_fp64(golden, models64, tag, objective, smiling):
    pred, emb, _ = models64
    cp0 = torch.from_numpy(golden[f"{tag}_cp0"])
    tmel = torch.from_numpy(golden[f"{tag}_tmel"])
    n = len(golden[f"{tag}_loss"])
    r = O.plan_inner_loop(pred, emb, cp0, tmel, n, objective=objective, smiling=smiling, log_cps=True)
    np.testing.assert_allclose(r["loss"][:, 0].numpy(), golden[f"{tag}_loss"], rtol=1e-12)
    np.testing.assert_allclose(r["terms"][:, 0, 2].numpy(), golden[f"{tag}_vel"], rtol=1e-12)
    np.testing.assert_allclose(r["terms"][:, 0, 3].numpy(), golden[f"{tag}_jerk"], rtol=1e-12)
    np.testing.assert_allclose(r["terms"][:, 0, 0].numpy(), golden[f"{tag}_mel"], rtol=1e-12)
    if objective != "acoustic":
        np.testing.assert_allclose(r["terms"][:, 0, 1].numpy(), golden[f"{tag}_sem"], rtol=1e-12)
    np.testing.assert_allclose(r["planned_cp"][0].numpy(), golden[f"{tag}_planned_cp"], atol=1e-14)
    np.testing.assert_allclose(torch.stack(r["cps"])[:, 0].numpy(), golden[f"{tag}_cp_steps"], atol=1e-14)
    np.testing.assert_allclose(r["pred_mel"][0].numpy(), golden[f"{tag}_pred_mel"], atol=1e-13)
    np.testing.assert_allclose(r["pred_semvec"][0].numpy(), golden[f"{tag}_pred_semvec"], atol=1e-13)


@pytest.mark.parametrize("tag", ["real32", "real32_smooth"])
def test_inner_loop_matches_real_plan_resynth_fp32(golden, models32, tag):
    torch.set_num_threads(1)
    pred, emb, _ = models32
    cp0 = torch.from_numpy(golden[f"{tag}_cp0"])
    tmel = torch.from_numpy(golden[f"{tag}_tmel"])
    n = len(golden[f"{tag}_loss"])
    r = O.plan_inner_loop(pred, emb, cp0, tmel, n)
    # the smooth init is chaotic: a 1-ulp summation-order difference flips Adam
    # signs of near-zero gradients, so only the first two steps are pinned tightly there
    k = n if tag == "real32" else 2
    np.testing.assert_allclose(r["loss"][:k, 0].numpy(), golden[f"{tag}_loss"][:k], rtol=2e-6)
    np.testing.assert_allclose(r["terms"][:k, 0, 0].numpy(), golden[f"{tag}_mel"][:k], rtol=2e-6)
    np.testing.assert_allclose(r["terms"][:k, 0, 1].numpy(), golden[f"{tag}_sem"][:k], rtol=2e-6)
    np.testing.assert_allclose(r["loss"][:, 0].numpy(), golden[f"{tag}_loss"], rtol=5e-2)
    if tag == "real32":          # the smooth init is chaotic: only the iid one is pinned on cps
        np.testing.assert_allclose(r["planned_cp"][0].numpy(), golden[f"{tag}_planned_cp"], atol=1e-6)


def test_batched_loop_and_solo_equivalence(golden, models32):
    torch.set_num_threads(1)
    pred, emb, _ = models32
    cp0, tmel = torch.from_numpy(golden["b3_cp0"]), torch.from_numpy(golden["b3_tmel"])
    r = O.plan_inner_loop(pred, emb, cp0, tmel, 5, log_grads=True)
    np.testing.assert_allclose(r["loss"].numpy(), golden["b3_loss"], rtol=2e-6)
    np.testing.assert_allclose(torch.stack(r["grads"]).numpy(), golden["b3_grads"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(r["planned_cp"].numpy(), golden["b3_planned_cp"], atol=1e-6)
    # B words batched == B solo batch-1 runs (how the reference is used)
    np.testing.assert_allclose(golden["b3_planned_cp"], golden["b3_solo_planned_cp"], atol=1e-6)


def test_model_forwards(golden, models32, models64):
    pred, emb, _ = models32
    inv = models64[2]
    with torch.no_grad():
        y = pred(torch.from_numpy(golden["fw_x"]))
        assert y.shape == (2, 8, 60)                       # odd T=17 -> floor(17/2) frames
        np.testing.assert_allclose(y.numpy(), golden["fw_y"], atol=1e-6)
        lens = tuple(torch.tensor(int(v)) for v in golden["em_lens"])
        np.testing.assert_allclose(emb(torch.from_numpy(golden["em_x"]), lens).numpy(), golden["em_y"], atol=1e-6)
        np.testing.assert_allclose(inv(torch.from_numpy(golden["inv_x"])).numpy(), golden["inv_y"], atol=1e-12)
        # the oracle's inverse model follows the input dtype (the reference's is fp64-only)
        y32 = inv.float()(torch.from_numpy(golden["inv_x"]).float())
        inv.double()
        np.testing.assert_allclose(y32.numpy(), golden["inv_y"], atol=5e-5)


def test_manual_restatement_matches_autograd(golden, models64):
    """Hand-derived BPTT / adjoint stencils / Adam == autograd + torch.optim.Adam (fp64, tight)."""
    pred, emb, _ = models64
    cp0 = torch.from_numpy(golden["b3_cp0"]).double()
    tmel = torch.from_numpy(golden["b3_tmel"]).double()
    for objective in ("acoustic_semvec", "acoustic", "semvec"):
        r = O.plan_inner_loop(pred, emb, cp0, tmel, 3, objective=objective, log_grads=True, log_cps=True)
        pw, ew = pred.state_dict(), emb.state_dict()
        x = cp0.clone()
        m = torch.zeros_like(x)
        v = torch.zeros_like(x)
        for k in range(3):
            terms, total, dcp, mel, sv = O.manual_step(pw, ew, x, tmel, r["target_semvec"], objective)
            np.testing.assert_allclose(total.numpy(), r["loss"][k].numpy(), rtol=1e-12)
            np.testing.assert_allclose(dcp.numpy(), r["grads"][k].numpy(), rtol=1e-9, atol=1e-13)
            x, m, v = O.manual_adam_step(x, dcp, m, v, k + 1)
            x = x.clamp(-O.CLAMP, O.CLAMP)
        np.testing.assert_allclose(x.numpy(), r["planned_cp"].numpy(), atol=1e-12)


def test_edge_cases():
    """Shortest trajectories the stencils accept, odd T, and the eps=0 RMSE NaN the reference has."""
    pred, emb, _ = O.build_reference_models(0, 16, torch.float64, with_inverse=False)
    cp0, tmel = O.synthetic_inputs(2, 13, seed=1, dtype=torch.float64)   # jerk needs T >= 13
    r = O.plan_inner_loop(pred, emb, cp0, tmel, 2)
    assert torch.isfinite(r["loss"]).all() and r["pred_mel"].shape == (2, 6, 60)
    # zero mel error -> RMSE(eps=0) has a NaN gradient (paule/paule.py:68); the oracle keeps that behaviour
    with torch.no_grad():
        exact = pred(cp0)
    r2 = O.plan_inner_loop(pred, emb, cp0, exact, 1, objective="acoustic")
    assert torch.isnan(r2["planned_cp"]).any()


# ---- optional loss branches against the real plan_resynth with use_speech_classifier / use_somatosensory_feedback
def test_branch_weights_regenerate_bit_identically(golden_branches):
    mods = O.build_branch_models(torch.float64)
    assert [O.state_dict_digest(m) for m in mods] == list(golden_branches["digest64"])
    np.testing.assert_array_equal(mods[3].linear.weight.numpy(), golden_branches["cls_w"])


@pytest.mark.parametrize("objective", ["acoustic_semvec", "acoustic", "semvec"])
def test_classifier_branch_matches_real_plan_resynth_fp64(golden_branches, models64, objective):
    g, (pred, emb, _) = golden_branches, models64
    cls = O.build_branch_models(torch.float64)[3]
    tag = f"cls_{objective}"
    n = len(g[f"{tag}_loss"])
    r = O.plan_inner_loop_branches(pred, emb, torch.from_numpy(g["cp0"]), torch.from_numpy(g["tmel"]), n,
                                   objective=objective, speech_classifier=cls)
    np.testing.assert_allclose(r["loss"][:, 0].numpy(), g[f"{tag}_loss"], rtol=1e-12)
    np.testing.assert_allclose(r["aux"][:, 0, 0].numpy(), g[f"{tag}_cls"], rtol=1e-12)
    np.testing.assert_allclose(r["planned_cp"][0].numpy(), g[f"{tag}_planned_cp"], atol=1e-14)


def test_somatosensory_branch_matches_real_plan_resynth_fp64(golden_branches, models64):
    g, (pred, emb, _) = golden_branches, models64
    cp_tube, tube_mel, tube_emb, _ = O.build_branch_models(torch.float64)
    n = len(g["soma_loss"])
    r = O.plan_inner_loop_branches(pred, emb, torch.from_numpy(g["cp0"]), torch.from_numpy(g["tmel"]), n,
                                   cp_tube_model=cp_tube, tube_mel_model=tube_mel, tube_embedder=tube_emb)
    np.testing.assert_allclose(r["loss"][:, 0].numpy(), g["soma_loss"], rtol=1e-12)
    np.testing.assert_allclose(r["aux"][:, 0, 1].numpy(), g["soma_tube_mel"], rtol=1e-12)
    np.testing.assert_allclose(r["aux"][:, 0, 2].numpy(), g["soma_tube_sem"], rtol=1e-12)
    np.testing.assert_allclose(r["planned_cp"][0].numpy(), g["soma_planned_cp"], atol=1e-14)


# ---- the model-path gradient: what the LSTM BPTT kernels produce, isolated from the 10^5 x larger
# smoothness gradient, pinned against the REAL reference's xx_new.grad (tests/golden/make_grad_golden.py, fp64)
@pytest.mark.parametrize("init", ["iid", "smooth"])
@pytest.mark.parametrize("objective", ["acoustic_semvec", "acoustic", "semvec"])
def test_model_path_gradient_matches_the_real_reference(golden_grad, models64, init, objective):
    pred, emb, _ = models64
    g = golden_grad
    assert [O.state_dict_digest(m) for m in models64] == list(g["digest64"])
    tmel = torch.from_numpy(g[f"{init}_tmel"])
    tag = f"{init}_{objective}"
    for k in range(g[f"{tag}_cps"].shape[0]):
        cp = torch.from_numpy(g[f"{tag}_cps"][k])[None]
        want_total, want_model = g[f"{tag}_grad"][k], g[f"{tag}_grad_model"][k]
        got_model = O.model_path_grad(pred, emb, cp, tmel, objective=objective)[0].numpy()
        scale = np.abs(want_model).max()
        assert scale > 1e-6     # the signal exists
        # the golden model part is (total - smoothness) in fp64: its absolute accuracy is ~1e-16 x |total| <= 1e-13
        np.testing.assert_allclose(got_model, want_model, atol=2e-9 * scale + 1e-13)
        got_total = got_model + O.manual_smooth_grad(cp)[0].numpy()
        np.testing.assert_allclose(got_total, want_total, rtol=1e-10, atol=1e-12)
        # the hand-derived BPTT (the formulas of the CUDA kernels) gives the same model part
        pw = {k_: v.double() for k_, v in pred.state_dict().items()}
        ew = {k_: v.double() for k_, v in emb.state_dict().items()}
        lens = (torch.tensor(tmel.shape[1]),)
        with torch.no_grad():
            tsv = emb(tmel, lens)
        _, _, dcp, _, _ = O.manual_step(pw, ew, cp, tmel, tsv, objective)
        np.testing.assert_allclose(dcp[0].numpy() - O.manual_smooth_grad(cp)[0].numpy(), want_model, atol=1e-7 * scale)


def test_model_path_gradient_ragged_words_are_planned_alone(models64):
    pred, emb, _ = models64
    gen = torch.Generator().manual_seed(3)
    cp = torch.rand(2, 30, 30, generator=gen, dtype=torch.float64) - 0.5
    tmel = torch.rand(2, 15, 60, generator=gen, dtype=torch.float64)
    g = O.model_path_grad(pred, emb, cp, tmel, lens=[30, 21])
    solo = O.model_path_grad(pred, emb, cp[1:2, :21], tmel[1:2, :10])
    np.testing.assert_allclose(g[1, :21].numpy(), solo[0].numpy(), atol=1e-15)
    assert torch.all(g[1, 21:] == 0)
