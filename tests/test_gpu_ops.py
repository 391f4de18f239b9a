"""Op-level parity of the CUDA kernels (through the C ABI / torch.library ops) against the CPU oracle.

Run on the B200 box: python -m pytest tests -m gpu
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


def _np(t):
    return t.detach().cpu().double().numpy()


def test_library_loaded_and_device_ok(dev):
    from paule_b200 import _lib
    lib = _lib.load()
    assert lib.paule_version() >= 100
    assert lib.paule_device_check() == 0


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 33, 30), (200, 2880, 30), (64, 300, 720), (130, 60, 2880), (96, 720, 60)])
def test_linear_rows_plain(dev, M, N, K):
    from paule_b200 import ops
    g = torch.Generator().manual_seed(M * 131 + N * 7 + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    out = torch.full((M, N), float("nan"), device=dev)
    ops.linear_rows_(out, a.to(dev), w.to(dev), b.to(dev), M, (1, K, 0), (1, N, 0))
    ref = a.double() @ w.double().t() + b.double()
    np.testing.assert_allclose(_np(out), ref.numpy(), rtol=2e-5, atol=2e-5)
    # accumulate and no bias
    out2 = out.clone()
    ops.linear_rows_(out2, a.to(dev), w.to(dev), None, M, (1, K, 0), (1, N, 0), accumulate=True)
    np.testing.assert_allclose(_np(out2), (2 * ref - b.double()).numpy(), rtol=4e-5, atol=4e-5)


def test_linear_rows_maps_and_pooling(dev):
    """batch-first -> time-major on load, pair pooling on load, batch-first on store."""
    from paule_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, T, K, N = 5, 9, 30, 17
    x = torch.randn(B, T, K, generator=g)
    w = torch.randn(N, K, generator=g)
    b = torch.randn(N, generator=g)
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    y = torch.empty(T, B, N, device=dev)
    ops.linear_rows_(y, xd, wd, bd, T * B, (B, K, T * K), (1, N, 0))
    ref = (x.double() @ w.double().t() + b.double()).transpose(0, 1)
    np.testing.assert_allclose(_np(y), ref.numpy(), rtol=1e-5, atol=1e-5)
    # pooled: time-major input [T,B,K] -> [T//2,B,N], and batch-first store
    xt = x.transpose(0, 1).contiguous()
    yp = ops.linear_tm(xt.to(dev), wd, bd, True, False)
    pooled = 0.5 * (xt[0:2 * (T // 2):2] + xt[1:2 * (T // 2):2]).double()
    refp = pooled @ w.double().t() + b.double()
    assert yp.shape == (T // 2, B, N)
    np.testing.assert_allclose(_np(yp), refp.numpy(), rtol=1e-5, atol=1e-5)
    ybf = ops.linear_tm(xt.to(dev), wd, bd, True, True)
    np.testing.assert_allclose(_np(ybf), refp.transpose(0, 1).numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(_np(ops.transpose_btc(xd)), xt.double().numpy())


@pytest.mark.parametrize("B,T,I,H", [(1, 5, 30, 720), (3, 14, 60, 720), (33, 6, 720, 720), (4, 9, 10, 48), (2, 7, 180, 100)])
def test_lstm_layer_forward_backward(dev, B, T, I, H):
    """One LSTM layer: forward vs torch.nn.LSTM (CPU, the reference's operator) and input-gradient BPTT vs autograd."""
    from paule_b200 import ops
    torch.manual_seed(B * 1000 + T * 10 + I + H)
    lstm = torch.nn.LSTM(I, H, num_layers=1, batch_first=True).double()
    x = (torch.rand(B, T, I, dtype=torch.double) - 0.5).requires_grad_()
    out, _ = lstm(x)
    gout = torch.randn(B, T, H, dtype=torch.double)
    (out * gout).sum().backward()

    w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev),
                        lstm.bias_hh_l0.to(dev))
    xd = x.detach().float().to(dev).requires_grad_()
    h, gates, c = ops.lstm_layer_fwd(xd, True, w.w_ih, w.w_hh, w.bias)
    assert h.shape == (T, B, H) and gates.shape == (T, B, 4 * H) and c.shape == (T, B, H)
    np.testing.assert_allclose(_np(h).transpose(1, 0, 2), out.detach().numpy(), rtol=2e-5, atol=2e-6)
    (h * gout.float().transpose(0, 1).to(dev)).sum().backward()
    scale = x.grad.abs().max().item()
    np.testing.assert_allclose(_np(xd.grad), x.grad.numpy(), rtol=1e-4, atol=1e-5 * scale)
    # stash untouched by backward (the op works on a copy), second backward gives the same result
    h2, gates2, c2 = ops.lstm_layer_fwd(xd.detach(), True, w.w_ih, w.w_hh, w.bias)
    assert torch.equal(gates2, gates) and torch.equal(c2, c)


def test_lstm_backward_pooled_and_last_modes(dev):
    """dh_mode 2 (adjoint of pair pooling, odd T) and dh_last (embedder head) of paule_lstm_seq_bwd_f32."""
    from paule_b200 import _lib, ops
    torch.manual_seed(5)
    B, T, I, H = 3, 7, 12, 40
    lstm = torch.nn.LSTM(I, H, batch_first=True).double()
    x = (torch.rand(B, T, I, dtype=torch.double) - 0.5).requires_grad_()
    out, _ = lstm(x)
    Tm = T // 2
    gp = torch.randn(B, Tm, H, dtype=torch.double)
    gl = torch.randn(B, H, dtype=torch.double)
    pooled = 0.5 * (out[:, 0:2 * Tm:2] + out[:, 1:2 * Tm:2])
    ((pooled * gp).sum() + (out[:, -1] * gl).sum()).backward()

    w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev),
                        lstm.bias_hh_l0.to(dev))
    h, gates, c = ops.lstm_layer_fwd(x.detach().float().to(dev), True, w.w_ih, w.w_hh, w.bias)
    lib = _lib.load()
    da = gates.clone()
    scratch = torch.empty(B, H, device=dev)
    dh_seq = gp.float().transpose(0, 1).contiguous().to(dev)
    dh_last = gl.float().to(dev)
    _lib.check(lib.paule_lstm_seq_bwd_f32(da.data_ptr(), c.data_ptr(), w.w_hh_t.data_ptr(), dh_seq.data_ptr(), 2,
                                          dh_last.data_ptr(), scratch.data_ptr(), T, B, H, ops._stream()))
    dx = torch.empty(B, T, I, device=dev)
    ops.linear_rows_(dx, da, w.w_ih_t, None, T * B, (1, 4 * H, 0), (B, I, T * I))
    np.testing.assert_allclose(_np(dx), x.grad.numpy(), rtol=1e-4, atol=1e-5 * x.grad.abs().max().item())


@pytest.mark.parametrize("objective", ["acoustic_semvec", "acoustic", "semvec"])
@pytest.mark.parametrize("B,T", [(1, 13), (3, 40), (2, 77), (5, 200)])
def test_plan_loss_and_gradients(dev, objective, B, T):
    """criterion + analytic gradients vs autograd of the oracle's per_word_losses (fp64 truth)."""
    from paule_b200 import ops
    g = torch.Generator().manual_seed(T + B)
    Tm = T // 2
    mel = torch.rand(B, Tm, 60, generator=g, dtype=torch.double).requires_grad_()
    tmel = torch.rand(B, Tm, 60, generator=g, dtype=torch.double)
    sv = torch.randn(B, 300, generator=g, dtype=torch.double).requires_grad_()
    tsv = torch.randn(B, 300, generator=g, dtype=torch.double)
    cp = (torch.rand(B, T, 30, generator=g, dtype=torch.double) - 0.5).requires_grad_()
    total, terms = O.per_word_losses(mel, tmel, sv, tsv, cp, objective)
    total.sum().backward()

    def tm(t):
        return t.detach().float().transpose(0, 1).contiguous().to(dev)

    out_terms, dmel, dsv, dcp = ops.plan_loss(tm(mel), tm(tmel), sv.detach().float().to(dev), tsv.float().to(dev),
                                              tm(cp), ops.OBJECTIVES[objective])
    np.testing.assert_allclose(_np(out_terms[:, 0]), total.detach().numpy(), rtol=2e-5)
    np.testing.assert_allclose(_np(out_terms[:, 1:]), terms.detach().numpy(), rtol=2e-5)
    zero = torch.zeros(1, dtype=torch.double)
    gm = mel.grad if mel.grad is not None else zero.expand_as(mel)
    gs = sv.grad if sv.grad is not None else zero.expand_as(sv)
    np.testing.assert_allclose(_np(dmel).transpose(1, 0, 2), gm.numpy(), rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(_np(dsv), gs.numpy(), rtol=2e-5, atol=1e-9)
    scale = cp.grad.abs().max().item()
    np.testing.assert_allclose(_np(dcp).transpose(1, 0, 2), cp.grad.numpy(), rtol=1e-4, atol=2e-6 * scale)


@pytest.mark.parametrize("smiling,past", [(False, 0), (True, 0), (False, 4), (True, 6)])
def test_adam_clamp_matches_torch_optim(dev, smiling, past):
    """Five Adam steps + clamp (+ smiling, past_cp) vs torch.optim.Adam on CPU fp32 (paule.py:1199-1211)."""
    from paule_b200 import ops
    g = torch.Generator().manual_seed(17)
    T, B, Cc = 21, 3, 30
    x0 = (torch.rand(T, B, Cc, generator=g) - 0.5) * 2.2
    past_cp = torch.rand(past, B, Cc, generator=g) if past else None
    x = x0.clone().requires_grad_()
    opt = torch.optim.Adam([x], lr=0.01)
    xd = x0.clone().to(dev)
    m, v = torch.zeros_like(xd), torch.zeros_like(xd)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    for k in range(5):
        ga = torch.randn(T, B, Cc, generator=g) * (10.0 ** (k - 3))
        gb = torch.randn(T, B, Cc, generator=g) * 1e-3
        opt.zero_grad()
        x.grad = (ga + gb).clone()
        opt.step()
        with torch.no_grad():
            x.data = x.data.clamp(-1.05, 1.05)
            if smiling:
                x.data[:, :, 4] = -1.0
                x.data[:, :, 1] = 1.0
            if past:
                x.data[0:past] = past_cp
        ops.adam_clamp_(xd, ga.to(dev), gb.to(dev), m, v, step, 0.01, 0.9, 0.999, 1e-8, 1.05, smiling,
                        past_cp.to(dev) if past else None)
        np.testing.assert_allclose(_np(xd), x.detach().double().numpy(), rtol=0, atol=2e-7)
    assert int(step.item()) == 5


def test_model_forwards_match_reference_golden(dev, golden):
    """ForwardModel / EmbeddingModel / InverseModel on CUDA vs outputs of the REFERENCE's modules (golden)."""
    import paule_b200 as P
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=720)
    assert [O.state_dict_digest(m) for m in (pred, emb, inv)] == list(golden["digest32"])
    pred, emb, inv = pred.to(dev), emb.to(dev), inv.to(dev)
    with torch.no_grad():
        y = pred(torch.from_numpy(golden["fw_x"]).to(dev))
        assert y.shape == (2, 8, 60)
        np.testing.assert_allclose(_np(y), golden["fw_y"], atol=2e-6)
        lens = tuple(torch.tensor(int(v)) for v in golden["em_lens"])      # ragged lens (variable-length words)
        sv = emb(torch.from_numpy(golden["em_x"]).to(dev), lens)
        np.testing.assert_allclose(_np(sv), golden["em_y"], atol=2e-6)
        cp = inv(torch.from_numpy(golden["inv_x"]).float().to(dev))
        assert cp.shape == (2, 24, 30)
        np.testing.assert_allclose(_np(cp), golden["inv_y"], atol=5e-5)


def test_module_autograd_input_gradient(dev):
    """d loss / d cp through ForwardModel -> EmbeddingModel on CUDA == autograd of the oracle modules (fp64)."""
    import paule_b200 as P
    pred64, emb64, _ = O.build_reference_models(0, 720, torch.float64, with_inverse=False)
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720).to(dev)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720).to(dev)
    cp0, tmel = O.synthetic_inputs(2, 24, seed=9, dtype=torch.float64)
    x = cp0.clone().requires_grad_()
    lens = (torch.tensor(12), torch.tensor(12))
    mel = pred64(x)
    sv = emb64(mel, lens)
    ((mel - tmel) ** 2).sum().add((sv ** 2).sum()).backward()
    xd = cp0.float().to(dev).requires_grad_()
    meld = pred(xd)
    svd = emb(meld, lens)
    ((meld - tmel.float().to(dev)) ** 2).sum().add((svd ** 2).sum()).backward()
    np.testing.assert_allclose(_np(meld), mel.detach().numpy(), atol=2e-6)
    np.testing.assert_allclose(_np(xd.grad), x.grad.numpy(), rtol=2e-4, atol=1e-5 * x.grad.abs().max().item())


def test_cpu_tensors_are_rejected(dev):
    import paule_b200 as P
    from paule_b200 import _lib
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=32)
    with pytest.raises(_lib.PauleB200Error):
        pred(torch.zeros(1, 20, 30))


def test_mel_embedding_model_matches_reference_golden(dev):
    """MelEmbeddingModelMelSmoothResidualUpsampling forward (mel-channel residual convs -> LSTM stack -> h at lens-1 ->
    post_linear -> LeakyReLU -> upsampling) against the reference's own output on the same seeded weights, ragged lens."""
    import os
    import paule_b200 as P
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mel_embedder_golden.npz"))
    torch.manual_seed(7)
    m = P.MelEmbeddingModelMelSmoothResidualUpsampling(hidden_size=96, num_lstm_layers=2, post_upsampling_size=256).to(dev)
    y = m(torch.from_numpy(g["x"]).to(dev), [int(l) for l in g["lens"]])
    np.testing.assert_allclose(_np(y), g["y64"], atol=2e-5)
    np.testing.assert_allclose(_np(y), g["y32"].astype(np.float64), atol=2e-5)


@pytest.mark.parametrize("B,T", [(2, 17), (70, 9)])
def test_module_math_bf16_runs_the_tensor_core_recurrences(dev, B, T):
    """``module.math = MATH_BF16``: ForwardModel / EmbeddingModel / InverseModel forward (and the input gradient) on the
    persistent tcgen05 kernels against the same modules on the fp32 kernels.  Bounds as in test_gpu_tc.py (bf16 operands)."""
    import paule_b200 as P
    from paule_b200 import ops
    torch.manual_seed(0)
    pred = P.ForwardModel(num_lstm_layers=1, hidden_size=720).to(dev)
    emb = P.EmbeddingModel(num_lstm_layers=2, hidden_size=720).to(dev)
    inv = P.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=720).to(dev)
    g = torch.Generator().manual_seed(B + T)
    x = (torch.rand(B, T, 30, generator=g) - 0.5).to(dev)
    mel = torch.rand(B, T // 2, 60, generator=g).to(dev)
    lens = [T // 2] * B
    out = {}
    for math in (ops.MATH_FP32, ops.MATH_BF16):
        for m in (pred, emb, inv):
            m.math = math
        xr = x.clone().requires_grad_()
        y = pred(xr)                                         # [B,T//2,60]
        sv = emb(y, lens)                                    # [B,300]: ForwardModel -> EmbeddingModel, as in the planner
        (sv.square().sum() + y.square().sum()).backward()
        with torch.no_grad():
            cp = inv(mel)
        out[math] = (y.detach(), sv.detach(), xr.grad.detach(), cp)
    (y0, s0, g0, c0), (y1, s1, g1, c1) = out[ops.MATH_FP32], out[ops.MATH_BF16]
    assert not torch.equal(y0, y1), "bf16 mode produced bit-identical results: the tensor-core path did not run"
    np.testing.assert_allclose(y1.cpu().numpy(), y0.cpu().numpy(), atol=2e-3)
    np.testing.assert_allclose(s1.cpu().numpy(), s0.cpu().numpy(), atol=2e-3)
    np.testing.assert_allclose(c1.cpu().numpy(), c0.cpu().numpy(), atol=3e-3)
    np.testing.assert_allclose(g1.cpu().numpy(), g0.cpu().numpy(), atol=2e-2 * g0.abs().max().item())
    # weight gradients (continue-learning) flow through the tensor-core op as well
    pred.math = ops.MATH_BF16
    for p_ in pred.parameters():
        p_.requires_grad_(True)
        p_.grad = None            # (the loop above already accumulated into .grad: parameters require grad by default)
    pred(x).square().sum().backward()
    gw = pred.lstm.weight_hh_l0.grad
    assert gw is not None and torch.isfinite(gw).all() and gw.abs().max() > 0
    pred.math = ops.MATH_FP32
    for p_ in pred.parameters():
        p_.grad = None
    pred(x).square().sum().backward()
    gw0 = pred.lstm.weight_hh_l0.grad
    np.testing.assert_allclose(gw.cpu().numpy(), gw0.cpu().numpy(), atol=3e-2 * gw0.abs().max().item())
