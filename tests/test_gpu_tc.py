"""tcgen05 persistent-RNN kernels (bf16 operands, fp32 accumulate / state) against the fp32 FFMA kernels.

Bounds (stated for bf16 operands; BASELINE.json north_star allows looser bounds than fp32's 1e-3): hidden / cell
states within 2e-3 / 4e-3 absolute of the fp32 kernels over the sequence, pre-activation gradients within 2e-2 of
their maximum.  Run on the B200 box: python -m pytest tests -m gpu
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from paule_b200 import _lib, ops
    _lib.require_device()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(30, 720, batch_first=True)
    w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev),
                        lstm.bias_hh_l0.to(dev), tc=True)
    return dev, _lib.load(), w


def _status(xchg):
    return int(xchg[2048:2052].view(torch.int32).item())    # kXchgErrOff: watchdog flag of the persistent kernels


# (450, 5), (1000, 4): more words than one launch holds -- passes of different layouts (tc_lstm.cuh, plan_passes)
@pytest.mark.parametrize("B,T", [(64, 24), (1, 9), (37, 16), (100, 12), (130, 7), (300, 6), (450, 5), (1000, 4)])
def test_tc_forward_and_backward_match_fp32_kernels(setup, B, T):
    from paule_b200 import _lib, ops
    dev, lib, w = setup
    H = 720
    g = torch.Generator(device="cpu").manual_seed(B * 100 + T)
    xp = (torch.randn(T, B, 4 * H, generator=g) * 0.5).to(dev)
    st = ops._stream()
    g0, h0, c0 = xp.clone(), torch.empty(T, B, H, device=dev), torch.empty(T, B, H, device=dev)
    _lib.check(lib.paule_lstm_seq_fwd_f32(g0.data_ptr(), w.w_hh.data_ptr(), h0.data_ptr(), c0.data_ptr(), T, B, H, st))
    g1, h1, c1 = xp.clone(), torch.zeros(T, B, H, device=dev), torch.zeros(T, B, H, device=dev)
    xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_lstm_seq_fwd(g1.data_ptr(), w.packed.data_ptr(), h1.data_ptr(), c1.data_ptr(),
                                         xchg.data_ptr(), None, T, B, 1,st))
    torch.cuda.synchronize()
    assert _status(xchg) == 0, "persistent kernel watchdog fired"
    np.testing.assert_allclose(h1.cpu().numpy(), h0.cpu().numpy(), atol=2e-3)
    np.testing.assert_allclose(c1.cpu().numpy(), c0.cpu().numpy(), atol=4e-3)
    np.testing.assert_allclose(g1.cpu().numpy(), g0.cpu().numpy(), atol=4e-3)

    # backward on the SAME (fp32-kernel) stash so that only the recurrent GEMM precision differs
    dh_seq = (torch.randn(T // 2, B, H, generator=g) * 1e-2).to(dev)
    dh_last = (torch.randn(B, H, generator=g) * 1e-2).to(dev)
    d0, scratch = g0.clone(), torch.empty(B, H, device=dev)
    _lib.check(lib.paule_lstm_seq_bwd_f32(d0.data_ptr(), c0.data_ptr(), w.w_hh_t.data_ptr(), dh_seq.data_ptr(), 2,
                                          dh_last.data_ptr(), scratch.data_ptr(), T, B, H, st))
    d1 = g0.clone()
    _lib.check(lib.paule_tc_lstm_seq_bwd(d1.data_ptr(), c0.data_ptr(), w.packed.data_ptr(), dh_seq.data_ptr(), 2,
                                         dh_last.data_ptr(), xchg.data_ptr(), None, T, B, 1,st))
    torch.cuda.synchronize()
    assert _status(xchg) == 0, "persistent kernel watchdog fired"
    scale = d0.abs().max().item()
    np.testing.assert_allclose(d1.cpu().numpy(), d0.cpu().numpy(), atol=2e-2 * scale)
    # mode 1 (full-rate external gradient), no dh_last
    dh_full = (torch.randn(T, B, H, generator=g) * 1e-2).to(dev)
    d0, d1 = g0.clone(), g0.clone()
    _lib.check(lib.paule_lstm_seq_bwd_f32(d0.data_ptr(), c0.data_ptr(), w.w_hh_t.data_ptr(), dh_full.data_ptr(), 1, None,
                                          scratch.data_ptr(), T, B, H, st))
    _lib.check(lib.paule_tc_lstm_seq_bwd(d1.data_ptr(), c0.data_ptr(), w.packed.data_ptr(), dh_full.data_ptr(), 1, None,
                                         xchg.data_ptr(), None, T, B, 1,st))
    torch.cuda.synchronize()
    np.testing.assert_allclose(d1.cpu().numpy(), d0.cpu().numpy(), atol=2e-2 * d0.abs().max().item())


def test_tc_is_repeatable(setup):
    """Same inputs -> the same outputs across launches up to fp32 summation order: the 12 MMA issuers of a step add their
    k-blocks into one shared TMEM accumulator in arrival order, so the last bits may differ (like a split-K GEMM).  A
    last-bit difference occasionally flips the bf16 rounding of one h value (2^-9 relative), which the recurrence then
    carries: the first steps must agree to fp32 round-off, the whole sequence to a few bf16 flips."""
    from paule_b200 import _lib, ops
    dev, lib, w = setup
    H, B, T = 720, 64, 40
    xp = torch.randn(T, B, 4 * H, device=dev) * 0.5
    outs = []
    for _ in range(3):
        g1, h1, c1 = xp.clone(), torch.zeros(T, B, H, device=dev), torch.zeros(T, B, H, device=dev)
        xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
        _lib.check(lib.paule_tc_lstm_seq_fwd(g1.data_ptr(), w.packed.data_ptr(), h1.data_ptr(), c1.data_ptr(),
                                             xchg.data_ptr(), None, T, B, 1,ops._stream()))
        torch.cuda.synchronize()
        outs.append(h1)
    for other in outs[1:]:
        d = (outs[0] - other).abs()
        assert d[:2].max().item() < 2e-6, "steps 0 and 1 see only fp32 summation-order noise"
        assert d.max().item() < 5e-4, "later steps: a few bf16 rounding flips of h, never more"
        assert (d > 5e-5).float().mean().item() < 1e-2


@pytest.mark.parametrize("B,T", [(64, 10), (37, 7), (100, 5), (500, 4)])   # 500: image groups that straddle two passes
def test_tc_gemm_over_image_sequences(setup, B, T):
    """paule_tc_gemm_img on the bf16 images a persistent kernel left behind == fp32 GEMM on the fp32 tensors
    (K = 720 input projection with bias; K = 2880 dX with N = 720 / 60 / 30, overwrite and accumulate)."""
    from paule_b200 import _lib, ops
    dev, lib, w = setup
    H = 720
    g = torch.Generator(device="cpu").manual_seed(B + T)
    st = ops._stream()
    # forward images of h
    xp = (torch.randn(T, B, 4 * H, generator=g) * 0.5).to(dev)
    gates, h, c = xp.clone(), torch.zeros(T, B, H, device=dev), torch.zeros(T, B, H, device=dev)
    xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
    himg = torch.zeros(lib.paule_tc_img_seq_bytes(T, B, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_lstm_seq_fwd(gates.data_ptr(), w.packed.data_ptr(), h.data_ptr(), c.data_ptr(),
                                         xchg.data_ptr(), himg.data_ptr(), T, B, 1, st))
    W = (torch.randn(4 * H, H, generator=g) / H ** 0.5).to(dev)
    bias = torch.randn(4 * H, generator=g).to(dev)
    pk = torch.empty(lib.paule_tc_gemm_packed_bytes(4 * H, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_gemm_pack(W.data_ptr(), pk.data_ptr(), 4 * H, 1, st))
    out = torch.full((T, B, 4 * H), float("nan"), device=dev)
    _lib.check(lib.paule_tc_gemm_img(himg.data_ptr(), pk.data_ptr(), bias.data_ptr(), out.data_ptr(), T, B, 4 * H, 1, 0, st))
    torch.cuda.synchronize()
    ref = h.double() @ W.double().t() + bias.double()
    np.testing.assert_allclose(out.cpu().double().numpy(), ref.cpu().numpy(), atol=1e-2, rtol=1e-2)
    # backward images of dA, three widths of W_ih^T
    dh = (torch.randn(T, B, H, generator=g) * 1e-2).to(dev)
    da = gates.clone()
    daimg = torch.zeros(lib.paule_tc_img_seq_bytes(T, B, 4), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_lstm_seq_bwd(da.data_ptr(), c.data_ptr(), w.packed.data_ptr(), dh.data_ptr(), 1, None,
                                         xchg.data_ptr(), daimg.data_ptr(), T, B, 1, st))
    for N in (720, 60, 30):
        Wt = (torch.randn(N, 4 * H, generator=g) / H ** 0.5).to(dev)
        pk2 = torch.empty(lib.paule_tc_gemm_packed_bytes(N, 4), dtype=torch.uint8, device=dev)
        _lib.check(lib.paule_tc_gemm_pack(Wt.data_ptr(), pk2.data_ptr(), N, 4, st))
        base = torch.randn(T, B, N, generator=g).to(dev)
        o1, o2 = torch.full((T, B, N), float("nan"), device=dev), base.clone()
        _lib.check(lib.paule_tc_gemm_img(daimg.data_ptr(), pk2.data_ptr(), None, o1.data_ptr(), T, B, N, 4, 0, st))
        _lib.check(lib.paule_tc_gemm_img(daimg.data_ptr(), pk2.data_ptr(), None, o2.data_ptr(), T, B, N, 4, 1, st))
        torch.cuda.synchronize()
        ref = da.double() @ Wt.double().t()
        scale = ref.abs().max().item()
        np.testing.assert_allclose(o1.cpu().double().numpy(), ref.cpu().numpy(), atol=1e-2 * scale)
        np.testing.assert_allclose(o2.cpu().double().numpy(), (ref + base.double()).cpu().numpy(), atol=1e-2 * scale + 1e-6)


@pytest.mark.parametrize("B,T,I", [(64, 12, 30), (37, 9, 60), (1, 5, 30), (130, 6, 30), (300, 5, 60)])
def test_tc_fused_input_projection(B, T, I):
    """paule_tc_lstm_seq_fwd_x (W_ih x_t inside the recurrence: hi/lo-split bf16 x for I <= 32, plain bf16 for I <= 64, bf16
    W_ih) against the unfused kernel fed with the fp32 projection x W_ih^T + b."""
    from paule_b200 import _lib, ops
    _lib.require_device()
    dev, lib, H = torch.device("cuda:0"), _lib.load(), 720
    torch.manual_seed(I)
    lstm = torch.nn.LSTM(I, H, batch_first=True)
    w = ops.LstmWeights(lstm.weight_ih_l0.to(dev), lstm.weight_hh_l0.to(dev), lstm.bias_ih_l0.to(dev), lstm.bias_hh_l0.to(dev), tc=True)
    g = torch.Generator(device="cpu").manual_seed(B * 7 + T)
    x = (torch.rand(T, B, I, generator=g) - 0.5).to(dev)
    st = ops._stream()
    xp = (x.reshape(-1, I) @ w.w_ih.t() + w.bias).reshape(T, B, 4 * H).contiguous()
    g0, h0, c0 = xp.clone(), torch.zeros(T, B, H, device=dev), torch.zeros(T, B, H, device=dev)
    xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_lstm_seq_fwd(g0.data_ptr(), w.packed.data_ptr(), h0.data_ptr(), c0.data_ptr(), xchg.data_ptr(), None, T, B, 1, st))
    ximg = torch.zeros(lib.paule_tc_x_image_bytes(T, B), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_x_image(x.data_ptr(), ximg.data_ptr(), T, B, I, st))
    g1 = torch.full((T, B, 4 * H), float("nan"), device=dev)
    h1, c1 = torch.zeros(T, B, H, device=dev), torch.zeros(T, B, H, device=dev)
    himg = torch.zeros(lib.paule_tc_img_seq_bytes(T, B, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_lstm_seq_fwd_x(g1.data_ptr(), w.packed.data_ptr(), w.bias.data_ptr(), ximg.data_ptr(), h1.data_ptr(),
                                           c1.data_ptr(), xchg.data_ptr(), himg.data_ptr(), T, B, 1, st))
    torch.cuda.synchronize()
    assert _status(xchg) == 0, "persistent kernel watchdog fired"
    tol = 1e-3 if I <= 32 else 3e-3     # plain-bf16 x (the mel) rounds the input to 2^-9 relative
    np.testing.assert_allclose(h1.cpu().numpy(), h0.cpu().numpy(), atol=tol)
    np.testing.assert_allclose(c1.cpu().numpy(), c0.cpu().numpy(), atol=2 * tol)
    np.testing.assert_allclose(g1.cpu().numpy(), g0.cpu().numpy(), atol=2 * tol)
    # the first step has no recurrent part: it isolates the projection itself
    np.testing.assert_allclose(g1[0].cpu().numpy(), g0[0].cpu().numpy(), atol=5e-4 if I <= 32 else 1.5e-3)   # bf16 W_ih


@pytest.mark.parametrize("hidden,layers", [(360, 1), (180, 4)])
def test_narrow_models_run_zero_padded_on_the_tensor_core_kernels(setup, hidden, layers):
    """Hidden sizes below 720 (the somatosensory models' 360 units, ForwardModel's default 4 x 180: paule/paule.py:231-249,
    paule/models.py:335-339) run on the persistent tcgen05 kernels zero-padded to 720 units -- exact up to the bf16 operand
    rounding: forward and the gradient wrt the input against torch.nn.LSTM on the CPU."""
    import paule_b200 as P
    from paule_b200 import ops
    dev = setup[0]
    torch.manual_seed(5)
    mine = P.ForwardModel(num_lstm_layers=layers, hidden_size=hidden)
    ref = torch.nn.LSTM(30, hidden, num_layers=layers, batch_first=True)
    ref.load_state_dict(mine.lstm.state_dict())
    lin = torch.nn.Linear(hidden, 60)
    lin.load_state_dict(mine.post_linear.state_dict())
    for p in mine.parameters():
        p.requires_grad_(False)
    mine = mine.to(dev)
    mine.math = ops.MATH_BF16
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(3, 24, 30, generator=g) - 0.5)
    xr = x.clone().requires_grad_()
    yr = torch.nn.functional.avg_pool1d(lin(ref(xr)[0]).transpose(1, 2), 2, stride=2).transpose(1, 2)
    w = torch.rand(yr.shape, generator=g) - 0.5
    (yr * w).sum().backward()
    xg = x.to(dev).requires_grad_()
    y = mine(xg)
    assert mine._pack.padded_from == hidden and mine._pack.layers[0].packed is not None
    (y * w.to(dev)).sum().backward()
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.detach().numpy(), atol=3e-3)
    scale = xr.grad.abs().max().item()
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xr.grad.numpy(), atol=3e-2 * scale)
    with pytest.raises(AssertionError):
        np.testing.assert_allclose(0 * xg.grad.cpu().numpy(), xr.grad.numpy(), atol=3e-2 * scale)


@pytest.mark.parametrize("B,T,N,K", [(37, 7, 720, 60), (100, 5, 720, 60), (500, 4, 720, 60), (64, 6, 2880, 30), (16, 3, 60, 64)])
def test_tc_gemm_narrow_k(setup, B, T, N, K):
    """paule_tc_a_image + paule_tc_gemm_img_k64 (K <= 64: one k-block; post_linear^T of the serial backward) == fp32 GEMM, overwrite
    and accumulate, with and without bias."""
    from paule_b200 import _lib, ops
    dev, lib, _ = setup
    g = torch.Generator(device="cpu").manual_seed(B * 7 + T)
    st = ops._stream()
    x = torch.randn(T, B, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    img = torch.zeros(lib.paule_tc_a_image_bytes(T, B), dtype=torch.uint8, device=dev)
    pk = torch.empty(lib.paule_tc_gemm_packed_bytes_k64(N), dtype=torch.uint8, device=dev)
    _lib.check(lib.paule_tc_gemm_pack_k64(W.data_ptr(), pk.data_ptr(), N, K, st))
    _lib.check(lib.paule_tc_a_image(x.data_ptr(), img.data_ptr(), T, B, K, st))
    base = torch.randn(T, B, N, generator=g).to(dev)
    o1, o2 = torch.full((T, B, N), float("nan"), device=dev), base.clone()
    _lib.check(lib.paule_tc_gemm_img_k64(img.data_ptr(), pk.data_ptr(), bias.data_ptr(), o1.data_ptr(), T, B, N, 0, st))
    _lib.check(lib.paule_tc_gemm_img_k64(img.data_ptr(), pk.data_ptr(), None, o2.data_ptr(), T, B, N, 1, st))
    torch.cuda.synchronize()
    ref = x.double() @ W.double().t()
    scale = ref.abs().max().item()
    np.testing.assert_allclose(o1.cpu().double().numpy(), (ref + bias.double()).cpu().numpy(), atol=1e-2 * scale)
    np.testing.assert_allclose(o2.cpu().double().numpy(), (ref + base.double()).cpu().numpy(), atol=1e-2 * scale)
