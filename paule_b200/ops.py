"""torch.library custom ops (namespace ``paule_b200::``) over the C ABI of libpaule_b200.so.

PyTorch is plumbing here: it owns the device buffers and the stream; every op below passes raw
device pointers to a hand-written sm_100a kernel.  CPU tensors are rejected -- there is no fallback.

Operator surface:
  lstm_layer_fwd / lstm_layer_bwd   one LSTM layer over a sequence (time-major), forward and input-gradient BPTT
  linear_rows                       y = x W^T + b with mapped rows (batch-first <-> time-major, pair pooling)
  plan_loss, adam_clamp_            criterion + analytic gradients; Adam + clamp (+ smiling / past_cp)
  plan_step / plan_forward          the fused inner step on a registered PlanContext
"""
from __future__ import annotations

import contextlib
import functools
import warnings
import weakref
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

MATH_FP32, MATH_BF16, MATH_BF16X3 = 0, 1, 2
OBJECTIVES = {"acoustic_semvec": 0, "acoustic": 1, "semvec": 2}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on(t: torch.Tensor):
    """Device guard: the C side launches on the CURRENT CUDA device and stream, so every op body runs with its tensor's
    device current (a module on cuda:1 while cuda:0 is current must not launch on device 0 with device-1 pointers)."""
    if isinstance(t, torch.Tensor) and t.is_cuda and t.device.index != torch.cuda.current_device():
        return torch.cuda.device(t.device)
    return contextlib.nullcontext()


def _guard(fn):
    """Run an op body with its first tensor argument's device current (see _on)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = next((a for a in args if isinstance(a, torch.Tensor)), None)
        with _on(t):
            return fn(*args, **kwargs)
    return wrapper


def tc_available() -> bool:
    """True when the tcgen05 recurrent kernels are in the library (hidden size 720)."""
    return _lib.load().paule_tc_packed_lstm_bytes(720, 30) > 0


def default_math(hidden_size: int) -> int:
    """The arithmetic a planner / module uses when none is given: the tensor-core path (bf16 operands, fp32 accumulate and
    state) for the hidden size it is built for (Paule's models, paule/paule.py:124,167), the fp32 kernels otherwise."""
    return MATH_BF16 if (int(hidden_size) == 720 and tc_available()) else MATH_FP32


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.PauleB200Error(f"{name} must be a CUDA tensor: paule_b200 has no CPU fallback "
                                  f"(the CPU restatement lives in oracle/ and is test infrastructure only)")
    if t.dtype != dtype:
        raise _lib.PauleB200Error(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.PauleB200Error(f"{name} must be contiguous")
    return t


# ------------------------------------------------------------------------------------------------
# raw wrappers (no autograd)
# ------------------------------------------------------------------------------------------------
def linear_rows_(out: torch.Tensor, a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], M: int,
                 a_map: Tuple[int, int, int], c_map: Tuple[int, int, int], a_pair: int = 0, accumulate: bool = False,
                 a_offset: int = 0) -> None:
    """out[cmap(r), :] (+)= a'[amap(r), :] @ w.T + bias; see paule_linear_f32 in include/paule_b200.h."""
    lib = _lib.load()
    N, K = w.shape
    with _on(out):
        _lib.check(lib.paule_linear_f32(a.data_ptr() + 4 * a_offset, w.data_ptr(), _p(bias), out.data_ptr(), M, N, K,
                                        a_map[0], a_map[1], a_map[2], a_pair, c_map[0], c_map[1], c_map[2],
                                        1 if accumulate else 0, _stream()), "paule_linear_f32")


def transpose_btc(x: torch.Tensor) -> torch.Tensor:
    """[B,T,C] -> [T,B,C] (or back)."""
    _chk(x, "x")
    n0, n1, c = x.shape
    out = torch.empty((n1, n0, c), device=x.device, dtype=x.dtype)
    with _on(x):
        _lib.check(_lib.load().paule_transpose_btc(x.data_ptr(), out.data_ptr(), n0, n1, c, _stream()), "paule_transpose_btc")
    return out


TC_HIDDEN = 720      # hidden size the persistent tcgen05 recurrent kernels are built for


def pad_lstm_params(w_ih: torch.Tensor, w_hh: torch.Tensor, b_ih: torch.Tensor, b_hh: torch.Tensor, input_padded: bool):
    """An LSTM layer with h <= 720 hidden units as a 720-unit layer: gate blocks of 720 rows, the layer's units first, zeros
    behind.  EXACT: a unit whose weights and biases are zero has pre-activation 0, so i = f = o = 1/2, g = 0, c = 0, h = 0 at
    every step; it receives no gradient and passes none on (its W_hh column is zero).  ``input_padded``: the layer's input is
    the padded output of the layer below (input columns padded likewise)."""
    H = TC_HIDDEN
    w_ih, w_hh, b_ih, b_hh = (t.detach().float() for t in (w_ih, w_hh, b_ih, b_hh))
    h, i = w_hh.shape[1], w_ih.shape[1]
    if h == H and not (input_padded and i < H):
        return w_ih, w_hh, b_ih, b_hh
    W_ih, W_hh = w_ih.new_zeros(4 * H, H if input_padded else i), w_hh.new_zeros(4 * H, H)
    B_ih, B_hh = b_ih.new_zeros(4 * H), b_hh.new_zeros(4 * H)
    for g in range(4):
        W_ih[g * H:g * H + h, :i] = w_ih[g * h:(g + 1) * h]
        W_hh[g * H:g * H + h, :h] = w_hh[g * h:(g + 1) * h]
        B_ih[g * H:g * H + h] = b_ih[g * h:(g + 1) * h]
        B_hh[g * H:g * H + h] = b_hh[g * h:(g + 1) * h]
    return W_ih, W_hh, B_ih, B_hh


def gemm_tn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a [R, M], b [R, N] (row-major, contiguous) -> a^T b [M, N]: the weight-gradient reduction over all (step, word) rows
    (paule_gemm_tn_f32)."""
    _chk(a, "a"); _chk(b, "b")
    R, M = a.shape
    N = b.shape[1]
    out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    with _on(a):
        _lib.check(_lib.load().paule_gemm_tn_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), R, M, N, M, N, 0, _stream()),
                   "paule_gemm_tn_f32")
    return out


def colsum(a: torch.Tensor) -> torch.Tensor:
    """a [R, M] -> column sums [M] (bias gradient; paule_colsum_f32)."""
    _chk(a, "a")
    R, M = a.shape
    out = torch.empty((M,), device=a.device, dtype=torch.float32)
    with _on(a):
        _lib.check(_lib.load().paule_colsum_f32(a.data_ptr(), out.data_ptr(), R, M, M, 0, _stream()), "paule_colsum_f32")
    return out


class LstmWeights:
    """Device-side operand pack of one LSTM layer (fp32 originals, transposes, summed bias, tcgen05 image).

    Repacked whenever the source parameters change (continue-learning, paule/paule.py:1372-1377)."""

    def __init__(self, w_ih: torch.Tensor, w_hh: torch.Tensor, b_ih: torch.Tensor, b_hh: torch.Tensor, tc: bool = False):
        for n, t in (("w_ih", w_ih), ("w_hh", w_hh), ("b_ih", b_ih), ("b_hh", b_hh)):
            if not t.is_cuda:
                raise _lib.PauleB200Error(f"LSTM parameter {n} is on {t.device}: move the module to a CUDA device")
        self.w_ih = w_ih.detach().float().contiguous()
        self.w_hh = w_hh.detach().float().contiguous()
        self.w_ih_t = self.w_ih.t().contiguous()
        self.w_hh_t = self.w_hh.t().contiguous()
        self.bias = (b_ih.detach().float() + b_hh.detach().float()).contiguous()
        self.hidden = self.w_hh.shape[1]
        self.input_size = self.w_ih.shape[1]
        self.packed: Optional[torch.Tensor] = None        # W_hh images of the persistent-RNN kernels
        self.packed_ih: Optional[torch.Tensor] = None     # W_ih image   (input-projection GEMM, only if I == H)
        self.packed_ih_t: Optional[torch.Tensor] = None   # W_ih^T image (dX GEMM)
        if tc:
            self.pack_tc()

    def pack_tc(self) -> None:
        lib = _lib.load()
        nbytes = lib.paule_tc_packed_lstm_bytes(self.hidden, self.input_size)
        if nbytes == 0:
            raise _lib.PauleB200Error("tensor-core path is not available for this layer shape")
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=self.w_hh.device)
        _lib.check(lib.paule_tc_pack_lstm(self.w_ih.data_ptr(), self.w_hh.data_ptr(), self.packed.data_ptr(),
                                          self.hidden, self.input_size, _stream()), "paule_tc_pack_lstm")
        dev = self.w_hh.device
        if self.input_size == self.hidden:
            self.packed_ih = torch.empty(lib.paule_tc_gemm_packed_bytes(4 * self.hidden, 1), dtype=torch.uint8, device=dev)
            _lib.check(lib.paule_tc_gemm_pack(self.w_ih.data_ptr(), self.packed_ih.data_ptr(), 4 * self.hidden, 1,
                                              _stream()), "paule_tc_gemm_pack")
        self.packed_ih_t = torch.empty(lib.paule_tc_gemm_packed_bytes(self.input_size, 4), dtype=torch.uint8, device=dev)
        _lib.check(lib.paule_tc_gemm_pack(self.w_ih_t.data_ptr(), self.packed_ih_t.data_ptr(), self.input_size, 4,
                                          _stream()), "paule_tc_gemm_pack")

    def as_struct(self) -> _lib.LstmLayer:
        return _lib.LstmLayer(self.w_ih.data_ptr(), self.w_hh.data_ptr(), self.w_ih_t.data_ptr(),
                              self.w_hh_t.data_ptr(), self.bias.data_ptr(), _p(self.packed), _p(self.packed_ih),
                              _p(self.packed_ih_t), self.input_size)


# ------------------------------------------------------------------------------------------------
# custom ops
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("paule_b200::lstm_layer_fwd", mutates_args=())
@_guard
def lstm_layer_fwd(x: torch.Tensor, batch_first_in: bool, w_ih: torch.Tensor, w_hh: torch.Tensor,
                   bias: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """x [B,T,I] (batch_first_in) or [T,B,I] -> (h [T,B,H], gates [T,B,4H] activated, c [T,B,H])."""
    _chk(x, "x"); _chk(w_ih, "w_ih"); _chk(w_hh, "w_hh"); _chk(bias, "bias")
    if batch_first_in:
        B, T, I = x.shape
        a_map = (B, I, T * I)
    else:
        T, B, I = x.shape
        a_map = (1, I, 0)
    H = w_hh.shape[1]
    gates = torch.empty((T, B, 4 * H), device=x.device, dtype=torch.float32)
    h = torch.empty((T, B, H), device=x.device, dtype=torch.float32)
    c = torch.empty((T, B, H), device=x.device, dtype=torch.float32)
    if T * B == 0:
        return h, gates, c
    linear_rows_(gates, x, w_ih, bias, T * B, a_map, (1, 4 * H, 0))
    _lib.check(_lib.load().paule_lstm_seq_fwd_f32(gates.data_ptr(), w_hh.data_ptr(), h.data_ptr(), c.data_ptr(),
                                                  T, B, H, _stream()), "paule_lstm_seq_fwd_f32")
    return h, gates, c


@lstm_layer_fwd.register_fake
def _(x, batch_first_in, w_ih, w_hh, bias):
    if batch_first_in:
        B, T, _ = x.shape
    else:
        T, B, _ = x.shape
    H = w_hh.shape[1]
    return x.new_empty((T, B, H)), x.new_empty((T, B, 4 * H)), x.new_empty((T, B, H))


@torch.library.custom_op("paule_b200::lstm_layer_bwd", mutates_args=())
@_guard
def lstm_layer_bwd(dh: torch.Tensor, gates: torch.Tensor, c: torch.Tensor, w_ih_t: torch.Tensor,
                   w_hh_t: torch.Tensor, batch_first_out: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """Input-gradient BPTT: dh [T,B,H] -> (dx [T,B,I] (or [B,T,I] if batch_first_out), da [T,B,4H] = d loss / d pre-activations)."""
    _chk(dh, "dh"); _chk(gates, "gates"); _chk(c, "c")
    T, B, H = dh.shape
    I = w_ih_t.shape[0]
    da = gates.clone()                      # the kernel turns the stash into d(pre-activation) in place
    scratch = torch.empty((B, H), device=dh.device, dtype=torch.float32)
    dx = torch.empty((B, T, I) if batch_first_out else (T, B, I), device=dh.device, dtype=torch.float32)
    if T * B == 0:
        return dx, da
    _lib.check(_lib.load().paule_lstm_seq_bwd_f32(da.data_ptr(), c.data_ptr(), w_hh_t.data_ptr(), dh.data_ptr(), 1,
                                                  None, scratch.data_ptr(), T, B, H, _stream()),
               "paule_lstm_seq_bwd_f32")
    c_map = (B, I, T * I) if batch_first_out else (1, I, 0)
    linear_rows_(dx, da, w_ih_t, None, T * B, (1, 4 * H, 0), c_map)
    return dx, da


@lstm_layer_bwd.register_fake
def _(dh, gates, c, w_ih_t, w_hh_t, batch_first_out):
    T, B, _ = dh.shape
    I = w_ih_t.shape[0]
    return dh.new_empty((B, T, I) if batch_first_out else (T, B, I)), torch.empty_like(gates)


def _lstm_weight_grads(da: torch.Tensor, x: torch.Tensor, h: torch.Tensor, batch_first_in: bool):
    """dW_ih = dA^T x, dW_hh = dA[1:]^T h[:-1], db = column sums of dA over all (step, word) rows -- the library's own reduction
    kernels (the reference gets them from autograd through aten::lstm, paule/paule.py:1376)."""
    T, B, G = da.shape
    x_tm = (x.transpose(0, 1) if batch_first_in else x).reshape(T * B, -1).float().contiguous()
    da2 = da.reshape(T * B, G)
    d_w_ih = gemm_tn(da2, x_tm)
    if T > 1:
        d_w_hh = gemm_tn(da[1:].reshape((T - 1) * B, G), h[:-1].reshape((T - 1) * B, -1).contiguous())
    else:
        d_w_hh = torch.zeros((G, h.shape[-1]), device=da.device, dtype=torch.float32)
    return d_w_ih, d_w_hh, colsum(da2)


def _lstm_layer_setup(ctx, inputs, output):
    x, batch_first_in, w_ih, w_hh, bias = inputs
    h, gates, c = output
    ctx.weight_grads = any(ctx.needs_input_grad[2:5])
    if ctx.weight_grads:   # continue-learning (paule/paule.py:1372-1377): also keep the operands of the dW GEMMs
        ctx.save_for_backward(gates, c, w_ih, w_hh, x, h)
    else:
        ctx.save_for_backward(gates, c, w_ih, w_hh)
    ctx.batch_first_in = batch_first_in


def _lstm_layer_backward(ctx, dh, dgates, dc):
    gates, c, w_ih, w_hh = ctx.saved_tensors[:4]
    if dh is None:
        return None, None, None, None, None
    dx, da = lstm_layer_bwd(dh.contiguous(), gates, c, w_ih.t().contiguous(), w_hh.t().contiguous(), ctx.batch_first_in)
    if not ctx.weight_grads:
        return dx, None, None, None, None      # planning needs input gradients only
    # weight gradients: reductions over all (step, word) rows of the saved operands (paule_gemm_tn_f32 / paule_colsum_f32)
    x, h = ctx.saved_tensors[4:6]
    d_w_ih, d_w_hh, d_b = _lstm_weight_grads(da, x, h, ctx.batch_first_in)
    return dx, None, d_w_ih, d_w_hh, d_b


lstm_layer_fwd.register_autograd(_lstm_layer_backward, setup_context=_lstm_layer_setup)


# ---- the same layer on the persistent tcgen05 kernels (H = 720, bf16 operands / fp32 state): what ``module.math = MATH_BF16``
# selects in paule_b200.models.  ``packed`` is LstmWeights(tc=True).packed.
def _tc_scratch(lib, B: int, device) -> torch.Tensor:
    """exchange scratch of one launch: zero-filled (its status word is sticky and starts at 0)"""
    return torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), dtype=torch.uint8, device=device)


CLAMPED_MSG = ("a recurrent-path gradient of the bf16 BPTT kernel reached the +-1.5 bound of the in-band exchange (or was "
               "NaN / Inf) and was clamped: gradients deviate from the reference's; use math=MATH_FP32 for this loss scale")


def _tc_check(xchg: torch.Tensor) -> None:
    code, clamped = (int(v) for v in xchg[2048:2056].view(torch.int32).tolist())   # kXchgErrOff, kXchgClampOff: one 8-byte read
    if clamped:        # module-level autograd (continue-learning, user losses): a silently clamped gradient is an error
        raise _lib.PauleB200Error(CLAMPED_MSG)
    if code != 0:
        raise _lib.PauleB200Error(f"persistent recurrent kernel watchdog fired (status {code}): results are invalid")


@torch.library.custom_op("paule_b200::lstm_layer_fwd_tc", mutates_args=())
@_guard
def lstm_layer_fwd_tc(x: torch.Tensor, batch_first_in: bool, w_ih: torch.Tensor, w_hh: torch.Tensor, bias: torch.Tensor,
                      packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """lstm_layer_fwd on the tensor-core path: inputs of at most 64 channels (cps, mel, tube) go through the fused input
    projection (paule_tc_x_image + paule_tc_lstm_seq_fwd_x), wider ones through the fp32 projection + paule_tc_lstm_seq_fwd."""
    _chk(x, "x"); _chk(w_ih, "w_ih"); _chk(w_hh, "w_hh"); _chk(bias, "bias")
    lib = _lib.load()
    H = w_hh.shape[1]
    if batch_first_in:
        B, T, I = x.shape
    else:
        T, B, I = x.shape
    gates = torch.empty((T, B, 4 * H), device=x.device, dtype=torch.float32)
    h = torch.empty((T, B, H), device=x.device, dtype=torch.float32)
    c = torch.empty((T, B, H), device=x.device, dtype=torch.float32)
    if T * B == 0:
        return h, gates, c
    xchg = _tc_scratch(lib, B, x.device)
    st = _stream()
    if I <= 64:
        x_tm = transpose_btc(x) if batch_first_in else x
        ximg = torch.zeros(lib.paule_tc_x_image_bytes(T, B), dtype=torch.uint8, device=x.device)
        _lib.check(lib.paule_tc_x_image(x_tm.data_ptr(), ximg.data_ptr(), T, B, I, st), "paule_tc_x_image")
        _lib.check(lib.paule_tc_lstm_seq_fwd_x(gates.data_ptr(), packed.data_ptr(), bias.data_ptr(), ximg.data_ptr(), h.data_ptr(),
                                               c.data_ptr(), xchg.data_ptr(), None, T, B, MATH_BF16, st),
                   "paule_tc_lstm_seq_fwd_x")
    else:
        a_map = (B, I, T * I) if batch_first_in else (1, I, 0)
        linear_rows_(gates, x, w_ih, bias, T * B, a_map, (1, 4 * H, 0))
        _lib.check(lib.paule_tc_lstm_seq_fwd(gates.data_ptr(), packed.data_ptr(), h.data_ptr(), c.data_ptr(), xchg.data_ptr(),
                                             None, T, B, MATH_BF16, st), "paule_tc_lstm_seq_fwd")
    _tc_check(xchg)
    return h, gates, c


@lstm_layer_fwd_tc.register_fake
def _(x, batch_first_in, w_ih, w_hh, bias, packed):
    if batch_first_in:
        B, T, _ = x.shape
    else:
        T, B, _ = x.shape
    H = w_hh.shape[1]
    return x.new_empty((T, B, H)), x.new_empty((T, B, 4 * H)), x.new_empty((T, B, H))


@torch.library.custom_op("paule_b200::lstm_layer_bwd_tc", mutates_args=())
@_guard
def lstm_layer_bwd_tc(dh: torch.Tensor, gates: torch.Tensor, c: torch.Tensor, w_ih_t: torch.Tensor, packed: torch.Tensor,
                      batch_first_out: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """lstm_layer_bwd with the recurrence on the persistent tcgen05 BPTT kernel (paule_tc_lstm_seq_bwd)."""
    _chk(dh, "dh"); _chk(gates, "gates"); _chk(c, "c")
    lib = _lib.load()
    T, B, H = dh.shape
    I = w_ih_t.shape[0]
    da = gates.clone()
    dx = torch.empty((B, T, I) if batch_first_out else (T, B, I), device=dh.device, dtype=torch.float32)
    if T * B == 0:
        return dx, da
    xchg = _tc_scratch(lib, B, dh.device)
    _lib.check(lib.paule_tc_lstm_seq_bwd(da.data_ptr(), c.data_ptr(), packed.data_ptr(), dh.data_ptr(), 1, None, xchg.data_ptr(),
                                         None, T, B, MATH_BF16, _stream()), "paule_tc_lstm_seq_bwd")
    c_map = (B, I, T * I) if batch_first_out else (1, I, 0)
    linear_rows_(dx, da, w_ih_t, None, T * B, (1, 4 * H, 0), c_map)
    _tc_check(xchg)
    return dx, da


@lstm_layer_bwd_tc.register_fake
def _(dh, gates, c, w_ih_t, packed, batch_first_out):
    T, B, _ = dh.shape
    I = w_ih_t.shape[0]
    return dh.new_empty((B, T, I) if batch_first_out else (T, B, I)), torch.empty_like(gates)


def _lstm_layer_tc_setup(ctx, inputs, output):
    x, batch_first_in, w_ih, w_hh, bias, packed = inputs
    h, gates, c = output
    ctx.weight_grads = any(ctx.needs_input_grad[2:5])
    if ctx.weight_grads:
        ctx.save_for_backward(gates, c, w_ih, packed, x, h)
    else:
        ctx.save_for_backward(gates, c, w_ih, packed)
    ctx.batch_first_in = batch_first_in


def _lstm_layer_tc_backward(ctx, dh, dgates, dc):
    gates, c, w_ih, packed = ctx.saved_tensors[:4]
    if dh is None:
        return None, None, None, None, None, None
    dx, da = lstm_layer_bwd_tc(dh.contiguous(), gates, c, w_ih.t().contiguous(), packed, ctx.batch_first_in)
    if not ctx.weight_grads:
        return dx, None, None, None, None, None
    x, h = ctx.saved_tensors[4:6]      # weight gradients as in _lstm_layer_backward
    d_w_ih, d_w_hh, d_b = _lstm_weight_grads(da, x, h, ctx.batch_first_in)
    return dx, None, d_w_ih, d_w_hh, d_b, None


lstm_layer_fwd_tc.register_autograd(_lstm_layer_tc_backward, setup_context=_lstm_layer_tc_setup)


@torch.library.custom_op("paule_b200::linear_tm", mutates_args=())
@_guard
def linear_tm(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, pool_pairs: bool, batch_first_out: bool
              ) -> torch.Tensor:
    """x [T,B,K] time-major -> y = pool?(x) W^T + b as [T',B,N] or batch-first [B,T',N]; T' = T//2 if pool_pairs."""
    _chk(x, "x"); _chk(w, "w"); _chk(bias, "bias")
    T, B, K = x.shape
    N = w.shape[0]
    To = T // 2 if pool_pairs else T
    y = torch.empty((B, To, N) if batch_first_out else (To, B, N), device=x.device, dtype=torch.float32)
    if To * B == 0:
        return y
    a_map = (B, 2 * B * K, K) if pool_pairs else (1, K, 0)
    c_map = (B, N, To * N) if batch_first_out else (1, N, 0)
    linear_rows_(y, x, w, bias, To * B, a_map, c_map, a_pair=B * K if pool_pairs else 0)
    return y


@linear_tm.register_fake
def _(x, w, bias, pool_pairs, batch_first_out):
    T, B, _ = x.shape
    To = T // 2 if pool_pairs else T
    return x.new_empty((B, To, w.shape[0]) if batch_first_out else (To, B, w.shape[0]))


@torch.library.custom_op("paule_b200::linear_tm_bwd", mutates_args=())
@_guard
def linear_tm_bwd(dy: torch.Tensor, w_t: torch.Tensor, T: int, pool_pairs: bool, batch_first_out: bool
                  ) -> torch.Tensor:
    """Adjoint of linear_tm wrt x: dy ([T',B,N] or [B,T',N]) -> dx [T,B,K]."""
    _chk(dy, "dy"); _chk(w_t, "w_t")
    if batch_first_out:
        B, To, N = dy.shape
        a_map = (B, N, To * N)
    else:
        To, B, N = dy.shape
        a_map = (1, N, 0)
    K = w_t.shape[0]
    if not pool_pairs:
        dx = torch.empty((T, B, K), device=dy.device, dtype=torch.float32)
        if T * B:
            linear_rows_(dx, dy, w_t, None, T * B, a_map, (1, K, 0))
        return dx
    dx = torch.zeros((T, B, K), device=dy.device, dtype=torch.float32)
    if To * B:
        half = dy * 0.5
        # frame 2k and 2k+1 both receive 0.5 * dy[k] W
        linear_rows_(dx, half, w_t, None, To * B, a_map, (B, 2 * B * K, K))
        linear_rows_(dx[1:], half, w_t, None, To * B, a_map, (B, 2 * B * K, K))
    return dx


@linear_tm_bwd.register_fake
def _(dy, w_t, T, pool_pairs, batch_first_out):
    B = dy.shape[0] if batch_first_out else dy.shape[1]
    return dy.new_empty((T, B, w_t.shape[0]))


def _linear_tm_setup(ctx, inputs, output):
    x, w, bias, pool_pairs, batch_first_out = inputs
    ctx.weight_grads = any(ctx.needs_input_grad[1:3])
    if ctx.weight_grads:
        ctx.save_for_backward(w, x)
    else:
        ctx.save_for_backward(w)
    ctx.T = x.shape[0]
    ctx.pool_pairs = pool_pairs
    ctx.batch_first_out = batch_first_out


def _linear_tm_backward(ctx, dy):
    w = ctx.saved_tensors[0]
    dx = linear_tm_bwd(dy.contiguous(), w.t().contiguous(), ctx.T, ctx.pool_pairs, ctx.batch_first_out)
    if not ctx.weight_grads:
        return dx, None, None, None, None
    x = ctx.saved_tensors[1]                                  # [T,B,K] time-major
    dy_tm = dy.transpose(0, 1) if ctx.batch_first_out else dy     # [T',B,N]
    if ctx.pool_pairs:
        To = dy_tm.shape[0]
        x = 0.5 * (x[0:2 * To:2] + x[1:2 * To:2])
    N = dy_tm.shape[-1]
    dy2 = dy_tm.reshape(-1, N).float().contiguous()
    d_w = gemm_tn(dy2, x.reshape(-1, x.shape[-1]).float().contiguous())
    return dx, d_w, colsum(dy2), None, None


linear_tm.register_autograd(_linear_tm_backward, setup_context=_linear_tm_setup)


@torch.library.custom_op("paule_b200::plan_loss", mutates_args=())
@_guard
def plan_loss(mel: torch.Tensor, tmel: torch.Tensor, sv: torch.Tensor, tsv: torch.Tensor, cp: torch.Tensor,
              objective: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Time-major inputs (mel/tmel [Tm,B,60], sv/tsv [B,300], cp [T,B,30]) ->
    (terms [B,6] = total, mel, semvec, vel, jerk, local_linear; dmel; dsv; dcp_smooth)."""
    for n, t in (("mel", mel), ("tmel", tmel), ("sv", sv), ("tsv", tsv), ("cp", cp)):
        _chk(t, n)
    lib = _lib.load()
    Tm, B, Cm = mel.shape
    T, _, Cc = cp.shape
    S = sv.shape[1]
    terms = torch.empty((B, 6), device=cp.device, dtype=torch.float32)
    dmel, dsv, dcp = torch.empty_like(mel), torch.empty_like(sv), torch.empty_like(cp)
    scratch = torch.empty(lib.paule_plan_loss_scratch_floats(T, B), device=cp.device, dtype=torch.float32)
    _lib.check(lib.paule_plan_loss_f32(mel.data_ptr(), tmel.data_ptr(), sv.data_ptr(), tsv.data_ptr(), cp.data_ptr(),
                                       terms.data_ptr(), dmel.data_ptr(), dsv.data_ptr(), dcp.data_ptr(),
                                       scratch.data_ptr(), T, Tm, B, Cc, Cm, S, objective, _stream()),
               "paule_plan_loss_f32")
    return terms, dmel, dsv, dcp


@plan_loss.register_fake
def _(mel, tmel, sv, tsv, cp, objective):
    return cp.new_empty((mel.shape[1], 6)), torch.empty_like(mel), torch.empty_like(sv), torch.empty_like(cp)


@torch.library.custom_op("paule_b200::adam_clamp_", mutates_args=("cp", "m", "v", "step_count"))
@_guard
def adam_clamp_(cp: torch.Tensor, g_a: torch.Tensor, g_b: Optional[torch.Tensor], m: torch.Tensor, v: torch.Tensor,
                step_count: torch.Tensor, lr: float, beta1: float, beta2: float, eps: float, clamp: float,
                smiling: bool, past_cp: Optional[torch.Tensor]) -> None:
    """In-place Adam step + clamp on time-major cp [T,B,C]; advances the device step counter first."""
    for n, t in (("cp", cp), ("g_a", g_a), ("m", m), ("v", v)):
        _chk(t, n)
    _chk(step_count, "step_count", torch.int32)
    lib = _lib.load()
    T, B, Cc = cp.shape
    past_T = 0 if past_cp is None else past_cp.shape[0]
    _lib.check(lib.paule_step_tick(step_count.data_ptr(), _stream()), "paule_step_tick")
    _lib.check(lib.paule_adam_clamp_f32(cp.data_ptr(), g_a.data_ptr(), _p(g_b), m.data_ptr(), v.data_ptr(),
                                        step_count.data_ptr(), lr, beta1, beta2, eps, clamp, 1 if smiling else 0,
                                        _p(past_cp), past_T, T, B, Cc, _stream()), "paule_adam_clamp_f32")


# ------------------------------------------------------------------------------------------------
# the fused planner step
# ------------------------------------------------------------------------------------------------
# weak references: a planner that goes out of scope releases its workspace, Adam state, loss log and CUDA graph
_PLAN_REGISTRY: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()


def register_plan(ctx_obj) -> int:
    key = id(ctx_obj)
    _PLAN_REGISTRY[key] = ctx_obj
    return key


def unregister_plan(key: int) -> None:
    _PLAN_REGISTRY.pop(key, None)


@torch.library.custom_op("paule_b200::plan_step",
                         mutates_args=("cp", "adam_m", "adam_v", "step_count", "loss_log", "pred_mel", "pred_sv",
                                       "workspace"))
@_guard
def plan_step(cp: torch.Tensor, adam_m: torch.Tensor, adam_v: torch.Tensor, step_count: torch.Tensor,
              loss_log: torch.Tensor, pred_mel: torch.Tensor, pred_sv: torch.Tensor, workspace: torch.Tensor,
              plan_key: int) -> None:
    """One fused inner planning step (paule_plan_step) on the PlanContext registered under plan_key."""
    ctx_obj = _PLAN_REGISTRY[plan_key]
    _lib.check(_lib.load().paule_plan_step(ctx_obj.struct_ref(), _stream()), "paule_plan_step")


@torch.library.custom_op("paule_b200::plan_embed", mutates_args=("sv", "workspace"))
@_guard
def plan_embed(mel: torch.Tensor, sv: torch.Tensor, workspace: torch.Tensor, plan_key: int) -> None:
    """EmbeddingModel forward of the registered plan on a time-major mel [Tm,B,Cm] -> sv [B,S] (paule/paule.py:533-535)."""
    _chk(mel, "mel"); _chk(sv, "sv")
    ctx_obj = _PLAN_REGISTRY[plan_key]
    _lib.check(_lib.load().paule_plan_embed(ctx_obj.struct_ref(), mel.data_ptr(), sv.data_ptr(), _stream()),
               "paule_plan_embed")


@torch.library.custom_op("paule_b200::plan_forward", mutates_args=("pred_mel", "pred_sv", "workspace"))
@_guard
def plan_forward(cp: torch.Tensor, pred_mel: torch.Tensor, pred_sv: torch.Tensor, workspace: torch.Tensor,
                 plan_key: int) -> None:
    """no_grad forward of both models on the registered PlanContext (paule/paule.py:822-824, :1460-1464)."""
    ctx_obj = _PLAN_REGISTRY[plan_key]
    _lib.check(_lib.load().paule_plan_forward(ctx_obj.struct_ref(), _stream()), "paule_plan_forward")
