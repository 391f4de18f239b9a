"""Optional loss branches of the planner.

* speech classifier (reference: paule/paule.py:210-225, :915, criterion :603-622 / :664-682 / :719-737;
  ``LinearClassifier`` paule/models.py:887-911): fused into the criterion kernel -- this module only extracts
  the 60 weights + bias for ``paule_plan.cls_w / cls_b``.
* somatosensory feedback (reference: paule/paule.py:227-273, :916-931, criterion :624-645):
  ``pred_tube = cp_tube_model(cp)`` [T,B,10], ``tube_mel_model(pred_tube)`` against the target mel (weight 5),
  ``tube_embedder(pred_tube)`` against the target semvec (weight 10).  Three more LSTM models of the same op set; their
  forward, analytic backward and loss run here on the library's LSTM / Linear / criterion kernels, launch for launch on
  the planner's stream, and hand ``extra_terms [B,2]`` / ``extra_grad [T,B,30]`` to the fused step, which adds them to the
  logged total and to d(loss)/d(cp) before Adam.  No host synchronisation: the branch is captured into the same CUDA graph.

  Arithmetic follows the planner: fp32 step kernels (parity anchor), or -- ``math=MATH_BF16`` -- the persistent tcgen05
  recurrences.  Those are built for 720 hidden units; the two 360-unit models (paule/paule.py:231-249) run on them ZERO-PADDED
  to 720 units, which is exact: a unit whose weights and biases are all zero has pre-activations 0, hence i = f = o = 1/2,
  g = 0, c = 0 and h = 0 for ever, receives no gradient and passes none on (its W_hh column and post_linear column are zero).
  64 words x 200 frames on B200: 68.9 ms per inner step with the branch on the fp32 step kernels, see DESIGN.md for the
  tensor-core figure.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, ops
from .models import EmbeddingModel, ForwardModel, _f32c


def classifier_operands(speech_classifier, Cm: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """(w [Cm], b [1]) fp32 on the device from a ``LinearClassifier``-like module (``.linear`` = Linear(Cm, 1))."""
    lin = getattr(speech_classifier, "linear", None)
    if lin is None or tuple(lin.weight.shape) != (1, Cm):
        raise NotImplementedError("the fused planner implements the reference's LinearClassifier(input_dim=60, output_dim=1) "
                                  "(paule/paule.py:215, paule/models.py:887-911); other classifiers are not supported")
    w = lin.weight.detach().to(device).float().reshape(Cm).contiguous()
    b = (lin.bias.detach().to(device).float() if lin.bias is not None else torch.zeros(1, device=device)).reshape(1).contiguous()
    return w, b


class SomatosensoryBranch:
    """cp -> tube -> (mel, semvec) loss branch on B words; buffers are static so the launches can be graph-captured."""

    def __init__(self, cp_tube_model: ForwardModel, tube_mel_model: ForwardModel, tube_embedder: EmbeddingModel,
                 B: int, T: int, C: int, device, math: int = ops.MATH_FP32):
        for name, m in (("cp_tube_model", cp_tube_model), ("tube_mel_model", tube_mel_model)):
            if not hasattr(m, "lstm") or m.lstm.num_layers != 1:
                raise NotImplementedError(f"{name}: expected a 1-layer ForwardModel (paule/paule.py:231-249)")
        if cp_tube_model.apply_half_sequence or not tube_mel_model.apply_half_sequence:
            raise NotImplementedError("cp_tube_model keeps the cp frame rate, tube_mel_model halves it (paule/paule.py:235,247)")
        if tube_embedder.lstm.num_layers != 2 or tube_embedder.post_upsampling_size != 0:
            raise NotImplementedError("tube_embedder: expected the 2-layer EmbeddingModel of paule/paule.py:258-262")
        self._models = (cp_tube_model, tube_mel_model, tube_embedder)
        self.B, self.T, self.C = B, T, C
        self.device = device
        self.dropout = float(tube_embedder.lstm.dropout)
        f32 = dict(device=device, dtype=torch.float32)
        self.extra_terms = torch.zeros((B, 2), **f32)
        self.extra_grad = torch.zeros((T, B, C), **f32)
        self.pred_tube = None
        # tensor-core path: every model as a 720-unit layer (the 360-unit ones zero-padded), static buffers, no host checks
        # inside the loop (the status words are read by BatchPlanner.check()).  The stochastic variant (active inter-layer
        # dropout of the shipped tube embedder) multiplies the fp32 h_0, which this path never materialises: fp32 kernels.
        hs = [m.lstm.hidden_size for m in self._models]
        self.tc = (math != ops.MATH_FP32 and ops.tc_available() and not self.stochastic and tube_embedder.lstm.hidden_size == 720
                   and all(h <= 720 for h in hs) and cp_tube_model.lstm.input_size <= 64 and tube_mel_model.lstm.input_size <= 64
                   and tube_embedder.lstm.input_size <= 64 and T % 2 == 0)
        self.refresh_weights()
        if self.tc:
            self._alloc_tc()

    @property
    def stochastic(self) -> bool:
        """The reference puts the tube embedder in training mode inside the loop (paule/paule.py:928): with its shipped
        dropout (0.7, :261) the inter-layer dropout is active while planning."""
        return self.dropout > 0.0

    def refresh_weights(self) -> None:
        def layer(lstm, k):
            return ops.LstmWeights(getattr(lstm, f"weight_ih_l{k}"), getattr(lstm, f"weight_hh_l{k}"),
                                   getattr(lstm, f"bias_ih_l{k}"), getattr(lstm, f"bias_hh_l{k}"))
        ct, tm, te = self._models
        self.w_ct, self.w_tm = layer(ct.lstm, 0), layer(tm.lstm, 0)
        self.w_e0, self.w_e1 = layer(te.lstm, 0), layer(te.lstm, 1)
        self.ct_w, self.ct_b = _f32c(ct.post_linear.weight), _f32c(ct.post_linear.bias)
        self.tm_w, self.tm_b = _f32c(tm.post_linear.weight), _f32c(tm.post_linear.bias)
        self.head_w, self.head_b = _f32c(te.linear_mapping.weight), _f32c(te.linear_mapping.bias)
        self.ct_w_t, self.tm_w_t, self.head_w_t = (w.t().contiguous() for w in (self.ct_w, self.tm_w, self.head_w))
        if getattr(self, "tc", False):
            H = 720

            def padded(lstm, k):      # as a 720-unit layer (ops.pad_lstm_params; exact, see the module docstring)
                return ops.LstmWeights(*ops.pad_lstm_params(getattr(lstm, f"weight_ih_l{k}"), getattr(lstm, f"weight_hh_l{k}"),
                                                            getattr(lstm, f"bias_ih_l{k}"), getattr(lstm, f"bias_hh_l{k}"),
                                                            input_padded=k > 0 and lstm.hidden_size < H), tc=True)

            def pad_cols(w):          # post_linear [out, h] -> [out, 720]
                out = w.new_zeros(w.shape[0], H)
                out[:, :w.shape[1]] = w
                return out.contiguous()
            self.p_ct, self.p_tm = padded(ct.lstm, 0), padded(tm.lstm, 0)
            self.p_e0, self.p_e1 = padded(te.lstm, 0), padded(te.lstm, 1)
            self.ct_wp, self.tm_wp = pad_cols(self.ct_w), pad_cols(self.tm_w)
            self.ct_wp_t, self.tm_wp_t = self.ct_wp.t().contiguous(), self.tm_wp.t().contiguous()

    # ------------------------------------------------------------------------------------------
    def forward(self, cp_tm: torch.Tensor, want_stash: bool = False):
        """cp [T,B,C] time-major -> (pred_tube [T,B,10], tube_mel [T//2,B,60], tube_semvec [B,300]) (+ the BPTT stash)."""
        if self.tc and not want_stash:
            tube, tube_mel, sv = self._forward_tc(cp_tm.contiguous())
            return tube.clone(), tube_mel.clone(), sv.clone()
        h_t, g_t, c_t = ops.lstm_layer_fwd(cp_tm, False, self.w_ct.w_ih, self.w_ct.w_hh, self.w_ct.bias)    # paule.py:917
        tube = ops.linear_tm(h_t, self.ct_w, self.ct_b, False, False)
        h_m, g_m, c_m = ops.lstm_layer_fwd(tube, False, self.w_tm.w_ih, self.w_tm.w_hh, self.w_tm.bias)     # :919
        tube_mel = ops.linear_tm(h_m, self.tm_w, self.tm_b, True, False)
        h_0, g_0, c_0 = ops.lstm_layer_fwd(tube, False, self.w_e0.w_ih, self.w_e0.w_hh, self.w_e0.bias)     # :929
        mask = None
        if want_stash and self.stochastic:       # nn.LSTM inter-layer dropout in training mode (:928)
            mask = (torch.rand_like(h_0) >= self.dropout).float() / (1.0 - self.dropout)
            h_0 = h_0 * mask
        h_1, g_1, c_1 = ops.lstm_layer_fwd(h_0, False, self.w_e1.w_ih, self.w_e1.w_hh, self.w_e1.bias)
        sv = ops.linear_tm(h_1[-1:].contiguous(), self.head_w, self.head_b, False, False)[0]
        if not want_stash:
            return tube, tube_mel, sv
        return tube, tube_mel, sv, (g_t, c_t, g_m, c_m, g_0, c_0, g_1, c_1, mask)

    # ------------------------------------------------------------------------------------------ tensor-core path
    def _alloc_tc(self) -> None:
        lib = _lib.load()
        T, B, H, dev = self.T, self.B, 720, self.device
        f32 = dict(device=dev, dtype=torch.float32)
        u8 = dict(device=dev, dtype=torch.uint8)
        self._stash = {k: (torch.empty((T, B, 4 * H), **f32), torch.empty((T, B, H), **f32)) for k in ("ct", "tm", "e0", "e1")}
        self._h = {k: torch.empty((T, B, H), **f32) for k in ("ct", "tm", "e1")}
        self._xchg = torch.zeros(lib.paule_tc_rnn_xchg_bytes(B), **u8)             # status word starts at 0, sticky
        self._ximg_cp = torch.zeros(lib.paule_tc_x_image_bytes(T, B), **u8)
        self._ximg_tube = torch.zeros(lib.paule_tc_x_image_bytes(T, B), **u8)
        self._h0_img = torch.zeros(lib.paule_tc_img_seq_bytes(T, B, 1), **u8)
        self._da_img = torch.zeros(lib.paule_tc_img_seq_bytes(T, B, 4), **u8)
        self._tube = torch.empty((T, B, self.ct_w.shape[0]), **f32)
        self._tube_mel = torch.empty((T // 2, B, self.tm_w.shape[0]), **f32)
        self._sv = torch.empty((B, self.head_w.shape[0]), **f32)
        self._dh = torch.empty((T, B, H), **f32)
        self._dh_last = torch.empty((B, H), **f32)
        self._dtube = torch.empty((T, B, self.ct_w.shape[0]), **f32)

    def status_words(self):
        """(status, clamp notice) of the branch's persistent kernels, or None on the fp32 path (BatchPlanner.check reads them)."""
        if not self.tc:
            return None
        return tuple(int(v) for v in self._xchg[2048:2056].view(torch.int32).tolist())

    def _fwd_x(self, L, ximg, key, h, img=None):
        lib, st = _lib.load(), ops._stream()
        gates, c = self._stash[key]
        _lib.check(lib.paule_tc_lstm_seq_fwd_x(gates.data_ptr(), L.packed.data_ptr(), L.bias.data_ptr(), ximg.data_ptr(),
                                               None if h is None else h.data_ptr(), c.data_ptr(), self._xchg.data_ptr(),
                                               None if img is None else img.data_ptr(), self.T, self.B, ops.MATH_BF16, st),
                   "paule_tc_lstm_seq_fwd_x")

    def _bwd(self, L, key, dh_seq, dh_last, dx, accumulate=False):
        """BPTT of one layer on the tcgen05 kernel (dA as bf16 images) + dX = dA W_ih on the tcgen05 GEMM."""
        lib, st = _lib.load(), ops._stream()
        gates, c = self._stash[key]
        _lib.check(lib.paule_tc_lstm_seq_bwd_img(gates.data_ptr(), c.data_ptr(), L.packed.data_ptr(),
                                                 None if dh_seq is None else dh_seq.data_ptr(), 0 if dh_seq is None else 1,
                                                 None if dh_last is None else dh_last.data_ptr(), self._xchg.data_ptr(),
                                                 self._da_img.data_ptr(), self.T, self.B, ops.MATH_BF16, st),
                   "paule_tc_lstm_seq_bwd_img")
        _lib.check(lib.paule_tc_gemm_img(self._da_img.data_ptr(), L.packed_ih_t.data_ptr(), None, dx.data_ptr(), self.T, self.B,
                                         L.input_size, 4, 1 if accumulate else 0, st), "paule_tc_gemm_img")

    def _forward_tc(self, cp_tm: torch.Tensor):
        lib, st = _lib.load(), ops._stream()
        T, B = self.T, self.B
        _lib.check(lib.paule_tc_x_image(cp_tm.data_ptr(), self._ximg_cp.data_ptr(), T, B, cp_tm.shape[2], st), "paule_tc_x_image")
        self._fwd_x(self.p_ct, self._ximg_cp, "ct", self._h["ct"])                                        # paule.py:917
        ops.linear_rows_(self._tube, self._h["ct"], self.ct_wp, self.ct_b, T * B, (1, 720, 0), (1, self._tube.shape[2], 0))
        _lib.check(lib.paule_tc_x_image(self._tube.data_ptr(), self._ximg_tube.data_ptr(), T, B, self._tube.shape[2], st),
                   "paule_tc_x_image")
        self._fwd_x(self.p_tm, self._ximg_tube, "tm", self._h["tm"])                                      # :919
        Cm = self._tube_mel.shape[2]
        ops.linear_rows_(self._tube_mel, self._h["tm"], self.tm_wp, self.tm_b, (T // 2) * B, (B, 2 * B * 720, 720), (1, Cm, 0),
                         a_pair=B * 720)                                                                  # post_linear + pool
        self._fwd_x(self.p_e0, self._ximg_tube, "e0", None, self._h0_img)                                 # :929, layer 0
        g1, c1 = self._stash["e1"]
        _lib.check(lib.paule_tc_gemm_img(self._h0_img.data_ptr(), self.p_e1.packed_ih.data_ptr(), self.p_e1.bias.data_ptr(),
                                         g1.data_ptr(), T, B, 4 * 720, 1, 0, st), "paule_tc_gemm_img")
        _lib.check(lib.paule_tc_lstm_seq_fwd(g1.data_ptr(), self.p_e1.packed.data_ptr(), self._h["e1"].data_ptr(), c1.data_ptr(),
                                             self._xchg.data_ptr(), None, T, B, ops.MATH_BF16, st), "paule_tc_lstm_seq_fwd")
        S = self._sv.shape[1]
        ops.linear_rows_(self._sv, self._h["e1"], self.head_w, self.head_b, B, (1, 720, 0), (1, S, 0), a_offset=(T - 1) * B * 720)
        return self._tube, self._tube_mel, self._sv

    def _run_tc(self, cp_tm, target_mel_tm, target_sv) -> None:
        T, B = self.T, self.B
        tube, tube_mel, sv = self._forward_tc(cp_tm)
        terms, dmel, dsv, _ = ops.plan_loss(tube_mel, target_mel_tm, sv, target_sv, cp_tm, ops.OBJECTIVES["acoustic_semvec"])
        self.extra_terms.copy_(terms[:, 1:3])
        # embedder: head^T -> layer 1 (gradient on the last step only) -> layer 0 -> d(tube)
        ops.linear_rows_(self._dh_last, dsv, self.head_w_t, None, B, (1, dsv.shape[1], 0), (1, 720, 0))
        self._bwd(self.p_e1, "e1", None, self._dh_last, self._dh)                          # dh0 [T,B,720]
        self._bwd(self.p_e0, "e0", self._dh, None, self._dtube)
        # tube -> mel model: post_linear^T with the un-pooling, BPTT, += d(tube)
        dh_m = ops.linear_tm_bwd(dmel, self.tm_wp_t, T, True, False)
        self._bwd(self.p_tm, "tm", dh_m, None, self._dtube, accumulate=True)
        # cp -> tube model
        dh_t = ops.linear_tm_bwd(self._dtube, self.ct_wp_t, T, False, False)
        self._bwd(self.p_ct, "ct", dh_t, None, self.extra_grad)

    def run(self, cp_tm: torch.Tensor, target_mel_tm: torch.Tensor, target_sv: torch.Tensor) -> None:
        """Evaluate the two tube terms and their gradient for the current cps into ``extra_terms`` / ``extra_grad``."""
        if self.tc:
            return self._run_tc(cp_tm, target_mel_tm, target_sv)
        T, B = self.T, self.B
        tube, tube_mel, sv, (g_t, c_t, g_m, c_m, g_0, c_0, g_1, c_1, mask) = self.forward(cp_tm, want_stash=True)
        # 5 rmse(tube_mel, target_mel), 10 rmse(tube_semvec, target_semvec) and their gradients: the criterion kernel with the
        # acoustic_semvec objective (TUBE_MEL_WEIGHT = MEL_WEIGHT, TUBE_SEMANTIC_WEIGHT = SEMANTIC_WEIGHT, paule.py:598-599)
        terms, dmel, dsv, _ = ops.plan_loss(tube_mel, target_mel_tm, sv, target_sv, cp_tm, ops.OBJECTIVES["acoustic_semvec"])
        self.extra_terms.copy_(terms[:, 1:3])
        # backward: embedder head -> l1 -> (dropout) -> l0 -> d(tube); tube_mel post_linear^T (un-pool) -> LSTM -> d(tube)
        dh1 = torch.zeros((T, B, self.w_e1.hidden), device=cp_tm.device, dtype=torch.float32)
        dh1[-1] = ops.linear_tm_bwd(dsv.unsqueeze(0).contiguous(), self.head_w_t, 1, False, False)[0]
        dh0, _ = ops.lstm_layer_bwd(dh1, g_1, c_1, self.w_e1.w_ih_t, self.w_e1.w_hh_t, False)
        if mask is not None:
            dh0 = dh0 * mask
        dtube, _ = ops.lstm_layer_bwd(dh0, g_0, c_0, self.w_e0.w_ih_t, self.w_e0.w_hh_t, False)
        dh_m = ops.linear_tm_bwd(dmel, self.tm_w_t, T, True, False)
        dtube_m, _ = ops.lstm_layer_bwd(dh_m, g_m, c_m, self.w_tm.w_ih_t, self.w_tm.w_hh_t, False)
        dtube = dtube + dtube_m
        dh_t = ops.linear_tm_bwd(dtube, self.ct_w_t, T, False, False)
        dcp, _ = ops.lstm_layer_bwd(dh_t, g_t, c_t, self.w_ct.w_ih_t, self.w_ct.w_hh_t, False)
        self.extra_grad.copy_(dcp)
