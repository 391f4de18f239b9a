"""Optional loss branches of the planner.

* speech classifier (reference: paule/paule.py:210-225, :915, criterion :603-622 / :664-682 / :719-737;
  ``LinearClassifier`` paule/models.py:887-911): fused into the criterion kernel -- this module only extracts
  the 60 weights + bias for ``paule_plan.cls_w / cls_b``.
* somatosensory feedback (reference: paule/paule.py:227-273, :916-931, criterion :624-645):
  ``pred_tube = cp_tube_model(cp)`` [T,B,10], ``tube_mel_model(pred_tube)`` against the target mel (weight 5),
  ``tube_embedder(pred_tube)`` against the target semvec (weight 10).  Three more LSTM models of the same op set; their
  forward, analytic backward and loss run here on the library's LSTM / Linear / criterion kernels (fp32: two of the models
  have 360 hidden units, which the tcgen05 recurrent kernels -- specialised for 720 -- do not cover), launch for launch on
  the planner's stream, and hand ``extra_terms [B,2]`` / ``extra_grad [T,B,30]`` to the fused step, which adds them to the
  logged total and to d(loss)/d(cp) before Adam.  No host synchronisation: the branch is captured into the same CUDA graph.
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
                 B: int, T: int, C: int, device):
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
        self.refresh_weights()

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

    # ------------------------------------------------------------------------------------------
    def forward(self, cp_tm: torch.Tensor, want_stash: bool = False):
        """cp [T,B,C] time-major -> (pred_tube [T,B,10], tube_mel [T//2,B,60], tube_semvec [B,300]) (+ the BPTT stash)."""
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

    def run(self, cp_tm: torch.Tensor, target_mel_tm: torch.Tensor, target_sv: torch.Tensor) -> None:
        """Evaluate the two tube terms and their gradient for the current cps into ``extra_terms`` / ``extra_grad``."""
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
