"""Batched, device-resident driver of the planning inner loop (the new component SURVEY.md section 8
asks for: the reference plans one word per call, paule/paule.py:539,826; words are independent, so a
batch is B independent plans run in lock-step with per-word losses).

``BatchPlanner`` owns (through PyTorch) every device buffer of one planning job -- cps, Adam state,
targets, activation stash -- laid out time-major, and drives ``paule_plan_step`` (one C call per inner
step, or one CUDA-graph replay).  No host synchronisation happens inside the loop: the loss terms of
every step go to a device ring buffer that is copied out once at the end.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import Dict, Optional

import torch

from . import _lib, ops
from .models import EmbeddingModel, ForwardModel, _f32c


class BatchPlanner:
    """Plans B words at once on one GPU.

    Parameters mirror what ``Paule.plan_resynth`` sets up before its loop (paule/paule.py:585-597,797):
    ``initial_cp`` [B,T,30], ``target_mel`` [B,T//2,60], ``target_semvec`` [B,300] or None (then it is the
    embedder's output on the target mel, :533-535), Adam(lr) with torch defaults, clamp 1.05.

    ``lengths`` (optional, B ints): cp frames of every word of a RAGGED batch.  ``initial_cp`` and
    ``target_mel`` are then padded to the longest word; word b is planned exactly as if it were alone (the reference plans
    one word per call): its mel / semvec / smoothness terms are means over its own ``lengths[b]`` (``lengths[b] // 2`` mel)
    frames, its semvec is read at its own last mel frame, and the padding frames receive a zero gradient.

    Optional loss branches (mutually exclusive as in paule/paule.py:117-118): ``speech_classifier`` -- a
    ``LinearClassifier`` whose term 0.1 BCEWithLogits(classifier(pred_mel), 0) is fused into the criterion kernel;
    ``somatosensory`` -- ``(cp_tube_model, tube_mel_model, tube_embedder)``, see ``branches.SomatosensoryBranch``.

    ``math``: ``ops.MATH_FP32`` (FFMA kernels, the parity anchor) or ``ops.MATH_BF16`` (persistent tcgen05 recurrences + tcgen05
    GEMMs: bf16 operands, fp32 accumulation, state, loss and Adam); ``None`` = ``ops.default_math(hidden_size)``, i.e. the
    tensor-core path for Paule's 720-unit models.

    The embedder is evaluated deterministically (eval-mode semantics): the reference switches it to ``train()`` inside the
    loop (paule/paule.py:923), which only matters for an embedder with LSTM dropout > 0 -- such an embedder is rejected here
    (Paule's embedder has dropout = 0, paule/paule.py:167).

    A planner can be re-armed for another job of the same shape with ``reset()``: workspace, packed weights and the captured
    CUDA graph are kept (this is what ``Paule.plan_resynth`` does when it is called in a loop)."""

    def __init__(self, pred_model: ForwardModel, embedder: EmbeddingModel, initial_cp: torch.Tensor,
                 target_mel: torch.Tensor, target_semvec: Optional[torch.Tensor] = None, *, lr: float = 0.01,
                 objective: str = "acoustic_semvec", smiling: bool = False, past_cp: Optional[torch.Tensor] = None,
                 log_semantics: bool = True, log_gradients: bool = False, max_log_steps: int = 1024,
                 math: Optional[int] = None, use_cuda_graph: bool = True, lengths=None,
                 speech_classifier=None, somatosensory=None):
        _lib.require_device()
        if objective not in ops.OBJECTIVES:
            raise ValueError("objective has to be one of 'acoustic_semvec', 'acoustic' or 'semvec'")
        if pred_model.lstm.num_layers != 1 or embedder.lstm.num_layers != 2 or embedder.post_upsampling_size != 0:
            raise NotImplementedError("the fused planner implements Paule's configuration: 1-layer ForwardModel, "
                                      "2-layer EmbeddingModel without post-upsampling (paule/paule.py:124,167)")
        if not pred_model.apply_half_sequence:
            raise NotImplementedError("the fused planner expects apply_half_sequence=True")
        if getattr(embedder.lstm, "dropout", 0) > 0:
            raise NotImplementedError("the fused planner evaluates the embedder without inter-layer dropout; the reference plans "
                                      "with embedder.train() (paule/paule.py:923), so an embedder with LSTM dropout > 0 would "
                                      "be planned differently -- Paule's embedder has dropout=0 (paule/paule.py:167)")
        dev = initial_cp.device
        if dev.type != "cuda":
            raise _lib.PauleB200Error("BatchPlanner needs CUDA tensors: there is no CPU fallback")
        self.device = dev
        self._guard = torch.cuda.device(dev)      # the C side launches on the current device (ops._on)
        self._guard.__enter__()
        try:
            self._init(pred_model, embedder, initial_cp, target_mel, target_semvec, lr, objective, smiling, past_cp,
                       log_semantics, log_gradients, max_log_steps, math, use_cuda_graph, lengths, speech_classifier,
                       somatosensory)
        finally:
            self._guard.__exit__(None, None, None)

    def _init(self, pred_model, embedder, initial_cp, target_mel, target_semvec, lr, objective, smiling, past_cp,
              log_semantics, log_gradients, max_log_steps, math, use_cuda_graph, lengths, speech_classifier, somatosensory):
        dev = self.device
        B, T, C = initial_cp.shape
        Tm = T // 2
        if target_mel.shape[0] != B or target_mel.shape[1] != Tm:
            raise ValueError(f"initial_cp {tuple(initial_cp.shape)} does not match target_mel {tuple(target_mel.shape)}")
        if T < 13:
            raise ValueError("cp trajectories need at least 13 frames (three nested 5-point stencils, util.py:634-636)")
        H = pred_model.lstm.hidden_size
        if embedder.lstm.hidden_size != H:
            raise NotImplementedError("pred_model and embedder must share the hidden size")
        Cm, S = target_mel.shape[2], embedder.linear_mapping.out_features
        self.B, self.T, self.Tm, self.H, self.C, self.Cm, self.S = B, T, Tm, H, C, Cm, S
        if math is None:
            math = ops.default_math(H)
        # tensor-core math with models narrower than 720 units: every layer zero-padded to 720 units (exact)
        self._pad = math != ops.MATH_FP32 and H < ops.TC_HIDDEN
        if self._pad:
            H = ops.TC_HIDDEN
            self.H = H
        self.objective, self.math = objective, math
        self.lengths, self.word_frames = None, None
        if lengths is not None:
            ln = [int(v) for v in lengths]
            if len(ln) != B:
                raise ValueError(f"lengths has {len(ln)} entries for {B} words")
            if min(ln) < 13 or max(ln) > T:
                raise ValueError("every word needs 13 <= lengths[b] <= initial_cp.shape[1] cp frames")
            if past_cp is not None:
                raise NotImplementedError("past_cp with ragged batches")
            if any(v != T for v in ln):
                self.lengths = ln
                self.word_frames = torch.tensor(ln, device=dev, dtype=torch.int32)
        tc = math != ops.MATH_FP32
        # ---- weights (replicated per GPU; repacked by refresh_weights() after continue-learning)
        self._pred, self._emb = pred_model, embedder
        self._tc = tc
        self.refresh_weights()
        # ---- state, time-major
        f32 = dict(device=dev, dtype=torch.float32)
        self.cp = ops.transpose_btc(initial_cp.float().contiguous())                  # [T,B,C]
        self.adam_m = torch.zeros_like(self.cp)
        self.adam_v = torch.zeros_like(self.cp)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int32)
        self.target_mel = ops.transpose_btc(target_mel.float().contiguous())          # [Tm,B,Cm]
        self.past_cp = None
        if past_cp is not None:
            pc = past_cp.float().contiguous()
            if pc.dim() == 2:
                pc = pc.unsqueeze(0).expand(B, -1, -1).contiguous()
            self.past_cp = ops.transpose_btc(pc)
        self.max_log_steps = int(max_log_steps)
        self.loss_log = torch.zeros((self.max_log_steps, B, 6), **f32)
        self.pred_mel = torch.empty((Tm, B, Cm), **f32)
        self.pred_sv = torch.zeros((B, S), **f32)
        self.grad_out = torch.empty_like(self.cp) if log_gradients else None
        # ---- optional loss branches
        if speech_classifier is not None and somatosensory is not None:
            raise NotImplementedError("at the moment you have to choose either to use `use_somatosenrosry_feedback=True` OR to use `use_speech_classifier=True` or none")
        self.cls_w = self.cls_b = self.soma = self.aux_log = None
        if speech_classifier is not None:
            from .branches import classifier_operands
            self.cls_w, self.cls_b = classifier_operands(speech_classifier, Cm, dev)
        if somatosensory is not None:
            if objective != "acoustic_semvec":
                # the reference's 'acoustic' / 'semvec' somatosensory criteria use undefined names (paule/paule.py:697,:750)
                raise NotImplementedError("somatosensory feedback is defined for objective='acoustic_semvec' only "
                                          "(paule/paule.py:624-645)")
            if self.lengths is not None:
                raise NotImplementedError("somatosensory feedback with ragged batches")
            from .branches import SomatosensoryBranch
            self.soma = SomatosensoryBranch(*somatosensory, B=B, T=T, C=C, device=dev, math=math)
        if self.cls_w is not None or self.soma is not None:
            self.aux_log = torch.zeros((self.max_log_steps, B, 3), **f32)
        lib = _lib.load()
        ws_bytes = lib.paule_plan_workspace_bytes(B, T, H, C, Cm, S, math)
        # zero-filled ONCE: the pad rows / pad columns of the bf16 operand images inside must be exact zeros
        self.workspace = torch.zeros(ws_bytes, device=dev, dtype=torch.uint8)
        off = lib.paule_plan_status_offset(B, T, H, C, Cm, S, math)
        self._status_off = None if off >= ws_bytes else int(off)
        self._grad_lstm_off = int(lib.paule_plan_grad_lstm_offset(B, T, H, C, Cm, S, math))
        self._warned_clamp = False
        self.target_sv = torch.zeros((B, S), **f32)
        self.hp = dict(lr=float(lr), beta1=0.9, beta2=0.999, eps=1e-8, clamp=1.05)
        self.smiling, self.log_semantics = bool(smiling), bool(log_semantics)
        self._struct = None
        self._build_struct()
        self._key = ops.register_plan(self)
        self._graph = None
        self._use_graph = bool(use_cuda_graph)
        self.steps_done = 0
        # target semvec = embedder(target_mel) under no_grad (paule/paule.py:533-535) unless given
        if target_semvec is None:
            self.target_sv.copy_(self.embed(self.target_mel))
        else:
            self.target_sv.copy_(target_semvec.to(dev).float().reshape(B, S))

    # ------------------------------------------------------------------------------------------
    def refresh_weights(self) -> None:
        """(Re)pack the model weights; call after the models' parameters changed."""
        p, e = self._pred, self._emb

        def layer(lstm, k):
            q = [getattr(lstm, f"{n}_l{k}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
            if self._pad:      # narrower than the tensor-core kernels' 720 units: zero-padded (exact, ops.pad_lstm_params)
                q = ops.pad_lstm_params(*q, input_padded=(lstm is e.lstm and k == 1))
            return ops.LstmWeights(*q, tc=self._tc)

        def cols(w):           # Linear weight [out, h] reading a padded h: zero columns behind the model's own units
            w = _f32c(w)
            if not self._pad or w.shape[1] == self.H:
                return w
            out = w.new_zeros(w.shape[0], self.H)
            out[:, :w.shape[1]] = w
            return out.contiguous()
        self.w_fwd, self.w_e0, self.w_e1 = layer(p.lstm, 0), layer(e.lstm, 0), layer(e.lstm, 1)
        self.post_w, self.post_b = cols(p.post_linear.weight), _f32c(p.post_linear.bias)
        self.post_w_t = self.post_w.t().contiguous()
        self.post_packed = None
        if self._tc:
            # AvgPool1d(2,2) o Linear == one GEMM over K = 2 x 720 with the weights [0.5 W | 0.5 W]: the two frames of
            # a pair are the two K segments of one row, so the pooled projection runs on the h_t images unchanged
            lib = _lib.load()
            w2 = torch.cat((0.5 * self.post_w, 0.5 * self.post_w), dim=1).contiguous()
            n_out = w2.shape[0]
            self.post_packed = torch.empty(lib.paule_tc_gemm_packed_bytes(n_out, 2), dtype=torch.uint8, device=w2.device)
            _lib.check(lib.paule_tc_gemm_pack(w2.data_ptr(), self.post_packed.data_ptr(), n_out, 2, ops._stream()),
                       "paule_tc_gemm_pack")
        self.bwd_fused_packed = None
        if self._tc:
            # backward layer wavefront: d/d(pooled forward-model h) = dA_0 (W_ih0 W_post) in one streaming GEMM -- the weights
            # as the GEMM wants them: [N = H, K = 4H] = post_linear.weight^T x weight_ih_l0^T
            lib = _lib.load()
            wc = (self.post_w.t() @ self.w_e0.w_ih.t()).contiguous()
            self.bwd_fused_packed = torch.empty(lib.paule_tc_gemm_packed_bytes(wc.shape[0], 4), dtype=torch.uint8, device=wc.device)
            _lib.check(lib.paule_tc_gemm_pack(wc.data_ptr(), self.bwd_fused_packed.data_ptr(), wc.shape[0], 4, ops._stream()),
                       "paule_tc_gemm_pack")
        self.post_t_packed = None
        if self._tc and self.post_w_t.shape[1] <= 64 and self.post_w_t.shape[1] % 2 == 0:
            # serial backward: d/d(pooled forward-model h) = dmel W_post on the tcgen05 GEMM (K = 60: one k-block)
            lib = _lib.load()
            n_h, k_m = self.post_w_t.shape
            self.post_t_packed = torch.empty(lib.paule_tc_gemm_packed_bytes_k64(n_h), dtype=torch.uint8, device=self.post_w_t.device)
            _lib.check(lib.paule_tc_gemm_pack_k64(self.post_w_t.data_ptr(), self.post_t_packed.data_ptr(), n_h, k_m, ops._stream()),
                       "paule_tc_gemm_pack_k64")
        self.head_w, self.head_b = cols(e.linear_mapping.weight), _f32c(e.linear_mapping.bias)
        self.head_w_t = self.head_w.t().contiguous()
        if getattr(self, "_struct", None) is not None:
            self._build_struct()
            self._graph = None

    def _build_struct(self) -> None:
        s = _lib.Plan()
        s.B, s.T, s.H, s.C, s.Cm, s.S = self.B, self.T, self.H, self.C, self.Cm, self.S
        s.objective, s.math = ops.OBJECTIVES[self.objective], self.math
        s.smiling, s.log_slot_count = int(self.smiling), self.max_log_steps
        s.log_semantics = int(self.log_semantics)
        s.fwd, s.emb0, s.emb1 = self.w_fwd.as_struct(), self.w_e0.as_struct(), self.w_e1.as_struct()
        s.post_w, s.post_w_t, s.post_b = self.post_w.data_ptr(), self.post_w_t.data_ptr(), self.post_b.data_ptr()
        s.post_packed = None if self.post_packed is None else self.post_packed.data_ptr()
        s.head_w, s.head_w_t, s.head_b = self.head_w.data_ptr(), self.head_w_t.data_ptr(), self.head_b.data_ptr()
        s.cp, s.adam_m, s.adam_v = self.cp.data_ptr(), self.adam_m.data_ptr(), self.adam_v.data_ptr()
        s.step_count = self.step_count.data_ptr()
        s.target_mel, s.target_sv = self.target_mel.data_ptr(), self.target_sv.data_ptr()
        s.past_cp = None if self.past_cp is None else self.past_cp.data_ptr()
        s.past_T = 0 if self.past_cp is None else self.past_cp.shape[0]
        s.lr, s.beta1, s.beta2 = self.hp["lr"], self.hp["beta1"], self.hp["beta2"]
        s.eps, s.clamp = self.hp["eps"], self.hp["clamp"]
        s.loss_log, s.pred_mel, s.pred_sv = self.loss_log.data_ptr(), self.pred_mel.data_ptr(), self.pred_sv.data_ptr()
        s.grad_out = None if self.grad_out is None else self.grad_out.data_ptr()
        s.workspace, s.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel()
        s.word_frames = None if self.word_frames is None else self.word_frames.data_ptr()
        s.cls_w = None if self.cls_w is None else self.cls_w.data_ptr()
        s.cls_b = None if self.cls_b is None else self.cls_b.data_ptr()
        s.extra_terms = None if self.soma is None else self.soma.extra_terms.data_ptr()
        s.extra_grad = None if self.soma is None else self.soma.extra_grad.data_ptr()
        s.aux_log = None if self.aux_log is None else self.aux_log.data_ptr()
        s.bwd_fused_packed = None if self.bwd_fused_packed is None else self.bwd_fused_packed.data_ptr()
        s.post_t_packed = None if self.post_t_packed is None else self.post_t_packed.data_ptr()
        self._struct = s

    def struct_ref(self):
        return ctypes.byref(self._struct)

    def launches_per_step(self) -> int:
        """kernel launches of libpaule_b200.so per inner step of this planner (paule_plan_step_launches)"""
        with torch.cuda.device(self.device):
            return int(_lib.load().paule_plan_step_launches(self.struct_ref()))

    # ------------------------------------------------------------------------------------------
    def embed(self, mel_tm: torch.Tensor) -> torch.Tensor:
        """semvec [B,S] of a time-major mel [Tm,B,Cm] through the planner's own embedder kernels, in the planner's math
        (no grad; ragged batches read every word's own last frame)."""
        mel_tm = mel_tm.float().contiguous()
        if tuple(mel_tm.shape) != (self.Tm, self.B, self.Cm):
            raise ValueError(f"embed expects a time-major mel {(self.Tm, self.B, self.Cm)}, got {tuple(mel_tm.shape)}")
        sv = torch.empty((self.B, self.S), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            ops.plan_embed(mel_tm, sv, self.workspace, self._key)
        return sv

    def reset(self, initial_cp: torch.Tensor, target_mel: torch.Tensor, target_semvec: Optional[torch.Tensor] = None,
              past_cp: Optional[torch.Tensor] = None, lengths=None) -> None:
        """Re-arm the planner for another job of the SAME shape and options: new cps / targets, zeroed Adam state and loss
        log.  Every buffer keeps its address, so the captured CUDA graph, the workspace and the packed weights are reused."""
        B, T, C = initial_cp.shape
        if (B, T, C) != (self.B, self.T, self.C) or tuple(target_mel.shape) != (self.B, self.Tm, self.Cm):
            raise ValueError("reset() needs the shapes the planner was built for")
        if (past_cp is None) != (self.past_cp is None) or (lengths is not None and self.word_frames is None) or \
                (lengths is None and self.word_frames is not None):
            raise ValueError("reset() cannot add or remove past_cp / ragged lengths")
        with torch.cuda.device(self.device):
            self.cp.copy_(ops.transpose_btc(initial_cp.to(self.device).float().contiguous()))
            self.target_mel.copy_(ops.transpose_btc(target_mel.to(self.device).float().contiguous()))
            if past_cp is not None:
                pc = past_cp.to(self.device).float().contiguous()
                if pc.dim() == 2:
                    pc = pc.unsqueeze(0).expand(B, -1, -1).contiguous()
                pc = ops.transpose_btc(pc)
                if pc.shape != self.past_cp.shape:
                    raise ValueError("reset() needs a past_cp of the same length")
                self.past_cp.copy_(pc)
            if lengths is not None:
                ln = [int(v) for v in lengths]
                if len(ln) != B or min(ln) < 13 or max(ln) > T:
                    raise ValueError("every word needs 13 <= lengths[b] <= initial_cp.shape[1] cp frames")
                self.lengths = ln
                self.word_frames.copy_(torch.tensor(ln, dtype=torch.int32))
            self.adam_m.zero_(); self.adam_v.zero_(); self.step_count.zero_(); self.loss_log.zero_()
            if self.aux_log is not None:
                self.aux_log.zero_()
            self.steps_done = 0
            if target_semvec is None:
                self.target_sv.copy_(self.embed(self.target_mel))
            else:
                self.target_sv.copy_(target_semvec.to(self.device).float().reshape(self.B, self.S))

    def forward(self):
        """no_grad predictions for the current cps -> (pred_mel [B,Tm,Cm], pred_semvec [B,S])."""
        with torch.cuda.device(self.device):
            ops.plan_forward(self.cp, self.pred_mel, self.pred_sv, self.workspace, self._key)
            return ops.transpose_btc(self.pred_mel), self.pred_sv.clone()

    def _one_step(self) -> None:
        if self.soma is not None:   # tube terms + their d/d(cp) for the CURRENT cps, consumed by the fused step below
            self.soma.run(self.cp, self.target_mel, self.target_sv)
        ops.plan_step(self.cp, self.adam_m, self.adam_v, self.step_count, self.loss_log, self.pred_mel, self.pred_sv,
                      self.workspace, self._key)

    def step(self, n: int = 1) -> None:
        """n inner steps (optimizer.zero_grad() ... clamp, paule/paule.py:911-1211), no host sync."""
        if n <= 0:
            return
        if self.steps_done + n > self.max_log_steps:
            raise ValueError(f"loss log holds {self.max_log_steps} steps; construct with a larger max_log_steps")
        with torch.cuda.device(self.device):
            self._step(n)

    def _step(self, n: int) -> None:
        if self._use_graph and self._graph is None and not (self.soma is not None and self.soma.stochastic):
            # warm up on a side stream, then capture one step; Adam's step counter and the log slot live on the
            # device, so a replay is a full, correct step.
            state = [self.cp, self.adam_m, self.adam_v, self.step_count, self.loss_log]
            if self.aux_log is not None:
                state.append(self.aux_log)
            snap = [t.clone() for t in state]
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._one_step()
            torch.cuda.current_stream(self.device).wait_stream(s)
            for dst, src in zip(state, snap):
                dst.copy_(src)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._one_step()
            for dst, src in zip(state, snap):
                dst.copy_(src)
            self._graph = g
        for _ in range(n):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._one_step()
        self.steps_done += n

    # ------------------------------------------------------------------------------------------
    def check(self) -> None:
        """Raise if a persistent kernel's watchdog fired (an inter-CTA wait exceeded 4 s: results are invalid).  One 4-byte
        device->host read; called whenever results leave the planner, never inside the inner loop."""
        if self._status_off is not None:
            code, clamped = (int(v) for v in self.workspace[self._status_off:self._status_off + 8].view(torch.int32).tolist())
            if clamped and not self._warned_clamp:   # informational: a recurrent gradient hit the exchange bound of the bf16 BPTT
                warnings.warn(ops.CLAMPED_MSG, RuntimeWarning, stacklevel=2)
                self._warned_clamp = True
            if code == 0 and self.soma is not None and self.soma.status_words() is not None:
                code, soma_clamped = self.soma.status_words()
                if soma_clamped and not self._warned_clamp:
                    warnings.warn(ops.CLAMPED_MSG, RuntimeWarning, stacklevel=2)
                    self._warned_clamp = True
            if code != 0:
                raise _lib.PauleB200Error(f"persistent recurrent kernel watchdog fired (status {code}): results are invalid")

    def planned_cp(self) -> torch.Tensor:
        """current cps, batch-first [B,T,C]."""
        self.check()
        return ops.transpose_btc(self.cp)

    def set_cp(self, cp_bf: torch.Tensor) -> None:
        self.cp.copy_(ops.transpose_btc(cp_bf.float().contiguous()))

    def losses(self) -> Dict[str, torch.Tensor]:
        """per-step, per-word loss terms logged BEFORE each update (paule/paule.py:988): tensors [steps,B]."""
        self.check()
        log = self.loss_log[: self.steps_done]
        out = {"total": log[..., 0], "mel": log[..., 1], "semvec": log[..., 2], "velocity": log[..., 3],
               "jerk": log[..., 4], "local_linear": log[..., 5]}
        if self.aux_log is not None:
            aux = self.aux_log[: self.steps_done]
            out.update({"speech_classifier": aux[..., 0], "tube_mel": aux[..., 1], "tube_semvec": aux[..., 2]})
        return out

    def last_grad(self) -> Optional[torch.Tensor]:
        return None if self.grad_out is None else ops.transpose_btc(self.grad_out)

    def last_grad_lstm(self) -> torch.Tensor:
        """d(mel + semvec terms)/d(cp) of the last step, batch-first [B,T,C]: the BPTT result alone, i.e. ``xx_new.grad``
        (paule/paule.py:1052) minus the gradient of the velocity / jerk / local-linear terms (which the Adam kernel adds).
        This is what pins the model path: on iid cps the smoothness gradient is 10^5 times larger and hides it."""
        n = self.T * self.B * self.C
        view = self.workspace[self._grad_lstm_off:self._grad_lstm_off + 4 * n].view(torch.float32).view(self.T, self.B, self.C)
        return ops.transpose_btc(view.contiguous())

    def close(self) -> None:
        """Release the device buffers (workspace, Adam state, loss log, CUDA graph); the planner is unusable afterwards."""
        ops.unregister_plan(getattr(self, "_key", 0))
        self._graph = None
        self.workspace = None
        self.soma = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
