"""B200-native counterparts of the reference's hot-path models (same names, constructor arguments,
``forward`` signatures and ``state_dict`` keys as /root/reference/paule/models.py), running on the
hand-written CUDA kernels of libpaule_b200.so.

* ``ForwardModel``                         models.py:326-356   cp [B,T,30] -> log-mel [B,T//2,60]
* ``EmbeddingModel``                       models.py:413-448   log-mel [B,Tm,60], lens -> semvec [B,300]
* ``InverseModelMelTimeSmoothResidual``    models.py:177-247   log-mel [B,Tm,60] -> cp [B,2Tm,30] (no grad)
* aliases named by BASELINE.json: ``InverseModel``, ``MelEmbeddingModel``

The modules hold their parameters in ``torch.nn.LSTM`` / ``Linear`` / ``Conv1d`` containers purely so
that initialisation and ``state_dict`` naming are identical to the reference (reference checkpoints
load unchanged); the containers' own ``forward`` is never called.  Arithmetic is fp32 (BASELINE.json
configs[0]); float64 parameters (the reference's shipped dtype) are converted when packed.
``forward`` is differentiable with respect to its *input* (what planning needs, paule/paule.py:1052);
parameters receive no gradient.  CPU tensors raise: there is no fallback.
"""
from __future__ import annotations

import contextlib
import threading
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import _lib, ops


def _versions(params: Sequence[torch.Tensor]) -> Tuple:
    return tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in params)


class _PackedLSTM:
    """Caches the device operand pack of an nn.LSTM container; repacks when a parameter changes."""

    def __init__(self, lstm: nn.LSTM):
        self.lstm = lstm
        self._key = None
        self.layers: List[ops.LstmWeights] = []

    def get(self, tc: bool = False) -> List[ops.LstmWeights]:
        """tc: also hold the tcgen05 weight images of the persistent-RNN kernels.  Those are built for 720 hidden units;
        narrower layers are packed ZERO-PADDED to 720 units (exact, ops.pad_lstm_params) and ``lstm_stack`` drops the pad
        columns of the output again."""
        params = list(self.lstm.parameters())
        key = (_versions(params), bool(tc))
        if key != self._key:
            self.layers = []
            pad = tc and self.lstm.hidden_size < ops.TC_HIDDEN
            for k in range(self.lstm.num_layers):
                p = [getattr(self.lstm, f"{n}_l{k}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
                if pad:
                    p = ops.pad_lstm_params(*p, input_padded=k > 0)
                self.layers.append(ops.LstmWeights(*p, tc=tc))
            self.padded_from = self.lstm.hidden_size if pad else None
            self._key = key
        return self.layers


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


def _check_input(x: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise _lib.PauleB200Error(f"{name}: expected a CUDA tensor; paule_b200 has no CPU fallback")
    _lib.require_device()
    return x.float().contiguous()


def _learning(params) -> bool:
    """True when a weight gradient is wanted: grad mode on and a parameter requires grad (continue-learning of the models,
    paule/paule.py:1372-1377).  Planning runs under requires_grad=False weights or no_grad and takes the cached operands."""
    return torch.is_grad_enabled() and any(p.requires_grad for p in params)


def _param(p: torch.Tensor) -> torch.Tensor:
    return p if (p.dtype == torch.float32 and p.is_contiguous()) else p.float().contiguous()


def lstm_stack(x: torch.Tensor, layers: List[ops.LstmWeights], lstm: nn.LSTM = None, hidden: int = None) -> torch.Tensor:
    """x [B,T,I] batch-first -> top layer's h, TIME-MAJOR [T,B,H].  With ``lstm`` given and a weight gradient wanted, the
    live parameters go into the ops (so autograd reaches them) instead of the cached, detached operand pack."""
    h = x
    learn = lstm is not None and _learning(lstm.parameters())
    true_h = lstm.hidden_size if lstm is not None else hidden      # the model's own width (zero-padded tensor-core layers)
    if learn and layers and layers[0].packed is not None and true_h is not None and layers[0].hidden != true_h:
        # weight gradients of a zero-padded layer would come back padded: continue-learning of narrow models stays on the fp32 ops
        layers = [ops.LstmWeights(*(getattr(lstm, f"{n}_l{k}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")))
                  for k in range(lstm.num_layers)]
    for k, L in enumerate(layers):
        if learn:
            w_ih, w_hh = _param(getattr(lstm, f"weight_ih_l{k}")), _param(getattr(lstm, f"weight_hh_l{k}"))
            bias = _param(getattr(lstm, f"bias_ih_l{k}")) + _param(getattr(lstm, f"bias_hh_l{k}"))
        else:
            w_ih, w_hh, bias = L.w_ih, L.w_hh, L.bias
        if L.packed is not None:      # module.math = MATH_BF16: the persistent tcgen05 kernels (bf16 operands, fp32 state)
            h, _, _ = ops.lstm_layer_fwd_tc(h, k == 0, w_ih, w_hh, bias, L.packed)
        else:
            h, _, _ = ops.lstm_layer_fwd(h, k == 0, w_ih, w_hh, bias)
    if true_h is not None and h.shape[-1] != true_h:      # zero-padded tensor-core layers: the model's own units only
        h = h[..., :true_h].contiguous()
    return h


_SCOPE = threading.local()


@contextlib.contextmanager
def math_scope(math: int):
    """Arithmetic of module forwards inside the ``with`` block for modules WITHOUT their own ``math`` attribute (used by
    ``Paule.plan_resynth`` to run the inverse model's recurrence in the planner's arithmetic without touching the module)."""
    old = getattr(_SCOPE, "math", None)
    _SCOPE.math = math
    try:
        yield
    finally:
        _SCOPE.math = old


def _use_tc(module) -> bool:
    """``module.math = ops.MATH_BF16`` selects the tensor-core recurrences for the module's own forward / backward (the fused
    planner has its own ``math`` argument); without the attribute an enclosing ``math_scope`` decides, else fp32.  Only the
    hidden size the kernels are built for; fp32 otherwise."""
    math = getattr(module, "math", None)
    if math is None:
        math = getattr(_SCOPE, "math", None)
    if math is None:
        math = ops.MATH_FP32
    return math != ops.MATH_FP32 and module.lstm.hidden_size <= ops.TC_HIDDEN and ops.tc_available()


class ForwardModel(nn.Module):
    """Predictive forward model, cp -> log-mel (reference: paule/models.py:326-356)."""

    def __init__(self, input_size=30, output_size=60, hidden_size=180, num_lstm_layers=4, apply_half_sequence=True):
        super().__init__()
        self.apply_half_sequence = apply_half_sequence
        if self.apply_half_sequence:
            self.half_sequence = nn.AvgPool1d(2, stride=2)   # parameter-free; kept for attribute parity
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers=num_lstm_layers, batch_first=True)
        self.post_linear = nn.Linear(hidden_size, output_size)
        self._pack = _PackedLSTM(self.lstm)

    def forward(self, x, *args):
        x = _check_input(x, "ForwardModel.forward(x)")
        h = lstm_stack(x, self._pack.get(_use_tc(self)), self.lstm)
        if _learning(self.post_linear.parameters()):
            return ops.linear_tm(h, _param(self.post_linear.weight), _param(self.post_linear.bias),
                                 bool(self.apply_half_sequence), True)
        return ops.linear_tm(h, _f32c(self.post_linear.weight), _f32c(self.post_linear.bias),
                             bool(self.apply_half_sequence), True)


class EmbeddingModel(nn.Module):
    """Mel -> semantic-vector embedder (reference: paule/models.py:413-448)."""

    def __init__(self, input_size=60, output_size=300, hidden_size=720, num_lstm_layers=1,
                 post_activation=torch.nn.LeakyReLU(), post_upsampling_size=0, dropout=0):
        super().__init__()
        self.post_upsampling_size = post_upsampling_size
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers=num_lstm_layers, batch_first=True, dropout=dropout)
        if post_upsampling_size > 0:
            self.post_linear = nn.Linear(hidden_size, post_upsampling_size)
            self.linear_mapping = nn.Linear(post_upsampling_size, output_size)
            self.post_activation = post_activation
        else:
            self.linear_mapping = nn.Linear(hidden_size, output_size)
        self._pack = _PackedLSTM(self.lstm)

    def forward(self, x, lens, *args):
        x = _check_input(x, "EmbeddingModel.forward(x)")
        if self.training and self.lstm.dropout > 0 and self.lstm.num_layers > 1:
            raise _lib.PauleB200Error("inter-layer LSTM dropout in training mode is not implemented on the CUDA path "
                                      "(Paule's embedder uses dropout=0: paule/paule.py:167)")
        h = lstm_stack(x, self._pack.get(_use_tc(self)), hidden=self.lstm.hidden_size)   # [T,B,H]
        B = x.shape[0]
        idx = torch.as_tensor([int(l) - 1 for l in lens], device=x.device, dtype=torch.long)
        if idx.numel() != B:
            raise ValueError(f"lens has {idx.numel()} entries for a batch of {B}")
        last = h[idx, torch.arange(B, device=x.device)].contiguous()          # models.py:442 (device gather)
        last = last.unsqueeze(0)                                              # [1,B,H] time-major
        if self.post_upsampling_size > 0:
            z = ops.linear_tm(last, _f32c(self.post_linear.weight), _f32c(self.post_linear.bias), False, False)
            z = self.post_activation(z)
            out = ops.linear_tm(z.contiguous(), _f32c(self.linear_mapping.weight), _f32c(self.linear_mapping.bias),
                                False, False)
        else:
            out = ops.linear_tm(last, _f32c(self.linear_mapping.weight), _f32c(self.linear_mapping.bias), False, False)
        return out[0]


class _MelChannelConv1D(nn.Module):
    """Parameter container with the reference's naming (models.py:142-151)."""

    def __init__(self, input_units, filter_size_channel):
        super().__init__()
        self.filter_size_channel = filter_size_channel
        assert input_units % filter_size_channel == 0, 'output_size has to devisible by %d' % filter_size_channel
        output_units = int(input_units // filter_size_channel)
        self.ConvLayers = nn.ModuleList(
            [nn.Conv1d(input_units, output_units, 5, padding=2, groups=output_units) for _ in range(filter_size_channel)])

    def packed(self) -> Tuple[torch.Tensor, torch.Tensor]:
        # channel c = fs*g + i uses ConvLayers[i].weight[g] on the mel rows c-1, c, c+1 (fs = 3)
        w = torch.stack([_f32c(l.weight) for l in self.ConvLayers], dim=1)     # [G, fs, 3, 5]
        b = torch.stack([_f32c(l.bias) for l in self.ConvLayers], dim=1)       # [G, fs]
        return w.reshape(-1, w.shape[2], 5).contiguous(), b.reshape(-1).contiguous()


class _TimeConvResBlock(nn.Module):
    """Parameter container (models.py:114-129), filter size 5, channelwise."""

    def __init__(self, units):
        super().__init__()
        self.band_conv1d_1 = nn.Conv1d(units, units, 5, padding=2, groups=units)
        self.band_conv1d_2 = nn.Conv1d(units, units, 5, padding=2, groups=units)


class InverseModelMelTimeSmoothResidual(nn.Module):
    """Inverse model, log-mel -> initial cp (reference: paule/models.py:177-247).  Forward only.

    Supported configuration = what ``Paule`` instantiates (paule/paule.py:146): identity activations,
    ``mel_smooth_filter_size=3``, ``time_filter_size=5``, ``lstm_resid=True``."""

    def __init__(self, input_size=60, output_size=30, hidden_size=180, num_lstm_layers=4, mel_smooth_layers=3,
                 mel_smooth_filter_size=3, mel_resid_activation=torch.nn.Identity(), resid_blocks=5,
                 time_filter_size=5, pre_resid_activation=torch.nn.Identity(),
                 post_resid_activation=torch.nn.Identity(), output_activation=torch.nn.Identity(), lstm_resid=True):
        super().__init__()
        for act in (mel_resid_activation, pre_resid_activation, post_resid_activation, output_activation):
            if not isinstance(act, torch.nn.Identity):
                raise NotImplementedError("only the Identity activations Paule uses are implemented on the CUDA path")
        if mel_smooth_filter_size != 3 or time_filter_size != 5 or not lstm_resid:
            raise NotImplementedError("CUDA path supports mel_smooth_filter_size=3, time_filter_size=5, lstm_resid=True")
        self.lstm_resid = lstm_resid
        self.MelBlocks = nn.ModuleList([_MelChannelConv1D(input_size, mel_smooth_filter_size)
                                        for _ in range(mel_smooth_layers)])
        self.lstm = nn.LSTM(3 * input_size, hidden_size, num_layers=num_lstm_layers, batch_first=True)
        self.post_linear = nn.Linear(hidden_size, output_size)
        self.ResidualConvBlocks = nn.ModuleList([_TimeConvResBlock(output_size) for _ in range(resid_blocks)])
        if self.lstm_resid and len(self.ResidualConvBlocks) > 0:
            self.resid_weighting = nn.Conv1d(2 * output_size, output_size, time_filter_size, padding=2,
                                             groups=output_size)
        self._pack = _PackedLSTM(self.lstm)

    def forward(self, x, *args):
        """Planning: the hand-written stencil kernels + LSTM ops, no autograd graph (the reference only ever calls the inverse
        model under no_grad while planning, paule/paule.py:551-557).  Continue-learning of the inverse model
        (``self.learnable = True``, grad mode on, parameters require grad; paule/paule.py:1413-1436, set by
        ``Paule.continue_learning_inv``): the differentiable restatement ``_forward_trainable`` -- outer-loop work, a few small
        batches between planning rounds."""
        if getattr(self, "learnable", False) and _learning(self.parameters()):
            return self._forward_trainable(_check_input(x, "InverseModel.forward(x)"))
        with torch.no_grad():
            return self._forward_kernels(x)

    def _forward_trainable(self, x):
        """models.py:221-247 with autograd: the recurrence and post_linear on the library's LSTM / Linear ops (own forward, BPTT
        and weight-gradient kernels), the per-channel stencils -- 1 % of the FLOPs -- as grouped ``conv1d`` library calls so
        that autograd reaches their weights."""
        F = nn.functional
        B, Tm, Cm = x.shape
        z = x.transpose(1, 2)                                                   # [B, mel, seq]
        for blk in self.MelBlocks:                                              # models.py:152-169, :224-228
            fs = blk.filter_size_channel
            outs = []
            for i, conv in enumerate(blk.ConvLayers):                           # layer i sees the mel axis shifted by (fs - 2) - i rows
                shift = (fs - 2) - i
                if shift > 0:
                    zi = F.pad(z, (0, 0, shift, 0))[:, :Cm, :]
                elif shift < 0:
                    zi = F.pad(z, (0, 0, 0, -shift))[:, -Cm:, :]
                else:
                    zi = z
                outs.append(F.conv1d(zi, conv.weight.float(), conv.bias.float(), padding=2, groups=conv.groups))
            z = torch.stack(outs, dim=2).reshape(B, Cm, Tm) + z
        z = z.transpose(1, 2)
        zero = z.new_zeros(B, 1, Cm)                                            # add_vel_and_acc_info, models.py:47-61
        vel = z[:, 1:] - z[:, :-1]
        acc = vel[:, 1:] - vel[:, :-1]
        feat = torch.cat((z, torch.cat((vel, zero), 1), torch.cat((zero, acc, zero), 1)), dim=2).contiguous()
        h = lstm_stack(feat, self._pack.get(_use_tc(self)), self.lstm)          # [Tm,B,H]
        y = ops.linear_tm(h, _param(self.post_linear.weight), _param(self.post_linear.bias), False, True)   # [B,Tm,30]
        mid = torch.cat(((y[:, :-1] + y[:, 1:]) / 2.0, y[:, -1:]), dim=1)       # double_sequence, models.py:63-81
        out = torch.stack((y, mid), dim=2).reshape(B, 2 * Tm, y.shape[2]).transpose(1, 2)   # [B,30,2Tm]
        raw = out
        for r in self.ResidualConvBlocks:                                       # models.py:131-139
            c1, c2 = r.band_conv1d_1, r.band_conv1d_2
            out = F.conv1d(F.conv1d(out, c1.weight.float(), c1.bias.float(), padding=2, groups=c1.groups),
                           c2.weight.float(), c2.bias.float(), padding=2, groups=c2.groups) + out
        if len(self.ResidualConvBlocks) > 0:
            Cc, S = out.shape[1], out.shape[2]
            mixed = torch.stack((out, raw), dim=2).reshape(B, 2 * Cc, S)        # models.py:241-243
            rw = self.resid_weighting
            out = F.conv1d(mixed, rw.weight.float(), rw.bias.float(), padding=2, groups=rw.groups)
        return out.transpose(1, 2).contiguous()

    def _forward_kernels(self, x):
        x = _check_input(x, "InverseModel.forward(x)")
        lib = _lib.load()
        st = ops._stream()
        B, Tm, Cm = x.shape
        cur = x
        for blk in self.MelBlocks:
            w, b = blk.packed()
            nxt = torch.empty_like(cur)
            _lib.check(lib.paule_melconv_res_f32(cur.data_ptr(), w.data_ptr(), b.data_ptr(), nxt.data_ptr(), B, Tm, Cm,
                                                 st), "paule_melconv_res_f32")
            cur = nxt
        feat = torch.empty((B, Tm, 3 * Cm), device=x.device, dtype=torch.float32)
        _lib.check(lib.paule_vel_acc_f32(cur.data_ptr(), feat.data_ptr(), B, Tm, Cm, st), "paule_vel_acc_f32")
        h = lstm_stack(feat, self._pack.get(_use_tc(self)), hidden=self.lstm.hidden_size)           # [Tm,B,H]
        y = ops.linear_tm(h, _f32c(self.post_linear.weight), _f32c(self.post_linear.bias), False, True)  # [B,Tm,30]
        Cc = y.shape[2]
        nb = len(self.ResidualConvBlocks)
        out = torch.empty((B, 2 * Tm, Cc), device=x.device, dtype=torch.float32)
        scratch = torch.empty((3, B, 2 * Tm, Cc), device=x.device, dtype=torch.float32)
        if nb > 0:
            res_w = torch.stack([torch.stack((_f32c(r.band_conv1d_1.weight).reshape(Cc, 5),
                                              _f32c(r.band_conv1d_2.weight).reshape(Cc, 5)))
                                 for r in self.ResidualConvBlocks]).contiguous()       # [nb,2,C,5]
            res_b = torch.stack([torch.stack((_f32c(r.band_conv1d_1.bias), _f32c(r.band_conv1d_2.bias)))
                                 for r in self.ResidualConvBlocks]).contiguous()       # [nb,2,C]
            mix_w = _f32c(self.resid_weighting.weight)                                 # [C,2,5]: [c][0]=smoothed, [c][1]=raw
            mix_b = _f32c(self.resid_weighting.bias)
            _lib.check(lib.paule_upsample_smooth_f32(y.data_ptr(), res_w.data_ptr(), res_b.data_ptr(), nb,
                                                     mix_w.data_ptr(), mix_b.data_ptr(), out.data_ptr(),
                                                     scratch.data_ptr(), B, Tm, Cc, st), "paule_upsample_smooth_f32")
        else:
            _lib.check(lib.paule_upsample_smooth_f32(y.data_ptr(), None, None, 0, None, None, out.data_ptr(),
                                                     scratch.data_ptr(), B, Tm, Cc, st), "paule_upsample_smooth_f32")
        return out


# names used by BASELINE.json's north_star
class MelEmbeddingModelMelSmoothResidualUpsampling(nn.Module):
    """Mel -> semantic-vector embedder with mel-channel smoothing and a post-upsampling layer (reference:
    paule/models.py:362-409; defined there but not instantiated by ``Paule``, which uses ``EmbeddingModel``).  Forward only;
    same ``state_dict`` keys as the reference.  Supported: ``mel_smooth_filter_size=3``, Identity ``mel_resid_activation``."""

    def __init__(self, input_size=60, output_size=300, hidden_size=180, num_lstm_layers=4, mel_smooth_layers=3,
                 mel_smooth_filter_size=3, mel_resid_activation=torch.nn.Identity(), post_activation=torch.nn.LeakyReLU(),
                 post_upsampling_size=8192):
        super().__init__()
        if not isinstance(mel_resid_activation, torch.nn.Identity) or mel_smooth_filter_size != 3:
            raise NotImplementedError("CUDA path supports mel_smooth_filter_size=3 and the Identity mel_resid_activation")
        self.mel_resid_activation = mel_resid_activation
        self.MelBlocks = nn.ModuleList([_MelChannelConv1D(input_size, mel_smooth_filter_size)
                                        for _ in range(mel_smooth_layers)])
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers=num_lstm_layers, batch_first=True)
        self.post_linear = nn.Linear(hidden_size, post_upsampling_size)
        self.upsampling = nn.Linear(post_upsampling_size, output_size)
        self.post_activation = post_activation
        self._pack = _PackedLSTM(self.lstm)

    @torch.no_grad()
    def forward(self, x, lens, *args):
        x = _check_input(x, "MelEmbeddingModelMelSmoothResidualUpsampling.forward(x)")
        lib = _lib.load()
        st = ops._stream()
        B, Tm, Cm = x.shape
        cur = x
        for blk in self.MelBlocks:                                            # x = x + conv(x), models.py:391-397
            w, b = blk.packed()
            nxt = torch.empty_like(cur)
            _lib.check(lib.paule_melconv_res_f32(cur.data_ptr(), w.data_ptr(), b.data_ptr(), nxt.data_ptr(), B, Tm, Cm,
                                                 st), "paule_melconv_res_f32")
            cur = nxt
        h = lstm_stack(cur, self._pack.get(_use_tc(self)), hidden=self.lstm.hidden_size)  # [T,B,H]
        idx = torch.as_tensor([int(l) - 1 for l in lens], device=x.device, dtype=torch.long)
        if idx.numel() != B:
            raise ValueError(f"lens has {idx.numel()} entries for a batch of {B}")
        last = h[idx, torch.arange(B, device=x.device)].contiguous().unsqueeze(0)      # models.py:403
        z = ops.linear_tm(last, _f32c(self.post_linear.weight), _f32c(self.post_linear.bias), False, False)
        z = self.post_activation(z)
        return ops.linear_tm(z.contiguous(), _f32c(self.upsampling.weight), _f32c(self.upsampling.bias), False, False)[0]


class LinearClassifier(nn.Module):
    """Speech / non-speech classifier on log-mel frames (reference: paule/models.py:887-911): Linear(input_dim -> 1) per frame,
    averaged over the word's frames -> one logit per word.  In the planner its loss term is fused into the criterion kernel
    (``paule_plan.cls_w``); this ``forward`` serves the produced side and stand-alone use."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.linear = nn.Linear(input_dim, output_dim)

    def forward(self, input_, *, src_lens=None):
        x = _check_input(input_, "LinearClassifier.forward(input_)")                 # [B,Tm,input_dim]
        y = ops.linear_tm(ops.transpose_btc(x), _f32c(self.linear.weight), _f32c(self.linear.bias), False, True)
        y = torch.squeeze(y, 2)                                                      # [B,Tm]
        if src_lens is None:
            return y.mean(dim=1)
        lens = torch.as_tensor([int(l) for l in src_lens], device=y.device)
        keep = torch.arange(y.shape[1], device=y.device).unsqueeze(0) < lens.unsqueeze(1)   # padded frames count as 0
        return (y * keep).sum(dim=1) / lens


class Generator(nn.Module):
    """Conditional GAN generator semvec (+ noise) -> cp or mel trajectory (reference: paule/models.py:594-652), used ONLY in the
    prologue of ``plan_resynth`` (``initialize_from='semvec'`` :558-565, missing acoustic target :515-522) -- once per call,
    outside the planning loop.  Same parameter names / ``state_dict`` as the reference; the forward is a short chain of
    library ops on the device (conv / batch-norm / linear upsampling), not a hand-written kernel."""

    def __init__(self, channel_noise=100, embed_size=300, fc_size=1024, inital_seq_length=4, hidden_size=256,
                 num_res_blocks=5, output_size=30):
        super().__init__()
        self.fc_size, self.hidden_size = fc_size, hidden_size
        self.fc_reshaped_size = int(fc_size / inital_seq_length)
        self.fully_connected = nn.Linear(channel_noise + embed_size, fc_size)
        chans = [self.fc_reshaped_size] + [hidden_size] * num_res_blocks
        self.res_blocks = nn.ModuleList(
            nn.Sequential(nn.Conv1d(c_in, c_out, kernel_size=5, stride=1, padding=2), nn.BatchNorm1d(c_out), nn.LeakyReLU(0.2))
            for c_in, c_out in zip(chans[:-1], chans[1:]))
        self.post_linear = nn.Linear(hidden_size, output_size)
        self.final_smoothing = nn.Conv1d(output_size, output_size, kernel_size=5, padding=2, groups=output_size)
        self.output_activation = nn.Tanh()

    def forward(self, x, length, vector):
        z = self.fully_connected(torch.cat([x, vector.unsqueeze(1)], dim=2))         # [B,1,fc]
        z = z.view(len(x), self.fc_reshaped_size, -1)                                # [B,fc/4,4]
        n = len(self.res_blocks)
        for i, block in enumerate(self.res_blocks):                                  # grow the sequence to `length` in n stages
            z = nn.functional.interpolate(z, size=int(length / (n - i)), mode='linear', align_corners=False)
            y = block(z)
            z = y + z if (i > 0 or self.fc_reshaped_size == self.hidden_size) else y
        z = self.post_linear(z.permute(0, 2, 1)).permute(0, 2, 1)                    # [B,out,length]
        z = self.final_smoothing(z) + z
        return self.output_activation(z.permute(0, 2, 1))


InverseModel = InverseModelMelTimeSmoothResidual
# BASELINE.json's north_star says "MelEmbeddingModel" for the mel-to-embedding model of the planning loop, which in the reference
# is EmbeddingModel (paule/paule.py:167); the class the reference calls MelEmbeddingModel... is available under its own name.
MelEmbeddingModel = EmbeddingModel
