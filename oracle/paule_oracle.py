"""CPU oracle for the PAULE gradient-planning inner loop.  TEST INFRASTRUCTURE ONLY.

Nothing in ``paule_b200/`` may import this file.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, as the checker or as
the timed CPU baseline -- never as the product path.

What it restates (all paths relative to /root/reference):

* ``ForwardModel``      paule/models.py:326-356   LSTM(30->H) -> Linear(H->60) -> AvgPool1d(2,2) over time
* ``EmbeddingModel``    paule/models.py:413-448   LSTM(60->H, L layers) -> h at lens-1 -> Linear(H->300)
* ``InverseModelMelTimeSmoothResidual``  paule/models.py:177-247 (+ helpers :47-81, :114-169)
* loss helpers          paule/util.py:564-574 (RMSE, eps=0), :600 (5-point stencil), :608-614 (local_linear),
                        :617-637 (vel/acc/jerk); weights paule/paule.py:592-597; criterion :647-662 (+ :705-717, :760-773)
* the inner loop        paule/paule.py:910-913, :921-924, :986-997, :1052, :1199-1211
* the optimiser         torch.optim.Adam([cp], lr)  paule/paule.py:797  (third-party arithmetic: torch)

Third-party arithmetic: the numbers of the reference are produced by PyTorch (``torch.nn.LSTM``,
autograd, ``torch.optim.Adam``; pin in the reference is ``torch >= 1.13.1``, pyproject.toml:33; this
image has torch 2.11.0).  The oracle therefore calls the *same* torch CPU operators through its own
module definitions, so that it is the reference's arithmetic, restated batched with per-word losses
(the reference is strictly batch-1: paule/paule.py:539,826).  A second, independent restatement
(``manual_*`` below: explicit LSTM cell loop, hand-derived BPTT, hand-written Adam) exists to check
the analytic gradients the CUDA kernels implement.

Pinning: the reference ships no golden vectors for this path (tests/test_paule.py:65-70 asserts
nothing), so the oracle is pinned against outputs of the reference itself: ``tests/golden/make_golden.py``
runs the real ``paule.paule.Paule.plan_resynth`` and the real ``paule.models`` here (where /root/reference
exists) and commits the vectors; ``tests/test_oracle.py`` checks this file against them.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Optional, Sequence

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

# paule/paule.py:592-597
MEL_WEIGHT = 5.0
VELOCITY_WEIGHT = 80.0
JERK_WEIGHT = 400.0
SEMANTIC_WEIGHT = 10.0
LOCAL_LINEAR_WEIGHT = 100_000.0
SPEECH_CLASSIFIER_WEIGHT = 0.1
TUBE_MEL_WEIGHT = MEL_WEIGHT
TUBE_SEMANTIC_WEIGHT = SEMANTIC_WEIGHT
CLAMP = 1.05  # paule/paule.py:1202


# ----------------------------------------------------------------------------------------------
# models (same parameter creation order as the reference constructors, so that torch.manual_seed(s)
# followed by construction yields bit-identical random-init weights; same state_dict keys)
# ----------------------------------------------------------------------------------------------
class OracleForwardModel(nn.Module):
    """paule/models.py:326-356."""

    def __init__(self, input_size=30, output_size=60, hidden_size=180, num_lstm_layers=4,
                 apply_half_sequence=True):
        super().__init__()
        self.apply_half_sequence = apply_half_sequence
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers=num_lstm_layers, batch_first=True)
        self.post_linear = nn.Linear(hidden_size, output_size)

    def forward(self, x, *args):
        out, _ = self.lstm(x)
        out = self.post_linear(out)
        if self.apply_half_sequence:
            # AvgPool1d(2, stride=2) over time, floor(T/2) frames (models.py:344,352-354)
            out = F.avg_pool1d(out.transpose(1, 2), 2, stride=2).transpose(1, 2)
        return out


class OracleEmbeddingModel(nn.Module):
    """paule/models.py:413-448."""

    def __init__(self, input_size=60, output_size=300, hidden_size=720, num_lstm_layers=1,
                 post_activation=None, post_upsampling_size=0, dropout=0):
        super().__init__()
        self.post_upsampling_size = post_upsampling_size
        self.lstm = nn.LSTM(input_size, hidden_size, num_layers=num_lstm_layers, batch_first=True,
                            dropout=dropout)
        if post_upsampling_size > 0:
            self.post_linear = nn.Linear(hidden_size, post_upsampling_size)
            self.linear_mapping = nn.Linear(post_upsampling_size, output_size)
            self.post_activation = post_activation if post_activation is not None else nn.LeakyReLU()
        else:
            self.linear_mapping = nn.Linear(hidden_size, output_size)

    def forward(self, x, lens, *args):
        out, _ = self.lstm(x)
        idx = torch.as_tensor([int(l) - 1 for l in lens], dtype=torch.long)
        out = out[torch.arange(out.shape[0]), idx, :]          # models.py:442
        if self.post_upsampling_size > 0:
            out = self.post_activation(self.post_linear(out))
        return self.linear_mapping(out)


class _OracleMelChannelConv(nn.Module):
    """paule/models.py:142-169 -- per-channel untied 3(mel) x 5(time) convolution."""

    def __init__(self, input_units, filter_size_channel):
        super().__init__()
        assert input_units % filter_size_channel == 0
        self.fs = filter_size_channel
        out_units = input_units // filter_size_channel
        self.ConvLayers = nn.ModuleList(
            [nn.Conv1d(input_units, out_units, 5, padding=2, groups=out_units)
             for _ in range(filter_size_channel)])

    def forward(self, x):                       # x [B, mel, seq]
        b, mel, seq = x.shape
        # neighbour stacks: layer i sees the mel axis shifted by (i - (fs-2)) rows, zero filled
        # (models.py:155-160): layers 0..fs-3 look "down" by fs-2-i rows, layer fs-2 is unshifted,
        # layer fs-1 looks one row "up".
        outs = []
        for i, conv in enumerate(self.ConvLayers):
            shift = (self.fs - 2) - i          # >0: rows move towards higher index
            if shift > 0:
                xi = F.pad(x, (0, 0, shift, 0))[:, :mel, :]
            elif shift < 0:
                xi = F.pad(x, (0, 0, 0, -shift))[:, -mel:, :]
            else:
                xi = x
            outs.append(conv(xi))
        # interleave: output channel fs*g + i = outs[i][:, g]   (models.py:166-167)
        return torch.stack(outs, dim=2).reshape(b, mel, seq)


class _OracleTimeConvResBlock(nn.Module):
    """paule/models.py:114-139 with filter_size 5, channelwise, identity activations."""

    def __init__(self, units):
        super().__init__()
        self.band_conv1d_1 = nn.Conv1d(units, units, 5, padding=2, groups=units)
        self.band_conv1d_2 = nn.Conv1d(units, units, 5, padding=2, groups=units)

    def forward(self, x):
        return self.band_conv1d_2(self.band_conv1d_1(x)) + x


class OracleInverseModel(nn.Module):
    """paule/models.py:177-247 (identity activations, lstm_resid=True).

    ``double_sequence`` follows the input dtype here; the reference hard-codes float64
    (models.py:76) and therefore only runs in fp64.
    """

    def __init__(self, input_size=60, output_size=30, hidden_size=180, num_lstm_layers=4,
                 mel_smooth_layers=3, mel_smooth_filter_size=3, resid_blocks=5, time_filter_size=5):
        super().__init__()
        assert time_filter_size == 5
        self.MelBlocks = nn.ModuleList(
            [_OracleMelChannelConv(input_size, mel_smooth_filter_size) for _ in range(mel_smooth_layers)])
        self.lstm = nn.LSTM(3 * input_size, hidden_size, num_layers=num_lstm_layers, batch_first=True)
        self.post_linear = nn.Linear(hidden_size, output_size)
        self.ResidualConvBlocks = nn.ModuleList(
            [_OracleTimeConvResBlock(output_size) for _ in range(resid_blocks)])
        self.resid_weighting = nn.Conv1d(2 * output_size, output_size, time_filter_size, padding=2,
                                         groups=output_size)

    @staticmethod
    def add_vel_and_acc_info(x):               # models.py:47-61
        z = x.new_zeros(x.shape[0], 1, x.shape[2])
        vel = x[:, 1:] - x[:, :-1]
        acc = vel[:, 1:] - vel[:, :-1]
        return torch.cat((x, torch.cat((vel, z), 1), torch.cat((z, acc, z), 1)), dim=2)

    @staticmethod
    def double_sequence(x):                    # models.py:63-81
        mid = torch.cat(((x[:, :-1] + x[:, 1:]) / 2.0, x[:, -1:]), dim=1)
        return torch.stack((x, mid), dim=2).reshape(x.shape[0], 2 * x.shape[1], x.shape[2])

    def forward(self, x, *args):
        x = x.transpose(1, 2)
        for blk in self.MelBlocks:
            x = blk(x) + x
        x = self.add_vel_and_acc_info(x.transpose(1, 2))
        out, _ = self.lstm(x)
        out = self.double_sequence(self.post_linear(out)).transpose(1, 2)     # [B, 30, 2Tm]
        raw = out
        for blk in self.ResidualConvBlocks:
            out = blk(out)
        b, c, s = out.shape
        mixed = torch.stack((out, raw), dim=2).reshape(b, 2 * c, s)            # models.py:241-243
        return self.resid_weighting(mixed).transpose(1, 2)


def build_reference_models(seed: int = 0, hidden_size: int = 720, dtype=torch.float32,
                           with_inverse: bool = True):
    """Random-init weights of SURVEY section 8(d): construct pred, embedder, inverse in this order
    under ``torch.manual_seed(seed)`` (paule/paule.py:124,146,167 hyper-parameters)."""
    torch.manual_seed(seed)
    pred = OracleForwardModel(num_lstm_layers=1, hidden_size=hidden_size)
    emb = OracleEmbeddingModel(num_lstm_layers=2, hidden_size=hidden_size)
    inv = OracleInverseModel(num_lstm_layers=1, hidden_size=hidden_size) if with_inverse else None
    mods = [m for m in (pred, emb, inv) if m is not None]
    for m in mods:
        m.to(dtype)
        for p in m.parameters():
            p.requires_grad_(False)
    return pred, emb, inv


def build_branch_models(dtype=torch.float32):
    """Seeded random-init models of the optional branches, as ``tests/golden/make_branches_golden.py::branch_models``
    builds them from the reference classes (paule/paule.py:227-273 hyper-parameters; tube embedder without dropout;
    classifier weight x30 so that its gradient matters): (cp_tube, tube_mel, tube_embedder, speech_classifier)."""
    torch.manual_seed(1)
    cp_tube = OracleForwardModel(num_lstm_layers=1, hidden_size=360, output_size=10, input_size=30,
                                 apply_half_sequence=False)
    tube_mel = OracleForwardModel(num_lstm_layers=1, hidden_size=360, output_size=60, input_size=10,
                                  apply_half_sequence=True)
    tube_emb = OracleEmbeddingModel(input_size=10, num_lstm_layers=2, hidden_size=720, dropout=0.0,
                                    post_upsampling_size=0)
    torch.manual_seed(2)
    cls = OracleLinearClassifier(input_dim=60, output_dim=1)
    with torch.no_grad():
        cls.linear.weight.mul_(30.0)
    mods = (cp_tube, tube_mel, tube_emb, cls)
    for m in mods:
        m.to(dtype)
        for p in m.parameters():
            p.requires_grad_(False)
    return mods


def state_dict_digest(module: nn.Module) -> str:
    h = hashlib.sha256()
    for k, v in module.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def synthetic_inputs(B: int, T: int, seed: int = 5, dtype=torch.float32, smooth: bool = False):
    """cp0 ~ U(-0.5,0.5) iid [B,T,30] (well-conditioned) or a sum of three slow
    sinusoids (|x|<=0.75, the chaotic regime); target_mel ~ U(0,1) [B,T//2,60]."""
    g = torch.Generator().manual_seed(seed)
    Tm = T // 2
    if smooth:
        t = torch.arange(T, dtype=torch.float64)[None, :, None]
        f = torch.rand(B, 1, 30, 3, generator=g, dtype=torch.float64) * 0.02 + 0.002
        ph = torch.rand(B, 1, 30, 3, generator=g, dtype=torch.float64) * 6.283185307179586
        cp0 = (0.25 * torch.sin(6.283185307179586 * f * t[..., None] + ph)).sum(-1)
    else:
        cp0 = torch.rand(B, T, 30, generator=g, dtype=torch.float64) - 0.5
    tmel = torch.rand(B, Tm, 60, generator=g, dtype=torch.float64)
    return cp0.to(dtype), tmel.to(dtype)


# ----------------------------------------------------------------------------------------------
# losses (per word: the reference is batch-1, so "mean over all elements" == per-word mean)
# ----------------------------------------------------------------------------------------------
def five_point(v):                              # util.py:600, delta_t = 1
    return (-v[:, 4:] + 8.0 * v[:, 3:-1] - 8.0 * v[:, 1:-3] + v[:, :-4]) / 12.0


def local_linear(v):                            # util.py:613-614
    return (2 * v[:, 1:-1] - v[:, :-2] - v[:, 2:]) / 2.0


def _wmean(t):
    return t.flatten(1).mean(1)


def per_word_losses(pred_mel, target_mel, pred_sv, target_sv, cp, objective="acoustic_semvec"):
    """criterion of paule/paule.py:647-662 (acoustic_semvec), :705-717 (acoustic), :760-773 (semvec).

    Returns (total[B], terms[B,5]) with terms = (mel, semvec, vel, jerk, local_linear), weighted.
    Terms not in the objective are still reported (as the reference logs the mel loss for the semvec
    objective, paule.py:1023) but excluded from ``total``."""
    vel = five_point(cp)                        # util.py:634
    jerk = five_point(five_point(vel))          # util.py:635-636
    ll = local_linear(cp)
    mel = MEL_WEIGHT * _wmean((pred_mel - target_mel) ** 2).sqrt()            # RMSE eps=0, paule.py:68
    sem = SEMANTIC_WEIGHT * _wmean((pred_sv - target_sv) ** 2).sqrt()
    v = VELOCITY_WEIGHT * _wmean(vel ** 2)                                    # MSE: paule.py:650
    j = JERK_WEIGHT * _wmean(jerk ** 2)
    l = LOCAL_LINEAR_WEIGHT * _wmean(ll ** 2)
    if objective == "acoustic_semvec":
        total = mel + v + j + sem + l           # summation order of paule.py:660
    elif objective == "acoustic":
        total = mel + v + j + l                 # paule.py:715
    elif objective == "semvec":
        total = v + j + sem + l                 # paule.py:771
    else:
        raise ValueError("objective has to be one of 'acoustic_semvec', 'acoustic' or 'semvec'")
    return total, torch.stack((mel, sem, v, j, l), dim=1)


# ----------------------------------------------------------------------------------------------
# the inner loop (bit-identical to plan_resynth for B=1)
# ----------------------------------------------------------------------------------------------
def plan_inner_loop(pred: nn.Module, emb: nn.Module, cp0: torch.Tensor, target_mel: torch.Tensor,
                    n_steps: int, *, target_semvec: Optional[torch.Tensor] = None, lr: float = 0.01,
                    objective: str = "acoustic_semvec", smiling: bool = False,
                    past_cp: Optional[torch.Tensor] = None, log_grads: bool = False,
                    log_cps: bool = False, teacher_cps: Optional[Sequence[torch.Tensor]] = None
                    ) -> Dict[str, object]:
    """Batched inner loop.  cp0 [B,T,30], target_mel [B,T//2,60].

    teacher_cps: if given, step k starts from teacher_cps[k] instead of the running cp (Adam state
    still carries over) -- used for teacher-forced parity, which is immune to chaotic drift."""
    B, T, _ = cp0.shape
    lens = tuple(torch.tensor(target_mel.shape[1]) for _ in range(B))
    with torch.no_grad():
        tsv = emb(target_mel, lens) if target_semvec is None else target_semvec      # paule.py:533-540
    x = cp0.clone().requires_grad_()                                                  # paule.py:585-590
    opt = torch.optim.Adam([x], lr=lr)                                                # paule.py:797
    out: Dict[str, object] = {"loss": [], "terms": [], "grads": [], "cps": [], "pred_mel": None}
    for k in range(n_steps):
        if teacher_cps is not None:
            with torch.no_grad():
                x.copy_(teacher_cps[k])
        opt.zero_grad()                                                               # :911
        mel = pred(x)                                                                 # :913
        sv = emb(mel, lens)                                                           # :924
        total, terms = per_word_losses(mel, target_mel, sv, tsv, x, objective)        # :986
        out["loss"].append(total.detach().clone())                                    # logged BEFORE the update (:988)
        out["terms"].append(terms.detach().clone())
        if log_cps:
            out["cps"].append(x.detach().clone())                                     # :1066
        total.sum().backward()                                                        # :1052 (words independent)
        if log_grads:
            out["grads"].append(x.grad.detach().clone())
        opt.step()                                                                    # :1199
        with torch.no_grad():
            x.data = x.data.clamp(-CLAMP, CLAMP)                                      # :1202
            if smiling:                                                               # :1203-1208
                x.data[:, :, 4] = -1.0
                x.data[:, :, 1] = 1.0
            if past_cp is not None:                                                   # :1210-1211
                x.data[:, 0:past_cp.shape[-2], :] = past_cp
    with torch.no_grad():
        out["pred_mel"] = pred(x)                                                     # :1460-1464
        out["pred_semvec"] = emb(out["pred_mel"], lens)
    out["planned_cp"] = x.detach().clone()
    out["target_semvec"] = tsv
    out["loss"] = torch.stack(out["loss"]) if out["loss"] else torch.zeros(0, B)
    out["terms"] = torch.stack(out["terms"]) if out["terms"] else torch.zeros(0, B, 5)
    return out


def model_path_grad(pred: nn.Module, emb: nn.Module, cp: torch.Tensor, target_mel: torch.Tensor,
                    target_sv: Optional[torch.Tensor] = None, objective: str = "acoustic_semvec",
                    lens: Optional[Sequence[int]] = None) -> torch.Tensor:
    """d(mel term + semvec term)/d(cp) [B,T,30] by autograd: ``xx_new.grad`` of paule/paule.py:1052 WITHOUT the velocity /
    jerk / local-linear terms, i.e. exactly what the LSTM BPTT kernels have to produce.  On the synthetic inputs this part is
    10^4..10^6 times smaller than the smoothness gradient, so a parity test on the total gradient cannot see it.

    ``lens`` (cp frames per word, ragged batches): word b is evaluated alone on its own ``lens[b]`` frames, as the reference
    does (one word per call); its gradient is zero on the padding."""
    B, T, _ = cp.shape
    g = torch.zeros_like(cp)
    words = range(B)
    for b in words:
        L = T if lens is None else int(lens[b])
        x = cp[b:b + 1, :L].clone().requires_grad_()
        tm = target_mel[b:b + 1, :L // 2]
        ln = (torch.tensor(L // 2),)
        with torch.no_grad():
            tsv = emb(tm, ln) if target_sv is None else target_sv[b:b + 1]
        mel = pred(x)
        loss = x.new_zeros(())
        if objective in ("acoustic_semvec", "acoustic"):
            loss = loss + MEL_WEIGHT * ((mel - tm) ** 2).mean().sqrt()
        if objective in ("acoustic_semvec", "semvec"):
            sv = emb(mel, ln)
            loss = loss + SEMANTIC_WEIGHT * ((sv - tsv) ** 2).mean().sqrt()
        loss.backward()
        g[b, :L] = x.grad
    return g


# ----------------------------------------------------------------------------------------------
# optional loss branches: speech classifier and somatosensory feedback
# ----------------------------------------------------------------------------------------------
class OracleLinearClassifier(nn.Module):
    """paule/models.py:887-911 without src_lens: Linear(60->1) per mel frame, mean over time -> one logit per word."""

    def __init__(self, input_dim=60, output_dim=1):
        super().__init__()
        self.linear = nn.Linear(input_dim, output_dim)

    def forward(self, x, *, src_lens=None):
        return self.linear(x).squeeze(2).mean(dim=1)


def plan_inner_loop_branches(pred: nn.Module, emb: nn.Module, cp0: torch.Tensor, target_mel: torch.Tensor,
                             n_steps: int, *, objective: str = "acoustic_semvec", lr: float = 0.01,
                             speech_classifier: Optional[nn.Module] = None, cp_tube_model: Optional[nn.Module] = None,
                             tube_mel_model: Optional[nn.Module] = None, tube_embedder: Optional[nn.Module] = None,
                             target_semvec: Optional[torch.Tensor] = None, log_grads: bool = False) -> Dict[str, object]:
    """Batched inner loop with ONE of the optional branches of paule/paule.py:

    * speech classifier (:210-225, :915, criterion :603-622 / :664-682 / :719-737):
      + 0.1 * BCEWithLogits(classifier(pred_mel), 0), per word;
    * somatosensory feedback (:227-273, :916-931, criterion :624-645; only the acoustic_semvec variant runs in the reference,
      the two others reference undefined names): pred_tube = cp_tube_model(cp), + 5 rmse(tube_mel_model(pred_tube), target_mel)
      + 10 rmse(tube_embedder(pred_tube), target_semvec).  The tube embedder is evaluated without dropout.

    Returns per-step ``loss`` [steps,B], ``terms`` [steps,B,5], ``aux`` [steps,B,3] (classifier, tube_mel, tube_semvec)."""
    B, T, _ = cp0.shape
    lens = tuple(torch.tensor(target_mel.shape[1]) for _ in range(B))
    tube_lens = tuple(torch.tensor(T) for _ in range(B))
    soma = cp_tube_model is not None
    if soma and objective != "acoustic_semvec":
        raise NotImplementedError("the reference's somatosensory criterion only runs for objective='acoustic_semvec'")
    with torch.no_grad():
        tsv = emb(target_mel, lens) if target_semvec is None else target_semvec
    x = cp0.clone().requires_grad_()
    opt = torch.optim.Adam([x], lr=lr)
    out: Dict[str, object] = {"loss": [], "terms": [], "aux": [], "grads": []}
    zero = torch.zeros(B, dtype=cp0.dtype)
    for _ in range(n_steps):
        opt.zero_grad()
        mel = pred(x)
        sv = emb(mel, lens)
        _, terms = per_word_losses(mel, target_mel, sv, tsv, x, objective)
        l_mel, l_sem, l_v, l_j, l_ll = terms.unbind(1)
        l_cls, l_tm, l_ts = zero, zero, zero
        if speech_classifier is not None:
            z = speech_classifier(mel)                                                      # :915
            l_cls = SPEECH_CLASSIFIER_WEIGHT * F.binary_cross_entropy_with_logits(z, torch.zeros_like(z), reduction="none")
            if objective == "acoustic_semvec":
                total = l_mel + l_v + l_j + l_sem + l_cls + l_ll                              # :621
            elif objective == "acoustic":
                total = l_mel + l_v + l_j + l_ll + l_cls                                      # :681
            else:
                total = l_v + l_j + l_sem + l_cls + l_ll                                      # :736
        elif soma:
            tube = cp_tube_model(x)                                                         # :917
            tube_mel = tube_mel_model(tube)                                                 # :919
            tube_sv = tube_embedder(tube, tube_lens)                                        # :929
            l_tm = TUBE_MEL_WEIGHT * _wmean((tube_mel - target_mel) ** 2).sqrt()
            l_ts = TUBE_SEMANTIC_WEIGHT * _wmean((tube_sv - tsv) ** 2).sqrt()
            total = l_mel + l_v + l_j + l_sem + l_ll + l_tm + l_ts                            # :643
        else:
            raise ValueError("give speech_classifier or the three tube models")
        out["loss"].append(total.detach().clone())
        out["terms"].append(terms.detach().clone())
        out["aux"].append(torch.stack((l_cls, l_tm, l_ts), dim=1).detach().clone())
        total.sum().backward()
        if log_grads:
            out["grads"].append(x.grad.detach().clone())
        opt.step()
        with torch.no_grad():
            x.data = x.data.clamp(-CLAMP, CLAMP)
    out["planned_cp"] = x.detach().clone()
    out["target_semvec"] = tsv
    for k in ("loss", "terms", "aux"):
        out[k] = torch.stack(out[k])
    return out


# ----------------------------------------------------------------------------------------------
# independent manual restatement: explicit LSTM cell, hand-derived BPTT, analytic loss gradients
# and hand-written Adam.  These are the formulas the CUDA kernels implement.
# ----------------------------------------------------------------------------------------------
def manual_lstm_forward(x, w_ih, w_hh, b_ih, b_hh):
    """x [B,T,I] -> h [B,T,H] plus stash (i,f,g,o,c); gate order i,f,g,o (torch.nn.LSTM)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    hs, st = [], []
    xp = x @ w_ih.t() + (b_ih + b_hh)
    for t in range(T):
        a = xp[:, t] + h @ w_hh.t()
        i, f, g, o = a.split(H, dim=1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c_prev = c
        c = f * c_prev + i * g
        h = o * torch.tanh(c)
        hs.append(h)
        st.append((i, f, g, o, c, c_prev))
    return torch.stack(hs, 1), st


def manual_lstm_backward_input(dh_out, st, w_ih, w_hh):
    """dh_out [B,T,H] -> dx [B,T,I]; input gradients only (weight gradients are not needed to plan)."""
    B, T, H = dh_out.shape
    dh_rec = dh_out.new_zeros(B, H)
    dc = dh_out.new_zeros(B, H)
    dxs = [None] * T
    for t in reversed(range(T)):
        i, f, g, o, c, c_prev = st[t]
        dh = dh_out[:, t] + dh_rec
        tc = torch.tanh(c)
        do = dh * tc
        dc = dc + dh * o * (1 - tc * tc)
        da = torch.cat((dc * g * i * (1 - i), dc * c_prev * f * (1 - f),
                        dc * i * (1 - g * g), do * o * (1 - o)), dim=1)
        dc = dc * f
        dh_rec = da @ w_hh
        dxs[t] = da @ w_ih
    return torch.stack(dxs, 1)


def manual_smooth_grad(cp):
    """Analytic d/dcp of 80*mean(vel^2) + 400*mean(jerk^2) + 1e5*mean(ll^2) per word (adjoint stencils)."""
    B, T, C = cp.shape

    def d5_adj(r, n_in):                        # adjoint of five_point: r has length n_in-4
        out = r.new_zeros(B, n_in, C)
        out[:, 4:] += -r / 12.0
        out[:, 3:-1] += 8.0 * r / 12.0
        out[:, 1:-3] += -8.0 * r / 12.0
        out[:, :-4] += r / 12.0
        return out

    vel = five_point(cp)
    acc = five_point(vel)
    jerk = five_point(acc)
    ll = local_linear(cp)
    g = d5_adj(vel * (2.0 * VELOCITY_WEIGHT / vel[0].numel()), T)
    gj = jerk * (2.0 * JERK_WEIGHT / jerk[0].numel())
    g = g + d5_adj(d5_adj(d5_adj(gj, T - 8), T - 4), T)
    r = ll * (2.0 * LOCAL_LINEAR_WEIGHT / ll[0].numel())
    gl = cp.new_zeros(B, T, C)
    gl[:, 1:-1] += r
    gl[:, :-2] += -0.5 * r
    gl[:, 2:] += -0.5 * r
    return g + gl


def manual_adam_step(x, g, m, v, step, lr=0.01, b1=0.9, b2=0.999, eps=1e-8):
    """torch/optim/adam.py::_single_tensor_adam (no weight decay, no amsgrad): returns new (x, m, v)."""
    m = m + (g - m) * (1 - b1)                   # exp_avg.lerp_(grad, 1-beta1)
    v = v * b2 + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + eps
    x = x - (lr / bc1) * (m / denom)
    return x, m, v


def manual_step(pw: Dict[str, torch.Tensor], ew: Dict[str, torch.Tensor], cp, target_mel, target_sv,
                objective="acoustic_semvec"):
    """One forward + analytic backward of the whole path; returns (terms[B,5], total[B], dcp[B,T,30], mel, sv)."""
    B, T, _ = cp.shape
    Tm = T // 2
    hf, st_f = manual_lstm_forward(cp, pw["lstm.weight_ih_l0"], pw["lstm.weight_hh_l0"],
                                   pw["lstm.bias_ih_l0"], pw["lstm.bias_hh_l0"])
    hp = 0.5 * (hf[:, 0:2 * Tm:2] + hf[:, 1:2 * Tm:2])
    mel = hp @ pw["post_linear.weight"].t() + pw["post_linear.bias"]
    h0, st_0 = manual_lstm_forward(mel, ew["lstm.weight_ih_l0"], ew["lstm.weight_hh_l0"],
                                   ew["lstm.bias_ih_l0"], ew["lstm.bias_hh_l0"])
    h1, st_1 = manual_lstm_forward(h0, ew["lstm.weight_ih_l1"], ew["lstm.weight_hh_l1"],
                                   ew["lstm.bias_ih_l1"], ew["lstm.bias_hh_l1"])
    sv = h1[:, -1] @ ew["linear_mapping.weight"].t() + ew["linear_mapping.bias"]
    total, terms = per_word_losses(mel, target_mel, sv, target_sv, cp, objective)
    use_mel = objective in ("acoustic_semvec", "acoustic")
    use_sem = objective in ("acoustic_semvec", "semvec")
    e_mel = mel - target_mel
    e_sv = sv - target_sv
    rm = _wmean(e_mel ** 2).sqrt()
    rs = _wmean(e_sv ** 2).sqrt()
    dmel = (MEL_WEIGHT / (e_mel[0].numel() * rm))[:, None, None] * e_mel if use_mel else torch.zeros_like(mel)
    dh0 = torch.zeros_like(h0)
    if use_sem:
        dsv = (SEMANTIC_WEIGHT / (e_sv[0].numel() * rs))[:, None] * e_sv
        dh1 = torch.zeros_like(h1)
        dh1[:, -1] = dsv @ ew["linear_mapping.weight"]
        dh0 = manual_lstm_backward_input(dh1, st_1, ew["lstm.weight_ih_l1"], ew["lstm.weight_hh_l1"])
        dmel = dmel + manual_lstm_backward_input(dh0, st_0, ew["lstm.weight_ih_l0"], ew["lstm.weight_hh_l0"])
    dhp = dmel @ pw["post_linear.weight"]
    dhf = torch.zeros_like(hf)
    dhf[:, 0:2 * Tm:2] = 0.5 * dhp
    dhf[:, 1:2 * Tm:2] = 0.5 * dhp
    dcp = manual_lstm_backward_input(dhf, st_f, pw["lstm.weight_ih_l0"], pw["lstm.weight_hh_l0"])
    dcp = dcp + manual_smooth_grad(cp)
    return terms, total, dcp, mel, sv
