"""``Paule`` -- the reference's planner API (/root/reference/paule/paule.py:92-1551) over the B200 hot path.

Drop-in scope: the gradient-planning inner loop of ``plan_resynth``
(paule.py:910-1211) and the model calls around it (target semvec :533-535, inverse-model init :551-557,
initial/final predictions :822-824, :1460-1464), batched over words.  Same keyword arguments, same
``ValueError``s, same ``PlanningResults`` field list.

Differences, all forced by the scope (BASELINE.json north_star):
* VocalTractLab synthesis and the log-mel front-end stay host-side and are called at outer-loop boundaries only
  (``paule_b200/audio.py``: ctypes binding of the synthesiser the reference ships, librosa-free ``librosa_melspec``); they plug in
  through ``Paule(synthesizer=audio.make_synthesizer(audio.VocalTractLab(...)))``.  Without a synthesizer the ``prod_*`` /
  ``*_sig`` fields are ``None`` (empty lists for the per-step ones) and continue-learning is not run.
* ``target_acoustic`` may be a file name or ``(signal, rate)`` as in the reference, a log-mel array ``[Tm,60]`` (one word,
  results shaped like the reference's), ``[B,Tm,60]`` (a batch: every result gains a leading word axis, per-step losses become
  arrays of shape ``[B]``) or a list of per-word mels of different lengths (ragged batch).
* arithmetic: bf16 tensor-core operands with fp32 accumulation / state / loss / Adam for 720-unit models (``math=None``),
  fp32 kernels with ``math=ops.MATH_FP32`` (the reference ships fp64 CPU).
"""
from __future__ import annotations

import os
import random
from collections import namedtuple
from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from .models import EmbeddingModel, ForwardModel, Generator, InverseModelMelTimeSmoothResidual, LinearClassifier
from .models import math_scope as models_math_scope
from .planner import BatchPlanner

DIR = os.path.dirname(__file__)

# field list of paule/paule.py:57 (33 names)
PlanningResults = namedtuple('PlanningResults', "planned_cp, initial_cp, initial_sig, initial_sr, initial_prod_mel,initial_pred_mel, target_sig, target_sr, target_mel, prod_sig, prod_sr, prod_mel, pred_mel, initial_prod_semvec, initial_pred_semvec, prod_semvec, pred_semvec, prod_loss_steps, planned_loss_steps, planned_mel_loss_steps, vel_loss_steps, jerk_loss_steps, pred_semvec_loss_steps, prod_semvec_loss_steps, cp_steps, pred_semvec_steps, prod_semvec_steps, grad_steps, sig_steps, prod_mel_steps, pred_mel_steps, pred_model_loss, inv_model_loss")
# field lists of paule/paule.py:58-59
PlanningResultsWithSpeechClassifier = namedtuple('PlanningResultsWithSpeechClassifier', "planned_cp, initial_cp, initial_sig, initial_sr, initial_prod_mel, initial_pred_mel, target_sig, target_sr, target_mel, prod_sig, prod_sr, prod_mel, pred_mel, initial_prod_semvec, initial_pred_semvec, prod_semvec, pred_semvec, prod_loss_steps, planned_loss_steps, planned_mel_loss_steps, vel_loss_steps, jerk_loss_steps, pred_semvec_loss_steps, prod_semvec_loss_steps, pred_speech_classifier_loss_steps, prod_speech_classifier_loss_steps, cp_steps, pred_semvec_steps, prod_semvec_steps, grad_steps, sig_steps, prod_mel_steps, pred_mel_steps, pred_model_loss, inv_model_loss")
PlanningResultsWithSomatosensory = namedtuple('PlanningResultsWithSomatosensory', "planned_cp, initial_cp, initial_sig, initial_sr, initial_prod_mel,initial_pred_mel, initial_prod_tube, initial_pred_tube, initial_prod_tube_mel, initial_pred_tube_mel, target_sig, target_sr, target_mel, prod_sig, prod_sr, prod_mel, pred_mel, prod_tube, pred_tube, prod_tube_mel, pred_tube_mel, initial_prod_semvec, initial_pred_semvec, initial_prod_tube_semvec, initial_pred_tube_semvec, prod_semvec, pred_semvec, prod_tube_semvec, pred_tube_semvec, prod_loss_steps, planned_loss_steps, planned_mel_loss_steps, vel_loss_steps, jerk_loss_steps, pred_semvec_loss_steps, prod_semvec_loss_steps, prod_tube_loss_steps, pred_tube_mel_loss_steps,prod_tube_mel_loss_steps, pred_tube_semvec_loss_steps, prod_tube_semvec_loss_steps, cp_steps, pred_semvec_steps, prod_semvec_steps, grad_steps, sig_steps, prod_mel_steps, pred_mel_steps, prod_tube_steps, pred_tube_steps, prod_tube_mel_steps, pred_tube_mel_steps, prod_tube_semvec_steps, pred_tube_semvec_steps, pred_model_loss, inv_model_loss, tube_model_loss, tube_mel_model_loss")
BestSynthesisAcoustic = namedtuple('BestSynthesisAcoustic', "mel_loss, planned_cp, prod_sig, prod_mel, pred_mel")
BestSynthesisSemantic = namedtuple('BestSynthesisSemantic', "semvec_loss, planned_cp, prod_sig, prod_semvec, pred_semvec")
SubLosses = namedtuple('SubLosses', "mel_loss, semvec_loss, velocity_loss, jerk_loss, local_linear_loss, speech_classifier_loss, tube_mel_loss, tube_semvec_loss")

_PRETRAINED = {   # file names of paule/paule.py:126,148,169
    "pred": "pretrained_models/predictive/pred_model_common_voice_1_720_lr_0001_50_00001_50_000001_50_0000001_200.pt",
    "inv": "pretrained_models/inverse/inv_model_common_voice_3_1_720_5_lr_0001_50_00001_50_000001_50_0000001_200.pt",
    "emb": "pretrained_models/embedder/embed_model_common_voice_syn_rec_2_720_0_dropout_07_noise_6e05_rmse_lr_00001_200.pt",
    "cp_gen": "pretrained_models/cp_gan/conditional_trained_cp_generator_whole_critic_it_5_10_20_40_80_100_415.pt",
    "mel_gen": "pretrained_models/mel_gan/conditional_trained_mel_generator_synthesized_critic_it_5_10_20_40_80_100_400.pt",
    # optional branches, paule/paule.py:218,238,250,263
    "cls": "pretrained_models/speech_classifier/linear_model_rec_as_nonspeech.pt",
    "cp_tube": "pretrained_models/somatosensory/cp_to_tube_model_1_360_lr_0001_50_00001_100.pt",
    "tube_mel": "pretrained_models/somatosensory/tube_to_mel_model_1_360_lr_0001_50_00001_100.pt",
    "tube_emb": "pretrained_models/somatosensory/tube_to_vector_model_2_720_0_dropout_07_noise_6e05_rmse_lr_00001_200.pt",
}


def _load_pretrained(module, key, device):
    path = os.path.join(DIR, _PRETRAINED[key])
    if not os.path.exists(path):
        raise FileNotFoundError(f"pretrained weights {path} not found; pass the model explicitly "
                                f"(Paule(pred_model=..., inv_model=..., embedder=...)) or place the reference's "
                                f"pretrained_models/ directory next to {__file__}")
    module.load_state_dict(torch.load(path, map_location=device, weights_only=True))
    return module


def _to_host(t: torch.Tensor) -> np.ndarray:
    """device tensor -> host numpy array through a pinned staging buffer (PCIe at DMA speed instead of the pageable path: a
    1024-word job returns ~250 MB of trajectories and predictions per call)."""
    t = t.detach()
    if not t.is_cuda:
        return t.numpy().copy()
    if t.numel() * t.element_size() < (1 << 20):
        return t.cpu().numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


class Paule():
    """State of the planner: predictive, inverse and embedder model (reference: paule/paule.py:92-318)."""

    def __init__(self, *, pred_model=None, pred_optimizer=None, inv_model=None, inv_optimizer=None,
                 embedder=None, cp_gen_model=None, mel_gen_model=None,
                 use_somatosensory_feedback=False, cp_tube_model=None, tube_optimizer=None,
                 tube_mel_model=None, tube_mel_optimizer=None, tube_embedder=None,
                 continue_data=None, device=torch.device('cuda'), smiling=False,
                 use_speech_classifier=False, speech_classifier=None, speech_classifier_optimizer=None,
                 math=None, synthesizer=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.PauleB200Error("paule_b200.Paule runs on a B200 (device='cuda'); there is no CPU fallback")
        self.smiling = smiling
        # arithmetic of the planning loop: None = the tensor-core path (bf16 operands, fp32 accumulate / state / loss / Adam)
        # for 720-unit models, fp32 kernels otherwise (ops.default_math); ops.MATH_FP32 forces the fp32 parity anchor
        self.math = math
        # host-side stand-in for speak() + librosa_melspec() + normalize_mel_librosa() (paule/util.py:115-146,175-249):
        # callable(cp [T,30] float ndarray) -> normalised log-mel [T//2,60] ndarray.  Only called at outer-loop boundaries.
        self.synthesizer = synthesizer
        if use_somatosensory_feedback and use_speech_classifier:
            raise NotImplementedError("at the moment you have to choose either to use `use_somatosenrosry_feedback=True` OR to use `use_speech_classifier=True` or none")
        self.use_somatosensory_feedback = bool(use_somatosensory_feedback)
        self.use_speech_classifier = bool(use_speech_classifier)

        self.pred_model = pred_model if pred_model else _load_pretrained(
            ForwardModel(num_lstm_layers=1, hidden_size=720), "pred", self.device)
        self.pred_model = self.pred_model.to(self.device)
        self.inv_model = inv_model if inv_model else _load_pretrained(
            InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=720), "inv", self.device)
        self.inv_model = self.inv_model.to(self.device)
        self.embedder = embedder if embedder else _load_pretrained(
            EmbeddingModel(num_lstm_layers=2, hidden_size=720), "emb", self.device)
        self.embedder = self.embedder.to(self.device)
        # generative models (paule.py:186-208): only used in the prologue, for initialize_from='semvec' (:558-565) and for a
        # missing acoustic target (:515-522).  Injected, or the pretrained ones when their files are present; otherwise None
        # and the two prologue modes that need them raise.
        def _gen(given, key, output_size):
            if given is not None:
                return given.to(self.device).float().eval()   # planning is fp32 (the reference ships fp64 modules)
            if os.path.exists(os.path.join(DIR, _PRETRAINED[key])):
                return _load_pretrained(Generator(output_size=output_size), key, self.device).to(self.device).eval()
            return None
        self.cp_gen_model = _gen(cp_gen_model, "cp_gen", 30)
        self.mel_gen_model = _gen(mel_gen_model, "mel_gen", 60)

        # optional loss branches (paule.py:210-273)
        if self.use_speech_classifier:
            self.speech_classifier = speech_classifier if speech_classifier else _load_pretrained(
                LinearClassifier(input_dim=60, output_dim=1), "cls", self.device)
            self.speech_classifier = self.speech_classifier.to(self.device)
            self.speech_classifier.eval()
        if self.use_somatosensory_feedback:
            self.cp_tube_model = cp_tube_model if cp_tube_model else _load_pretrained(
                ForwardModel(num_lstm_layers=1, hidden_size=360, output_size=10, input_size=30,
                             apply_half_sequence=False), "cp_tube", self.device)
            self.tube_mel_model = tube_mel_model if tube_mel_model else _load_pretrained(
                ForwardModel(num_lstm_layers=1, hidden_size=360, output_size=60, input_size=10,
                             apply_half_sequence=True), "tube_mel", self.device)
            self.tube_embedder = tube_embedder if tube_embedder else _load_pretrained(
                EmbeddingModel(input_size=10, num_lstm_layers=2, hidden_size=720, dropout=0.7, post_upsampling_size=0),
                "tube_emb", self.device)
            self.cp_tube_model = self.cp_tube_model.to(self.device)
            self.tube_mel_model = self.tube_mel_model.to(self.device)
            self.tube_embedder = self.tube_embedder.to(self.device)
            self.tube_embedder.eval()

        self.continue_data = continue_data
        self.continue_data_limit = 1000
        self.pred_optimizer = pred_optimizer if pred_optimizer else torch.optim.Adam(self.pred_model.parameters(), lr=0.001)
        self.inv_optimizer = inv_optimizer if inv_optimizer else torch.optim.Adam(self.inv_model.parameters(), lr=0.001)
        self.best_synthesis_acoustic = None
        self.best_synthesis_semantic = None
        self.last_planner: Optional[BatchPlanner] = None
        self._planner_key = None      # plan_resynth in a loop re-arms the planner of the previous call (same shapes / options)

    # ---- host-side synthesis pipeline (reference: speak() + librosa_melspec() + normalize, paule.py:1097-1113)
    def _submit_synthesis(self, cps_np):
        """One synthesis job per word on a thread pool (VocalTractLab runs behind ctypes and releases the GIL).  Returns the
        futures; the GPU keeps planning while they run."""
        import concurrent.futures
        import os
        if getattr(self, "_synth_pool", None) is None:
            self._synth_pool = concurrent.futures.ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1))
        return [self._synth_pool.submit(self.synthesizer, c) for c in cps_np]

    @staticmethod
    def _synthesis_result(fut):
        """synthesizer may return the normalised log-mel [Tm,60] or (sig, sr, mel)."""
        r = fut.result()
        if isinstance(r, tuple) and len(r) == 3:
            return r[0], r[1], np.asarray(r[2], dtype=np.float32)
        return None, None, np.asarray(r, dtype=np.float32)

    def _produced_metrics(self, prod_mels, target_mels, target_semvec):
        """prod_loss = 5 rmse(prod_mel, target_mel) (paule.py:1109-1111), prod_semvec = embedder(prod_mel),
        prod_semvec_loss = 10 rmse(prod_semvec, target_semvec) (:1139-1146), per word.  Tiny, runs on the GPU with torch ops."""
        B = len(prod_mels)
        lens = [int(m.shape[0]) for m in prod_mels]
        pad = torch.zeros((B, max(lens), prod_mels[0].shape[1]), device=self.device)
        for b, m in enumerate(prod_mels):
            pad[b, :lens[b]] = torch.from_numpy(m).to(self.device)
        mel_loss = torch.stack([5.0 * torch.sqrt(torch.mean((pad[b, :lens[b]] - target_mels[b][:lens[b]]) ** 2)) for b in range(B)])
        with torch.no_grad():
            sv = self.embedder(pad, lens)
        sem_loss = 10.0 * torch.sqrt(torch.mean((sv - target_semvec) ** 2, dim=1))
        return mel_loss.cpu().numpy(), sv.cpu().numpy(), sem_loss.cpu().numpy()

    def continue_learning_pred(self, cps, mels, *, n_epochs=10, batch_size=8, shuffle=True):
        """Continue-learning of the predictive forward model on (cp, produced mel) pairs: the learning step of the outer
        loop (paule/paule.py:1361-1377) -- same-size batching (create_epoch_batches(same_size_batching=True), :349-369),
        ``Y_hat = pred_model(batch)``, RMSE criterion (:68), ``self.pred_optimizer`` -- on the GPU: forward and input-gradient
        BPTT on the CUDA LSTM kernels, weight gradients as GEMMs over the saved operands.  ``cps[i]`` is [T_i,30], ``mels[i]``
        [T_i // 2, 60].  Returns the mean loss of every epoch; the planner of the last plan_resynth call is repacked."""
        if len(cps) != len(mels):
            raise ValueError("cps and mels need the same number of samples")
        xs = [torch.as_tensor(np.ascontiguousarray(c)).float() for c in cps]
        ys = [torch.as_tensor(np.ascontiguousarray(m)).float() for m in mels]
        by_len = {}
        for i, x in enumerate(xs):
            if ys[i].shape[0] != x.shape[0] // 2:
                raise ValueError(f"sample {i}: {x.shape[0]} cp frames need {x.shape[0] // 2} mel frames, got {ys[i].shape[0]}")
            by_len.setdefault(int(x.shape[0]), []).append(i)
        was_training = self.pred_model.training
        self.pred_model.train()
        epoch_losses = []
        for _ in range(int(n_epochs)):
            batches = []
            for length in sorted(by_len):                       # same-size batching (:349-369)
                idx = list(by_len[length])
                if shuffle:
                    random.shuffle(idx)
                batches += [idx[k:k + batch_size] for k in range(0, len(idx), batch_size)]
            if shuffle:
                random.shuffle(batches)
            losses = []
            for j in batches:
                batch_input = torch.stack([xs[i] for i in j]).to(self.device)
                batch_output = torch.stack([ys[i] for i in j]).to(self.device)
                y_hat = self.pred_model(batch_input)                         # :1372
                self.pred_optimizer.zero_grad()                              # :1374
                pred_loss = torch.sqrt(torch.mean((y_hat - batch_output) ** 2))   # RMSELoss(eps=0), paule/util.py:570-572
                pred_loss.backward()                                         # :1376
                self.pred_optimizer.step()                                   # :1377
                losses.append(float(pred_loss.item()))
            epoch_losses.append(float(np.mean(losses)) if losses else float("nan"))
        self.pred_model.train(was_training)
        if self.last_planner is not None:
            self.last_planner.refresh_weights()      # repack: the planner holds bf16 / transposed copies of the weights
        return epoch_losses

    @staticmethod
    def _same_size_batches(lengths, batch_size, shuffle):
        """create_epoch_batches(same_size_batching=True) of the reference (paule/paule.py:349-369): batches of samples of one
        length, shuffled inside a length and across batches."""
        by_len = {}
        for i, n in enumerate(lengths):
            by_len.setdefault(int(n), []).append(i)
        batches = []
        for length in sorted(by_len):
            idx = list(by_len[length])
            if shuffle:
                random.shuffle(idx)
            batches += [idx[k:k + batch_size] for k in range(0, len(idx), batch_size)]
        if shuffle:
            random.shuffle(batches)
        return batches

    @staticmethod
    def cp_trajectory_loss(y_hat, tgts):
        """paule/util.py:640-671: RMSE of position, velocity, acceleration and jerk (five-point stencils; the deprecated ``lag``
        argument is ignored by the reference, so every derivative term counts three times).  Returns (loss, pos, vel, acc, jerk)."""
        def d5(v):
            return (-v[:, 4:] + 8.0 * v[:, 3:-1] - 8.0 * v[:, 1:-3] + v[:, :-4]) / 12.0

        def rmse(a, b):
            return torch.sqrt(torch.mean((a - b) ** 2))
        v_t = d5(tgts); a_t = d5(v_t); j_t = d5(a_t)
        v_h = d5(y_hat); a_h = d5(v_h); j_h = d5(a_h)
        pos = rmse(y_hat, tgts)
        vel, acc, jerk = 3.0 * rmse(v_h, v_t), 3.0 * rmse(a_h, a_t), 3.0 * rmse(j_h, j_t)
        return pos + vel + acc + jerk, pos, vel, acc, jerk

    def continue_learning_inv(self, mels, cps, *, n_epochs=10, batch_size=8, shuffle=True):
        """Continue-learning of the inverse model on (produced mel -> cp) pairs (paule/paule.py:1413-1436): same-size batching,
        ``Y_hat = inv_model(batch)``, ``cp_trajectory_loss`` (:294), ``self.inv_optimizer``.  ``mels[i]`` is [Tm_i, 60],
        ``cps[i]`` [2 Tm_i, 30].  Returns the mean loss of every epoch."""
        if len(cps) != len(mels):
            raise ValueError("cps and mels need the same number of samples")
        xs = [torch.as_tensor(np.ascontiguousarray(m)).float() for m in mels]
        ys = [torch.as_tensor(np.ascontiguousarray(c)).float() for c in cps]
        for i, (x, y) in enumerate(zip(xs, ys)):
            if y.shape[0] != 2 * x.shape[0]:
                raise ValueError(f"sample {i}: {x.shape[0]} mel frames need {2 * x.shape[0]} cp frames, got {y.shape[0]}")
        was_training = self.inv_model.training
        self.inv_model.train()
        self.inv_model.learnable = True          # the differentiable forward (models.InverseModel...forward)
        for p in self.inv_model.parameters():
            p.requires_grad_(True)
        epoch_losses = []
        for _ in range(int(n_epochs)):
            losses = []
            for j in self._same_size_batches([x.shape[0] for x in xs], batch_size, shuffle):
                batch_input = torch.stack([xs[i] for i in j]).to(self.device)
                batch_output = torch.stack([ys[i] for i in j]).to(self.device)
                y_hat = self.inv_model(batch_input)                         # :1427
                self.inv_optimizer.zero_grad()                              # :1429
                inv_loss = self.cp_trajectory_loss(y_hat, batch_output)[0]  # :1430-1432
                inv_loss.backward()                                         # :1433
                self.inv_optimizer.step()                                   # :1434
                losses.append(float(inv_loss.item()))
            epoch_losses.append(float(np.mean(losses)) if losses else float("nan"))
        self.inv_model.train(was_training)
        self.inv_model.learnable = False
        return epoch_losses

    def continue_learning_tube(self, cps, tubes, mels, *, n_epochs=10, batch_size=8, shuffle=True):
        """Continue-learning of the somatosensory models (paule/paule.py:1381-1405): ``cp_tube_model`` on (cp -> produced tube)
        and ``tube_mel_model`` on (produced tube -> produced mel), both with the RMSE criterion (:301-302) and their own
        optimizers.  ``tubes[i]`` is [T_i, 10] (VocalTractLab's tube extraction, host side).  Returns (tube_losses, tube_mel_losses)."""
        if not self.use_somatosensory_feedback:
            raise ValueError("continue_learning_tube needs use_somatosensory_feedback=True")
        if not (len(cps) == len(tubes) == len(mels)):
            raise ValueError("cps, tubes and mels need the same number of samples")
        xs = [torch.as_tensor(np.ascontiguousarray(c)).float() for c in cps]
        ts = [torch.as_tensor(np.ascontiguousarray(t)).float() for t in tubes]
        ms = [torch.as_tensor(np.ascontiguousarray(m)).float() for m in mels]
        if getattr(self, "tube_optimizer", None) is None:
            self.tube_optimizer = torch.optim.Adam(self.cp_tube_model.parameters(), lr=0.001)          # paule.py:296-298
        if getattr(self, "tube_mel_optimizer", None) is None:
            self.tube_mel_optimizer = torch.optim.Adam(self.tube_mel_model.parameters(), lr=0.001)
        for m_ in (self.cp_tube_model, self.tube_mel_model):
            m_.train()
            for p in m_.parameters():
                p.requires_grad_(True)

        def rmse(a, b):
            return torch.sqrt(torch.mean((a - b) ** 2))
        tube_losses, tube_mel_losses = [], []
        for _ in range(int(n_epochs)):
            l1, l2 = [], []
            for j in self._same_size_batches([x.shape[0] for x in xs], batch_size, shuffle):
                batch_input = torch.stack([xs[i] for i in j]).to(self.device)
                batch_tube = torch.stack([ts[i] for i in j]).to(self.device)
                batch_mel = torch.stack([ms[i] for i in j]).to(self.device)
                y_hat = self.cp_tube_model(batch_input)                     # :1388
                self.tube_optimizer.zero_grad()
                tube_loss = rmse(y_hat, batch_tube)
                tube_loss.backward()
                self.tube_optimizer.step()
                l1.append(float(tube_loss.item()))
                y_hat = self.tube_mel_model(batch_tube)                     # :1397
                self.tube_mel_optimizer.zero_grad()
                tube_mel_loss = rmse(y_hat, batch_mel)
                tube_mel_loss.backward()
                self.tube_mel_optimizer.step()
                l2.append(float(tube_mel_loss.item()))
            tube_losses.append(float(np.mean(l1)) if l1 else float("nan"))
            tube_mel_losses.append(float(np.mean(l2)) if l2 else float("nan"))
        return tube_losses, tube_mel_losses

    def plan_iterative(self, overlap=8):
        """Empty stub in the reference as well (paule/paule.py:383-388)."""
        pass

    def plan_resynth(self, *, learning_rate_planning=0.01, learning_rate_learning=0.001,
                     learning_rate_learning_inv=None,
                     target_acoustic=None,
                     target_semvec=None,
                     target_seq_length=None,
                     initial_cp=None,
                     past_cp=None,
                     initialize_from="acoustic",
                     objective="acoustic",
                     n_outer=5, n_inner=24,
                     continue_learning=True,
                     continue_learning_inv=False,
                     continue_learning_tube=False,
                     add_training_data_pred=False,
                     add_training_data_inv=False,
                     n_batches=3, batch_size=8, n_epochs=10,
                     log_ii=1,
                     log_semantics=True,
                     log_gradients=False,
                     log_signals=False,
                     log_cps=False,
                     plot=False,
                     seed=None,
                     verbose=True):
        """Plans resynthesis cp trajectories (reference: paule/paule.py:391-1551) for one word or a batch."""
        if seed:
            torch.manual_seed(seed)
            random.seed(seed)

        if target_acoustic is None and target_semvec is None:
            raise ValueError("Either target_acoustic or target_semvec has to be not None.")

        if learning_rate_learning:
            for param_group in self.pred_optimizer.param_groups:
                param_group['lr'] = learning_rate_learning
        if learning_rate_learning_inv:
            for param_group in self.inv_optimizer.param_groups:
                param_group['lr'] = learning_rate_learning_inv

        if log_ii is None:
            log_ii = n_inner
        if log_ii > n_inner:
            raise ValueError('results can only be logged between first and last planning step')

        # ---- ragged batches (new): a list of per-word mel arrays [Tm_b, 60] of different lengths.  The words
        # are padded to the longest one and planned in lock-step, each exactly as if it were alone.
        lengths = None
        if (isinstance(target_acoustic, (list, tuple)) and len(target_acoustic) > 0
                and all(isinstance(m, (np.ndarray, torch.Tensor)) and np.ndim(m) == 2 for m in target_acoustic)
                and len({int(np.shape(m)[0]) for m in target_acoustic}) > 1):
            mels = [torch.as_tensor(np.ascontiguousarray(m) if isinstance(m, np.ndarray) else m).float() for m in target_acoustic]
            lengths = [2 * int(m.shape[0]) for m in mels]
            if past_cp is not None:
                raise NotImplementedError("past_cp with ragged batches")
            if initialize_from == "semvec":
                raise NotImplementedError("initialize_from='semvec' with ragged batches")
            if initial_cp is None:   # the inverse model is not causal: initialise every word on its own (paule.py:551-557)
                with torch.no_grad():
                    inits = [self.inv_model(m.unsqueeze(0).to(self.device)).clamp(min=-1, max=1)[0] for m in mels]
                initialize_from = None
            else:
                if initialize_from is not None:
                    raise ValueError('one of initial_cp and initialize_from has to be None')
                inits = [torch.as_tensor(np.ascontiguousarray(c) if isinstance(c, np.ndarray) else c).float() for c in initial_cp]
                if len(inits) != len(mels) or any(c.shape[0] != L for c, L in zip(inits, lengths)):
                    raise ValueError("initial_cp needs one [2 * mel frames, 30] array per word")
            Tm_max = max(m.shape[0] for m in mels)
            target_acoustic = torch.zeros((len(mels), Tm_max, mels[0].shape[1]))
            initial_cp = torch.zeros((len(mels), 2 * Tm_max, inits[0].shape[1]))
            for b, (m, c) in enumerate(zip(mels, inits)):
                target_acoustic[b, :m.shape[0]] = m
                initial_cp[b, :c.shape[0]] = c.to(initial_cp.device)

        # ---- target parsing (paule.py:486-529)
        batched = False
        target_mel = None
        self._target_audio = (None, None)
        if isinstance(target_acoustic, str) or (isinstance(target_acoustic, (tuple, list)) and len(target_acoustic) == 2
                                                and not isinstance(target_acoustic[0], (list, tuple, torch.Tensor))
                                                and np.ndim(target_acoustic[0]) in (1, 2) and np.ndim(target_acoustic[1]) == 0):
            # audio target: a file name or (signal, sampling rate) -> normalised log-mel, shifted to min 0 (paule.py:487-496,
            # :523-529); the mel front-end runs on the host (paule_b200/audio.py), once per call
            from . import audio
            sig, sr = audio.read_audio(target_acoustic) if isinstance(target_acoustic, str) else target_acoustic
            sig = np.asarray(sig, dtype=np.float64)
            if sig.ndim == 2:
                sig = sig.mean(axis=1)                                         # stereo_to_mono, paule.py:489-490
            self._target_audio = (sig, int(sr))
            target_mel = torch.from_numpy(audio.target_mel_from_audio(sig, int(sr)).astype(np.float32)).unsqueeze(0)
            target_seq_length = target_mel.shape[1]
        elif target_acoustic is None:
            pass
        else:
            if isinstance(target_acoustic, torch.Tensor):
                target_mel = target_acoustic.detach().clone()
            else:
                target_mel = torch.from_numpy(np.ascontiguousarray(target_acoustic))
            if target_mel.dim() == 2:
                target_mel = target_mel.unsqueeze(0)
            elif target_mel.dim() == 3:
                batched = target_mel.shape[0] != 1 or (initial_cp is not None and np.ndim(initial_cp) == 3) or lengths is not None
            else:
                raise ValueError("target_acoustic has to be torch.Tensor at this point")
            target_seq_length = target_mel.shape[1]

        if target_acoustic is None and (target_seq_length is None or target_semvec is None):
            raise ValueError("if target_acoustic is None you need to give a target_seq_length and a target_semvec")
        elif target_acoustic is None:
            if self.mel_gen_model is None:
                raise NotImplementedError("planning without an acoustic target needs mel_gen_model (paule.py:515-522)")
            if not isinstance(target_semvec, torch.Tensor):
                target_semvec = torch.tensor(np.asarray(target_semvec), device=self.device)
            sv = target_semvec.reshape(-1, 300).detach().clone().to(self.device).float()
            noise = torch.randn(sv.shape[0], 1, 100, device=self.device)
            with torch.no_grad():
                target_mel = self.mel_gen_model(noise, target_seq_length, sv).detach().clone()
            batched = sv.shape[0] != 1

        target_mel = target_mel.to(self.device).float().contiguous()
        B = target_mel.shape[0]

        if target_semvec is not None:
            if not isinstance(target_semvec, torch.Tensor):
                target_semvec = torch.tensor(np.asarray(target_semvec), device=self.device)
            target_semvec = target_semvec.reshape(B, 300).detach().clone().to(self.device).float()

        # ---- 1.1 initial cp (paule.py:550-573)
        if initial_cp is None:
            if initialize_from == "acoustic":
                # the inverse model's recurrence runs in the planner's arithmetic (tensor-core path for 720 units) unless the
                # module carries its own `math` attribute
                inv_math = self.math if self.math is not None else ops.default_math(self.inv_model.lstm.hidden_size)
                with torch.no_grad(), models_math_scope(inv_math):
                    cp0 = self.inv_model(target_mel).clamp(min=-1, max=1)
            elif initialize_from == "semvec":
                if self.cp_gen_model is None:
                    raise NotImplementedError("initialize_from='semvec' needs cp_gen_model (paule.py:558-565)")
                if target_semvec is None:
                    with torch.no_grad():
                        target_semvec = self.embedder(target_mel, tuple(target_mel.shape[1] for _ in range(B)))
                noise = torch.randn(B, 1, 100, device=self.device)
                with torch.no_grad():
                    cp0 = self.cp_gen_model(noise, 2 * target_seq_length, target_semvec.reshape(B, 300)).detach().float()
            else:
                raise ValueError("initialize_from has to be either 'acoustic' or 'semvec'")
        else:
            if initialize_from is not None:
                raise ValueError('one of initial_cp and initialize_from has to be None')
            cp0 = initial_cp if isinstance(initial_cp, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(initial_cp))
            if cp0.dim() == 2:
                cp0 = cp0.unsqueeze(0)
            if not cp0.shape[1] == (target_mel.shape[1] * 2):
                raise ValueError(f"initial_cp {cp0.shape[1]}, target_mel {target_mel.shape[1] * 2}")
            if cp0.shape[0] != B:
                raise ValueError(f"initial_cp has {cp0.shape[0]} words, target_acoustic {B}")
            cp0 = cp0.to(self.device).float()

        if not past_cp is None and past_cp.shape[-2] % 2 != 0:
            raise ValueError("past_cp have to be None or the sequence length has to be an even number")
        if objective not in ('acoustic_semvec', 'acoustic', 'semvec'):
            raise ValueError("objective has to be one of 'acoustic_semvec', 'acoustic' or 'semvec'")

        past_t = None
        if past_cp is not None:
            past_t = past_cp if isinstance(past_cp, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(past_cp))
            past_t = past_t.to(self.device).float()
            if past_t.dim() == 2:
                past_t = past_t.unsqueeze(0).expand(B, -1, -1)
            cp0 = torch.cat((past_t, cp0), dim=1)                                   # paule.py:579
            # the reference prepends the *produced* mel of the past cps to the target (paule.py:867-870); without
            # VocalTractLab the predicted mel of the past cps is used instead
            with torch.no_grad():
                past_mel = self.pred_model(past_t.contiguous())
            target_mel = torch.cat((past_mel, target_mel), dim=1).contiguous()
        initial_cp_np = _to_host(cp0)

        n_steps = int(n_outer) * int(n_inner)
        with torch.cuda.device(self.device):
            planner = self._get_planner(cp0, target_mel, target_semvec, learning_rate_planning, objective, past_t,
                                        log_semantics, log_gradients, n_steps, lengths)
            return self._plan(planner, cp0, target_mel, initial_cp_np, lengths, batched, B, n_outer, n_inner, n_steps, log_ii,
                              log_cps, log_gradients, log_signals, log_semantics, objective, continue_learning,
                              add_training_data_pred, n_epochs, batch_size, continue_learning_inv)

    def _get_planner(self, cp0, target_mel, target_semvec, lr, objective, past_t, log_semantics, log_gradients, n_steps,
                     lengths) -> BatchPlanner:
        """The planner of this call: the previous call's one re-armed (workspace, packed weights and CUDA graph kept) when
        shapes, options and weights are unchanged -- ``plan_resynth`` is normally called in a loop over words -- else a new
        one (the old one's device memory is released first)."""
        from .models import _versions
        soma = self.use_somatosensory_feedback
        weights = _versions(list(self.pred_model.parameters()) + list(self.embedder.parameters())
                            + (list(self.speech_classifier.parameters()) if self.use_speech_classifier else []))
        key = (tuple(cp0.shape), tuple(target_mel.shape), float(lr), objective, bool(self.smiling),
               None if past_t is None else tuple(past_t.shape), bool(log_semantics), bool(log_gradients),
               lengths is not None, self.math, self.use_speech_classifier, weights)
        old = self.last_planner
        if (old is not None and not soma and self._planner_key == key and old.workspace is not None
                and old.max_log_steps >= max(n_steps, 1)):
            old.reset(cp0, target_mel, target_semvec, past_cp=past_t, lengths=lengths)
            return old
        if old is not None:
            old.close()
            self.last_planner = None
        planner = BatchPlanner(self.pred_model, self.embedder, cp0, target_mel, target_semvec,
                               lr=lr, objective=objective, smiling=self.smiling, past_cp=past_t,
                               log_semantics=log_semantics, log_gradients=log_gradients,
                               max_log_steps=max(n_steps, 1), math=self.math,
                               use_cuda_graph=True, lengths=lengths,
                               speech_classifier=self.speech_classifier if self.use_speech_classifier else None,
                               somatosensory=(self.cp_tube_model, self.tube_mel_model, self.tube_embedder)
                               if soma else None)
        self.last_planner = planner
        self._planner_key = key
        return planner

    def _plan(self, planner, cp0, target_mel, initial_cp_np, lengths, batched, B, n_outer, n_inner, n_steps, log_ii, log_cps,
              log_gradients, log_signals, log_semantics, objective, continue_learning, add_training_data_pred, n_epochs,
              batch_size, learn_inv=False):
        """Outer / inner loops, final predictions and result packing (paule/paule.py:822-1550) on an armed planner."""
        soma = planner.soma

        def tube_predictions():       # no_grad cp -> tube -> (mel, semvec) of the current cps (paule.py:1472-1487)
            with torch.no_grad():
                tube, tube_mel, tube_sv = soma.forward(planner.cp)
            return ops.transpose_btc(tube), ops.transpose_btc(tube_mel), tube_sv
        if soma is not None:
            initial_pred_tube, initial_pred_tube_mel, initial_pred_tube_semvec = tube_predictions()

        def out(t):
            a = _to_host(t)
            if lengths is not None and a.ndim == 3:   # ragged: per-word arrays without the padding frames
                per = 1 if a.shape[1] == cp0.shape[1] else 2
                return [a[b, :L // per] for b, L in enumerate(lengths)]
            return a if batched else a[0]

        # initial predictions (paule.py:822-824).  The forward pass of the FIRST inner step is exactly this prediction (the
        # step logs before it updates), so it is taken from there instead of running the models once more
        sem_in_step = objective != "acoustic" or log_semantics
        first_from_step = (soma is None and sem_in_step and n_steps >= 1 and not (log_cps or log_gradients)
                           and self.synthesizer is None and planner.steps_done == 0)
        if first_from_step:
            planner.step(1)
            initial_pred_mel, initial_pred_semvec = ops.transpose_btc(planner.pred_mel), planner.pred_sv.clone()
        else:
            initial_pred_mel, initial_pred_semvec = planner.forward()

        def word_cps():
            cur = _to_host(planner.planned_cp())
            return [cur[b, :L] for b, L in enumerate(lengths)] if lengths is not None else [cur[b] for b in range(cur.shape[0])]

        # produced side (paule.py:826-870, :1097-1164): synthesised on the host at the outer-loop boundaries, overlapped with
        # the GPU's next outer iteration; pending = [(outer index or -1 for the initial cps, cps, futures)]
        synth = self.synthesizer is not None
        pending, produced = [], []
        if synth:
            cps_init = word_cps()
            pending.append((-1, cps_init, self._submit_synthesis(cps_init)))

        cp_steps, grad_steps, pred_semvec_steps, pred_mel_steps = [], [], [], []
        pred_model_loss, inv_model_loss = [], []
        need_per_step = log_cps or log_gradients
        for ii_outer in range(n_outer):
            cp_steps_ii, pred_semvec_steps_ii, pred_mel_steps_ii = [], [], []
            if not need_per_step:
                # no host round trip inside the inner loop (the first step may already have run, see above)
                planner.step(n_inner - (1 if (first_from_step and ii_outer == 0) else 0))
            else:
                for ii in range(n_inner):
                    if log_cps and (ii + 1) % log_ii == 0:
                        cp_steps_ii.append(out(planner.planned_cp()))                # logged before the update (:1066)
                    planner.step(1)
                    if log_gradients:
                        grad_steps.append(planner.last_grad().clone())              # :1062-1063
            cp_steps.append(cp_steps_ii)
            pred_semvec_steps.append(pred_semvec_steps_ii)
            pred_mel_steps.append(pred_mel_steps_ii)
            # continue-learning (paule.py:1243-1454): the planned cps of this outer iteration are synthesised on the host
            # (VocalTractLab + librosa behind `synthesizer`), the produced mels train pred_model, the planner is repacked.
            # Without a synthesizer the models stay frozen (the reference needs VocalTractLab here).
            if synth:
                cur_list = word_cps()
                pending.append((ii_outer, cur_list, self._submit_synthesis(cur_list)))   # runs while the GPU plans on
            if continue_learning and synth:
                prod = [self._synthesis_result(f)[2] for f in pending[-1][2]]         # the learning step needs them now
                train_cps, train_mels = list(cur_list), list(prod)
                if add_training_data_pred and self.continue_data is not None and len(self.continue_data) > 0:
                    k = min(len(self.continue_data), len(cur_list))      # 50 % known data, 50 % produced (:1256-1268)
                    for i in random.sample(range(len(self.continue_data)), k=k):
                        row = self.continue_data.iloc[i] if hasattr(self.continue_data, "iloc") else self.continue_data[i]
                        train_cps.append(np.asarray(row["cp_norm"], dtype=np.float32))
                        train_mels.append(np.asarray(row["melspec_norm_synthesized"], dtype=np.float32))
                pred_model_loss += self.continue_learning_pred(train_cps, train_mels, n_epochs=n_epochs, batch_size=batch_size)
                if learn_inv:      # paule.py:1413-1436: the inverse model learns (produced mel -> cp) on the same samples
                    inv_model_loss += self.continue_learning_inv(train_mels, train_cps, n_epochs=n_epochs, batch_size=batch_size)

        # final predictions (paule.py:1456-1470)
        planned_cp = planner.planned_cp()
        pred_mel, pred_semvec = planner.forward()

        # collect the produced side: losses and semvecs per synthesis, best synthesis per word (:1158-1164)
        prod_loss_steps, prod_semvec_loss_steps, prod_mel_steps, prod_semvec_steps, sig_steps = [], [], [], [], []
        initial_prod_mel = initial_prod_semvec = prod_mel_out = prod_semvec_out = prod_sig = prod_sr = None
        if synth:
            tm = planner.target_mel.transpose(0, 1)                      # [B,Tm,60]
            t_mels = [tm[b] for b in range(B)]
            for idx, cps_k, futs in pending:
                res = [self._synthesis_result(f) for f in futs]
                mels_k = [r[2] for r in res]
                mel_loss, sv, sem_loss = self._produced_metrics(mels_k, t_mels, planner.target_sv)
                wrap = (lambda v: v) if batched else (lambda v: v[0])
                if idx < 0:
                    initial_prod_mel, initial_prod_semvec = wrap(mels_k), wrap(sv)
                    continue
                prod_loss_steps.append(wrap(mel_loss)); prod_semvec_loss_steps.append(wrap(sem_loss))
                prod_mel_steps.append(wrap(mels_k)); prod_semvec_steps.append(wrap(sv))
                if log_signals:
                    sig_steps.append(wrap([r[0] for r in res]))
                prod_mel_out, prod_semvec_out = wrap(mels_k), wrap(sv)
                prod_sig, prod_sr = wrap([r[0] for r in res]), res[0][1]
                for b in range(B):      # best synthesis so far, per word
                    if self.best_synthesis_acoustic is None:
                        self.best_synthesis_acoustic = [None] * B
                        self.best_synthesis_semantic = [None] * B
                    if len(self.best_synthesis_acoustic) != B:
                        self.best_synthesis_acoustic, self.best_synthesis_semantic = [None] * B, [None] * B
                    ba, bs = self.best_synthesis_acoustic[b], self.best_synthesis_semantic[b]
                    if ba is None or ba.mel_loss > float(mel_loss[b]):
                        self.best_synthesis_acoustic[b] = BestSynthesisAcoustic(float(mel_loss[b]), cps_k[b], res[b][0], mels_k[b], None)
                    if bs is None or bs.semvec_loss > float(sem_loss[b]):
                        self.best_synthesis_semantic[b] = BestSynthesisSemantic(float(sem_loss[b]), cps_k[b], res[b][0], sv[b], None)
        logs = {k: _to_host(v.contiguous()) for k, v in planner.losses().items()}   # the device-resident loss log, read once

        def per_step(name):
            rows = [logs[name][k] for k in range(n_steps) if (k % n_inner + 1) % log_ii == 0]
            return [r if batched else float(r[0]) for r in rows]

        sem_logged = objective in ('acoustic_semvec', 'semvec') or log_semantics
        sem_steps = per_step("semvec") if sem_logged else list()
        if self.use_speech_classifier:
            # produced side of the classifier term (paule.py:1116-1125): 0.1 BCEWithLogits(classifier(prod_mel), 0) per synthesis
            prod_cls = []
            if synth:
                for mels_k in prod_mel_steps:
                    lst = mels_k if batched else [mels_k]
                    with torch.no_grad():
                        z = torch.stack([self.speech_classifier(torch.from_numpy(np.asarray(m, dtype=np.float32))
                                                                .to(self.device).unsqueeze(0))[0] for m in lst])
                    v = (0.1 * torch.nn.functional.softplus(z)).cpu().numpy()
                    prod_cls.append(v if batched else float(v[0]))
            return PlanningResultsWithSpeechClassifier(
                out(planned_cp), (out(cp0) if lengths is not None else initial_cp_np) if batched else initial_cp_np[0], None,
                None, initial_prod_mel, out(initial_pred_mel), None, None, out(target_mel), prod_sig, prod_sr, prod_mel_out,
                out(pred_mel), initial_prod_semvec, out(initial_pred_semvec), prod_semvec_out, out(pred_semvec),
                prod_loss_steps, per_step("total"), per_step("mel"), per_step("velocity"), per_step("jerk"), sem_steps,
                prod_semvec_loss_steps, per_step("speech_classifier"), prod_cls, cp_steps, pred_semvec_steps,
                prod_semvec_steps, grad_steps, sig_steps, prod_mel_steps, pred_mel_steps, pred_model_loss, inv_model_loss)
        if self.use_somatosensory_feedback:
            # the produced tube side needs VocalTractLab's tube extraction (paule.py:1070-1095), which stays host-side and does
            # not ship: those fields are None / empty
            pred_tube, pred_tube_mel, pred_tube_semvec = tube_predictions()
            return PlanningResultsWithSomatosensory(
                out(planned_cp), (out(cp0) if lengths is not None else initial_cp_np) if batched else initial_cp_np[0], None,
                None, initial_prod_mel, out(initial_pred_mel), None, out(initial_pred_tube), None, out(initial_pred_tube_mel),
                None, None, out(target_mel), prod_sig, prod_sr, prod_mel_out, out(pred_mel), None, out(pred_tube), None,
                out(pred_tube_mel), initial_prod_semvec, out(initial_pred_semvec), None, out(initial_pred_tube_semvec),
                prod_semvec_out, out(pred_semvec), None, out(pred_tube_semvec),
                prod_loss_steps, per_step("total"), per_step("mel"), per_step("velocity"), per_step("jerk"), sem_steps,
                prod_semvec_loss_steps, list(), per_step("tube_mel"), list(), per_step("tube_semvec"), list(), cp_steps,
                pred_semvec_steps, prod_semvec_steps, grad_steps, sig_steps, prod_mel_steps, pred_mel_steps, list(), list(),
                list(), list(), list(), list(), pred_model_loss, inv_model_loss, list(), list())
        target_sig, target_sr = getattr(self, "_target_audio", (None, None))
        return PlanningResults(
            out(planned_cp), (out(cp0) if lengths is not None else initial_cp_np) if batched else initial_cp_np[0], None, None,
            initial_prod_mel, out(initial_pred_mel),
            target_sig, target_sr, out(target_mel), prod_sig, prod_sr, prod_mel_out, out(pred_mel), initial_prod_semvec,
            out(initial_pred_semvec), prod_semvec_out,
            out(pred_semvec), prod_loss_steps, per_step("total"), per_step("mel"), per_step("velocity"), per_step("jerk"),
            per_step("semvec") if sem_logged else list(), prod_semvec_loss_steps, cp_steps, pred_semvec_steps, prod_semvec_steps,
            grad_steps, sig_steps, prod_mel_steps, pred_mel_steps, pred_model_loss, inv_model_loss)


# BASELINE.json's north_star spells the class name in capitals
PAULE = Paule
