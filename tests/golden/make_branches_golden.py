#!/usr/bin/env python3
"""Golden vectors for the optional loss branches FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_branches_golden.py

Executes the unmodified ``paule.paule.Paule.plan_resynth`` (paule/paule.py:391) with
``use_speech_classifier=True`` (three objectives) and with ``use_somatosensory_feedback=True``
(objective acoustic_semvec -- the only one whose criterion runs in the reference, :624-645; the
``acoustic`` / ``semvec`` variants use undefined names, :697,:750) on seeded random-init models,
fp64, batch 1, VocalTractLab / tube extraction replaced by inert stubs as in ``make_golden.py``.
The tube embedder is constructed with ``dropout=0``: the reference puts it in training mode inside
the loop (:928), so the shipped ``dropout=0.7`` makes its planning stochastic and unpinnable.

Model weights are regenerated from seeds (``branch_models``); the sha256 of every state_dict is stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the librosa / matplotlib / soundfile stubs, imports the reference)

import torch  # noqa: E402

ref_models, ref_paule, O = MG.ref_models, MG.ref_paule, MG.O


def branch_models(dtype, models=ref_models):
    """Seeded random-init branch models; the classifier weight is scaled up so that its gradient matters."""
    torch.manual_seed(1)
    cp_tube = models.ForwardModel(num_lstm_layers=1, hidden_size=360, output_size=10, input_size=30,
                                  apply_half_sequence=False)
    tube_mel = models.ForwardModel(num_lstm_layers=1, hidden_size=360, output_size=60, input_size=10,
                                   apply_half_sequence=True)
    tube_emb = models.EmbeddingModel(input_size=10, num_lstm_layers=2, hidden_size=720, dropout=0.0,
                                     post_upsampling_size=0)
    torch.manual_seed(2)
    cls = models.LinearClassifier(input_dim=60, output_dim=1)
    with torch.no_grad():
        cls.linear.weight.mul_(30.0)
    return cp_tube.to(dtype), tube_mel.to(dtype), tube_emb.to(dtype), cls.to(dtype)


def run(dtype, cp0, tmel, n_inner, objective, branch):
    pred, emb, inv = MG.ref_models_seeded(dtype)
    cp_tube, tube_mel, tube_emb, cls = branch_models(dtype)
    gen = ref_models.Generator().to(dtype)
    gen2 = ref_models.Generator(output_size=60).to(dtype)
    kw = dict(pred_model=pred, inv_model=inv, embedder=emb, cp_gen_model=gen, mel_gen_model=gen2)
    if branch == "cls":
        kw.update(use_speech_classifier=True, speech_classifier=cls)
    else:
        kw.update(use_somatosensory_feedback=True, cp_tube_model=cp_tube, tube_mel_model=tube_mel, tube_embedder=tube_emb)
    pm = ref_paule.Paule(**kw)
    Tm, T = tmel.shape[0], cp0.shape[0]
    ref_paule.speak = lambda cp: (np.zeros((cp.shape[0] - 1) * 110), 44100)
    ref_paule.librosa_melspec = lambda sig, sr: np.zeros((Tm, 60), dtype=tmel.dtype)
    ref_paule.normalize_mel_librosa = lambda m: m
    ref_paule.speak_and_extract_tube_information = lambda cp: (
        np.zeros((cp.shape[0] - 1) * 110), 44100,
        dict(tube_length_cm=np.zeros((T, 40)), tube_area_cm2=np.zeros((T, 40)), incisor_pos_cm=np.zeros(T),
             tongue_tip_side_elevation=np.zeros(T), velum_opening_cm2=np.zeros(T)))
    ref_paule.get_area_info_within_oral_cavity = lambda length, area: np.zeros((T, 7), dtype=tmel.dtype)
    ref_paule.normalize_tube = lambda t: t
    return pm.plan_resynth(target_acoustic=tmel.copy(), initial_cp=cp0.copy(), initialize_from=None,
                           objective=objective, n_outer=1, n_inner=n_inner, log_ii=1, continue_learning=False,
                           verbose=False, log_semantics=False, log_cps=True)


def main():
    out = {}
    torch.set_num_threads(1)
    T, N = 40, 6
    cp0, tmel = O.synthetic_inputs(1, T, seed=5, dtype=torch.float64)
    out["cp0"], out["tmel"] = cp0.numpy(), tmel.numpy()
    for obj in ("acoustic_semvec", "acoustic", "semvec"):
        res = run(torch.float64, cp0[0].numpy(), tmel[0].numpy(), N, obj, "cls")
        tag = f"cls_{obj}"
        out[f"{tag}_planned_cp"] = np.asarray(res.planned_cp)
        out[f"{tag}_loss"] = np.asarray(res.planned_loss_steps)
        out[f"{tag}_mel"] = np.asarray(res.planned_mel_loss_steps)
        out[f"{tag}_cls"] = np.asarray(res.pred_speech_classifier_loss_steps)
        out[f"{tag}_cp_steps"] = np.stack([np.stack(c) for c in res.cp_steps])[0]
    res = run(torch.float64, cp0[0].numpy(), tmel[0].numpy(), N, "acoustic_semvec", "soma")
    out["soma_planned_cp"] = np.asarray(res.planned_cp)
    out["soma_loss"] = np.asarray(res.planned_loss_steps)
    out["soma_mel"] = np.asarray(res.planned_mel_loss_steps)
    out["soma_sem"] = np.asarray(res.pred_semvec_loss_steps)
    out["soma_tube_mel"] = np.asarray(res.pred_tube_mel_loss_steps)
    out["soma_tube_sem"] = np.asarray(res.pred_tube_semvec_loss_steps)
    out["soma_cp_steps"] = np.stack([np.stack(c) for c in res.cp_steps])[0]
    out["soma_pred_tube"] = np.asarray(res.pred_tube)
    out["soma_pred_tube_mel"] = np.asarray(res.pred_tube_mel)
    out["soma_pred_tube_semvec"] = np.asarray(res.pred_tube_semvec)
    models = branch_models(torch.float64)
    out["digest64"] = np.array([O.state_dict_digest(m) for m in models])
    out["cls_w"] = models[3].linear.weight.detach().numpy()
    out["cls_b"] = models[3].linear.bias.detach().numpy()
    out["torch_version"] = np.array(torch.__version__)
    path = os.path.join(HERE, "branches_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in ("cls_acoustic_semvec_loss", "cls_acoustic_semvec_cls", "soma_loss", "soma_tube_mel", "soma_tube_sem"):
        print(k, out[k])


if __name__ == "__main__":
    main()
