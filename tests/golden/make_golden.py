#!/usr/bin/env python3
"""Generate the golden vectors in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It executes, unmodified, the reference's ``paule.paule.Paule.plan_resynth`` (paule/paule.py:391) and
``paule.models`` (ForwardModel :326, EmbeddingModel :413, InverseModelMelTimeSmoothResidual :177)
on seeded random-init weights and synthetic inputs and stores inputs + outputs as small .npz files.
librosa / soundfile / matplotlib are absent in this image and VocalTractLab synthesis is outside
the hot path, so they are replaced by inert stubs exactly as SURVEY.md appendix A.1 describes
("VTL resynthesis disabled", BASELINE.json configs[0]).

Weights are NOT stored (35 MB); they are regenerated from ``torch.manual_seed(0)`` by constructing
pred -> embedder -> inverse in this order, and the sha256 of every state_dict is stored so that
a regeneration mismatch is detected instead of silently comparing different models.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


mpl = _stub("matplotlib")
mpl.pyplot = _stub("matplotlib.pyplot")
mpl.cm = _stub("matplotlib.cm")
lib = _stub("librosa", resample=lambda wav, **k: wav,
            amplitude_to_db=lambda S, ref=1.0: 20 * np.log10(np.maximum(1e-5, S) / ref))
lib.feature = _stub("librosa.feature", melspectrogram=lambda **k: np.full(
    (k["n_mels"], 1 + len(k["y"]) // k["hop_length"]), 1e-5))
lib.display = _stub("librosa.display")
_stub("soundfile", read=None, write=None)

import torch  # noqa: E402
from paule import models as ref_models  # noqa: E402
from paule import paule as ref_paule  # noqa: E402

from oracle import paule_oracle as O  # noqa: E402

H = 720


def ref_models_seeded(dtype):
    torch.manual_seed(0)
    pred = ref_models.ForwardModel(num_lstm_layers=1, hidden_size=H)
    emb = ref_models.EmbeddingModel(num_lstm_layers=2, hidden_size=H)
    inv = ref_models.InverseModelMelTimeSmoothResidual(num_lstm_layers=1, hidden_size=H)
    return pred.to(dtype), emb.to(dtype), inv.to(dtype)


def digests(pred, emb, inv):
    return {"pred": O.state_dict_digest(pred), "emb": O.state_dict_digest(emb), "inv": O.state_dict_digest(inv)}


def run_real_plan_resynth(dtype, cp0, tmel, n_inner, objective="acoustic_semvec", smiling=False):
    """The REAL reference planner, batch 1 (it is batch-1 only)."""
    pred, emb, inv = ref_models_seeded(dtype)
    gen = ref_models.Generator().to(dtype)
    gen2 = ref_models.Generator(output_size=60).to(dtype)
    pm = ref_paule.Paule(pred_model=pred, inv_model=inv, embedder=emb, cp_gen_model=gen,
                         mel_gen_model=gen2, smiling=smiling)
    Tm = tmel.shape[0]
    ref_paule.speak = lambda cp: (np.zeros((cp.shape[0] - 1) * 110), 44100)
    ref_paule.librosa_melspec = lambda sig, sr: np.zeros((Tm, 60), dtype=tmel.dtype)
    ref_paule.normalize_mel_librosa = lambda m: m
    res = pm.plan_resynth(target_acoustic=tmel.copy(), initial_cp=cp0.copy(), initialize_from=None,
                          objective=objective, n_outer=1, n_inner=n_inner, log_ii=1,
                          continue_learning=False, verbose=False, log_semantics=False, log_cps=True)
    return res


def main():
    out = {}
    torch.set_num_threads(1)                       # deterministic reduction order for the fp32 vectors
    p32, e32, i32 = ref_models_seeded(torch.float32)
    p64, e64, i64 = ref_models_seeded(torch.float64)
    d32, d64 = digests(p32, e32, i32), digests(p64, e64, i64)

    # --- G1: the real plan_resynth, fp64 (the reference's shipped dtype), B=1, T=40, 6 steps, iid init
    T, N = 40, 6
    cp0, tmel = O.synthetic_inputs(1, T, seed=5, dtype=torch.float64)
    for tag, obj, smile in (("real64", "acoustic_semvec", False), ("real64_ac", "acoustic", False),
                            ("real64_sv", "semvec", True)):
        res = run_real_plan_resynth(torch.float64, cp0[0].numpy(), tmel[0].numpy(), N, obj, smile)
        out[f"{tag}_cp0"] = cp0.numpy()
        out[f"{tag}_tmel"] = tmel.numpy()
        out[f"{tag}_planned_cp"] = np.asarray(res.planned_cp)
        out[f"{tag}_loss"] = np.asarray(res.planned_loss_steps)
        out[f"{tag}_mel"] = np.asarray(res.planned_mel_loss_steps)
        out[f"{tag}_vel"] = np.asarray(res.vel_loss_steps)
        out[f"{tag}_jerk"] = np.asarray(res.jerk_loss_steps)
        out[f"{tag}_sem"] = np.asarray(res.pred_semvec_loss_steps)
        out[f"{tag}_cp_steps"] = np.stack([np.stack(c) for c in res.cp_steps])[0]
        out[f"{tag}_pred_mel"] = np.asarray(res.pred_mel)
        out[f"{tag}_pred_semvec"] = np.asarray(res.pred_semvec)

    # --- G2: the real plan_resynth in fp32 (BASELINE.json configs[0] dtype), smooth + iid inits
    cp0s, tmels = O.synthetic_inputs(1, T, seed=7, dtype=torch.float32, smooth=True)
    for tag, c, m in (("real32", cp0.float(), tmel.float()), ("real32_smooth", cp0s, tmels)):
        res = run_real_plan_resynth(torch.float32, c[0].numpy(), m[0].numpy(), N)
        out[f"{tag}_cp0"] = c.numpy()
        out[f"{tag}_tmel"] = m.numpy()
        out[f"{tag}_planned_cp"] = np.asarray(res.planned_cp)
        out[f"{tag}_loss"] = np.asarray(res.planned_loss_steps)
        out[f"{tag}_mel"] = np.asarray(res.planned_mel_loss_steps)
        out[f"{tag}_vel"] = np.asarray(res.vel_loss_steps)
        out[f"{tag}_jerk"] = np.asarray(res.jerk_loss_steps)
        out[f"{tag}_sem"] = np.asarray(res.pred_semvec_loss_steps)

    # --- G3: reference modules + autograd + torch.optim.Adam, batched with per-word losses (B=3, fp32),
    #         including the gradient at every step (what the CUDA backward must reproduce)
    B = 3
    cpb, tmb = O.synthetic_inputs(B, T, seed=11, dtype=torch.float32)
    r = O.plan_inner_loop(p32, e32, cpb, tmb, 5, log_grads=True, log_cps=True)
    out["b3_cp0"], out["b3_tmel"] = cpb.numpy(), tmb.numpy()
    out["b3_loss"], out["b3_terms"] = r["loss"].numpy(), r["terms"].numpy()
    out["b3_grads"] = torch.stack(r["grads"]).numpy()
    out["b3_cps"] = torch.stack(r["cps"]).numpy()
    out["b3_planned_cp"] = r["planned_cp"].numpy()
    out["b3_tsv"] = r["target_semvec"].numpy()
    out["b3_pred_mel"] = r["pred_mel"].numpy()
    out["b3_pred_semvec"] = r["pred_semvec"].numpy()
    # batched == solo (the property the batched driver relies on)
    solo = [O.plan_inner_loop(p32, e32, cpb[i:i + 1], tmb[i:i + 1], 5)["planned_cp"] for i in range(B)]
    out["b3_solo_planned_cp"] = torch.cat(solo).numpy()

    # --- G4: model forwards of the reference modules (fp32 pred/emb, fp64 inverse: it is fp64-only)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 17, 30, generator=g) - 0.5                 # odd T: AvgPool drops the last frame
    mel = torch.rand(2, 12, 60, generator=g)
    with torch.no_grad():
        out["fw_x"], out["fw_y"] = x.numpy(), p32(x).numpy()
        lens = (torch.tensor(12), torch.tensor(7))
        out["em_x"], out["em_lens"], out["em_y"] = mel.numpy(), np.array([12, 7]), e32(mel, lens).numpy()
        out["inv_x"] = mel.double().numpy()
        out["inv_y"] = i64(mel.double()).numpy()

    out["digest32"] = np.array([d32["pred"], d32["emb"], d32["inv"]])
    out["digest64"] = np.array([d64["pred"], d64["emb"], d64["inv"]])
    out["torch_version"] = np.array(torch.__version__)
    path = os.path.join(HERE, "paule_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
