#!/usr/bin/env python3
"""Golden d(loss)/d(cp) of the REAL reference, fp64, for pinning the model-path gradient.

Run in the build container only (needs /root/reference):

    python tests/golden/make_grad_golden.py

Executes the unmodified ``paule.paule.Paule.plan_resynth(log_gradients=True, log_cps=True)`` (paule/paule.py:391; the
gradient is ``xx_new.grad`` after ``discrepancy.backward()``, :1052,:1062-1063) in float64 for the three objectives, on an
iid-uniform and on a smooth cp initialisation (B = 1, T = 40, 4 inner steps).  On iid cps the total gradient is dominated
(1e5 : 1) by the local-linear term, so the *model-path* part -- d(mel + semvec terms)/d(cp), what the LSTM BPTT kernels
produce -- is stored separately: total (fp64) minus the analytic smoothness gradient ``oracle.manual_smooth_grad`` (fp64;
``tests/test_oracle.py`` pins that function against autograd to 1e-12).  Weights regenerate from ``torch.manual_seed(0)``
as in make_golden.py (digests are checked by the tests).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the librosa / matplotlib / soundfile stubs, imports the reference)

import torch  # noqa: E402
from oracle import paule_oracle as O  # noqa: E402


def run(cp0, tmel, n_inner, objective):
    pred, emb, inv = MG.ref_models_seeded(torch.float64)
    gen = MG.ref_models.Generator().double()
    gen2 = MG.ref_models.Generator(output_size=60).double()
    pm = MG.ref_paule.Paule(pred_model=pred, inv_model=inv, embedder=emb, cp_gen_model=gen, mel_gen_model=gen2)
    Tm = tmel.shape[0]
    MG.ref_paule.speak = lambda cp: (np.zeros((cp.shape[0] - 1) * 110), 44100)
    MG.ref_paule.librosa_melspec = lambda sig, sr: np.zeros((Tm, 60), dtype=tmel.dtype)
    MG.ref_paule.normalize_mel_librosa = lambda m: m
    return pm.plan_resynth(target_acoustic=tmel.copy(), initial_cp=cp0.copy(), initialize_from=None, objective=objective,
                           n_outer=1, n_inner=n_inner, log_ii=1, continue_learning=False, verbose=False,
                           log_semantics=False, log_cps=True, log_gradients=True)


def main():
    torch.set_num_threads(1)
    T, N = 40, 4
    out = {}
    iid = O.synthetic_inputs(1, T, seed=5, dtype=torch.float64)
    smooth = O.synthetic_inputs(1, T, seed=7, dtype=torch.float64, smooth=True)
    for init, (cp0, tmel) in (("iid", iid), ("smooth", smooth)):
        out[f"{init}_cp0"], out[f"{init}_tmel"] = cp0.numpy(), tmel.numpy()
        for obj in ("acoustic_semvec", "acoustic", "semvec"):
            res = run(cp0[0].numpy(), tmel[0].numpy(), N, obj)
            cps = np.stack([np.stack(c) for c in res.cp_steps])[0]                       # [N,T,30], BEFORE each update
            grads = np.stack([g.double().numpy()[0] for g in res.grad_steps])            # [N,T,30] total xx_new.grad
            smooth_g = np.stack([O.manual_smooth_grad(torch.from_numpy(c)[None]).numpy()[0] for c in cps])
            tag = f"{init}_{obj}"
            out[f"{tag}_cps"], out[f"{tag}_grad"], out[f"{tag}_grad_model"] = cps, grads, grads - smooth_g
            out[f"{tag}_loss"] = np.asarray(res.planned_loss_steps)
            print(tag, "max|total| %.3e  max|model part| %.3e" % (np.abs(grads).max(), np.abs(grads - smooth_g).max()))
    p64, e64, i64 = MG.ref_models_seeded(torch.float64)
    d = MG.digests(p64, e64, i64)
    out["digest64"] = np.array([d["pred"], d["emb"], d["inv"]])
    path = os.path.join(HERE, "grad_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
