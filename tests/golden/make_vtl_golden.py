#!/usr/bin/env python3
"""Golden output of the REFERENCE's VocalTractLab wrappers for a fixed cp trajectory (build container only: needs
/root/reference and its pre-built synthesiser binary, which loads here):

    python tests/golden/make_vtl_golden.py

Runs the unmodified ``paule.util.speak`` (paule/util.py:175-249) and ``speak_and_extract_tube_information`` (:317-433) on
``inv_normalize_cp`` of a smooth seeded trajectory and stores the signal (float64, exact), the tube information and the cp
normalisation round trip.  ``tests/test_host.py`` replays the same cps through ``paule_b200.audio.VocalTractLab`` wherever the
synthesiser binary is available and expects bit-identical audio.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402,F401  (installs the librosa / matplotlib / soundfile stubs)

from paule import util as ref_util  # noqa: E402


def trajectory(T=60, seed=3):
    rng = np.random.RandomState(seed)
    t = np.arange(T)[:, None]
    f = rng.rand(1, 30) * 0.02 + 0.005
    ph = rng.rand(1, 30) * 6.283185307179586
    return 0.3 * np.sin(6.283185307179586 * f * t + ph)          # normalised cps, |x| <= 0.3


def main():
    cp_norm = trajectory()
    cp = ref_util.inv_normalize_cp(cp_norm)
    sig, sr = ref_util.speak(cp)
    sig2, sr2, tube = ref_util.speak_and_extract_tube_information(cp)
    out = {"cp_norm": cp_norm, "cp": cp, "sig": np.asarray(sig), "sr": np.array(sr), "sig_framewise": np.asarray(sig2),
           "tube_area_cm2": tube["tube_area_cm2"], "tube_length_cm": tube["tube_length_cm"],
           "incisor_pos_cm": tube["incisor_pos_cm"], "tongue_tip_side_elevation": tube["tongue_tip_side_elevation"],
           "velum_opening_cm2": tube["velum_opening_cm2"],
           "mel_mean_librosa_quoted": np.array(-83.52182518111363)}          # the constant util.py:136 quotes
    path = os.path.join(HERE, "vtl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; signal", np.asarray(sig).shape, "max |sig|", float(np.abs(sig).max()))


if __name__ == "__main__":
    main()
