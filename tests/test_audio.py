"""Host-side audio pipeline (paule_b200/audio.py).  CPU only.

* log-mel front-end: librosa is absent from this image, so the restatement of ``librosa_melspec`` (paule/util.py:115-120) is
  held to the constant the reference quotes for silence (util.py:136), to the published properties of the algorithm (frame
  count, Slaney filterbank normalisation, dB reference / floor / 80 dB dynamic range) and to an analytic sine;
* VocalTractLab binding: bit-identical audio and tube areas against the REFERENCE's own wrappers
  (tests/golden/make_vtl_golden.py), wherever the synthesiser binary is present (it ships with the reference, not here).
"""
import os

import numpy as np
import pytest

from paule_b200 import audio as A

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VTL_DIR = os.environ.get("PAULE_VTL_DIR", "/root/reference/paule/vocaltractlab_api")


def test_silence_gives_the_constant_the_reference_quotes():
    assert abs(A.MEL_MEAN - (-83.52182518111363)) < 1e-9            # paule/util.py:136
    m = A.mel_spectrogram(np.zeros(5000), 44100)
    assert m.shape == (1 + 5000 // 220, 60) and np.all(m == m[0, 0])
    np.testing.assert_allclose(A.normalize_mel(m), 0.0 * m, atol=1e-12)           # silence is the zero of the normalised scale
    np.testing.assert_allclose(A.inv_normalize_mel(A.normalize_mel(m)), m, atol=1e-9)


def test_mel_filterbank_is_slaney_normalised():
    fb = A.mel_filterbank()
    assert fb.shape == (60, 513) and np.all(fb >= 0)
    freqs = np.linspace(0, 22050, 513)
    centre = freqs[fb.argmax(1)]
    assert np.all(np.diff(centre) > 0) and centre[0] < 200 and 10000 < centre[-1] < 12000
    assert np.all(fb[:, freqs > 12000 + 1e-9] == 0)
    # Slaney normalisation: every triangle has unit area over frequency (up to the FFT-bin discretisation)
    area = fb.sum(1) * (freqs[1] - freqs[0])
    np.testing.assert_allclose(area[10:], 1.0, atol=0.12)
    # the mel scale is linear below 1 kHz and logarithmic above
    np.testing.assert_allclose(A._mel_to_hz(A._hz_to_mel([10.0, 500.0, 1000.0, 4000.0, 12000.0])), [10.0, 500.0, 1000.0, 4000.0, 12000.0])
    np.testing.assert_allclose(A._hz_to_mel(1000.0), 15.0)


def test_sine_lands_in_the_right_band_at_the_right_level():
    sr, f0, amp = 44100, 1000.0, 0.5
    t = np.arange(sr // 2) / sr
    m = A.mel_spectrogram(amp * np.sin(2 * np.pi * f0 * t), sr)
    assert m.shape == (1 + len(t) // 220, 60)
    fb = A.mel_filterbank()
    freqs = np.linspace(0, 22050, 513)
    band = int(np.argmin(np.abs(freqs[fb.argmax(1)] - f0)))
    mid = m[20:-20]
    assert np.all(np.abs(mid.argmax(1) - band) <= 1)
    # |STFT| of a sine at a bin centre = amp * sum(window) / 2 = amp * N / 4; through the filter's peak weight; in dB re 0.15
    k = int(round(f0 * 1024 / sr))
    peak = amp * 1024 / 4 * fb[:, k].max()
    want = 20 * np.log10(peak / 0.15)
    # the Hann main lobe spreads the line over the neighbouring bins (half amplitude each), which the filter sums up as well:
    # between the peak bin alone and twice that
    assert want - 0.5 < mid.max(1).mean() < want + 6.5
    # 80 dB dynamic range below the loudest cell
    assert m.min() >= m.max() - 80.0 - 1e-9


def test_target_from_audio_and_resampling():
    sr = 22050
    t = np.arange(sr // 4) / sr
    sig = 0.3 * np.sin(2 * np.pi * 440.0 * t)
    mel = A.target_mel_from_audio(np.stack([sig, sig], 1), sr)       # stereo, 22.05 kHz -> mono, 44.1 kHz
    assert mel.shape == (1 + (2 * len(t)) // 220, 60) and mel.min() == 0.0 and np.isfinite(mel).all()
    np.testing.assert_allclose(A.inv_normalize_cp(A.normalize_cp(np.ones(30))), np.ones(30), atol=1e-12)


@pytest.mark.skipif(not os.path.exists(os.path.join(VTL_DIR, "libVocalTractLabApi.so")),
                    reason="the VocalTractLab binary ships with the reference package, not with paule_b200")
def test_vocaltractlab_binding_matches_the_reference_wrappers():
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "vtl_golden.npz")))
    vtl = A.VocalTractLab(os.path.join(VTL_DIR, "libVocalTractLabApi.so"), os.path.join(VTL_DIR, "JD3.speaker"))
    assert "API 2" in vtl.version
    np.testing.assert_array_equal(A.inv_normalize_cp(g["cp_norm"]), g["cp"])
    sig, sr = vtl.speak(g["cp"])
    assert sr == int(g["sr"]) == 44100 and sig.shape == ((g["cp"].shape[0] - 1) * 110,)
    np.testing.assert_array_equal(sig, g["sig"])                                      # bit-identical audio
    sig2, sr2, tube = vtl.speak_and_tube(g["cp"])
    np.testing.assert_array_equal(sig2, g["sig_framewise"])
    np.testing.assert_array_equal(tube["tube_area_cm2"], g["tube_area_cm2"])
    np.testing.assert_array_equal(tube["incisor_pos_cm"], g["incisor_pos_cm"])
    np.testing.assert_array_equal(tube["velum_opening_cm2"], g["velum_opening_cm2"])
    # the synthesizer Paule(synthesizer=...) expects: normalised cps -> (signal, rate, normalised log-mel [T // 2, 60])
    s, r, mel = A.make_synthesizer(vtl)(g["cp_norm"])
    np.testing.assert_array_equal(s, g["sig"])
    assert mel.shape == (g["cp"].shape[0] // 2, 60) and mel.dtype == np.float32 and np.isfinite(mel).all()
    with pytest.raises(FileNotFoundError):
        A.VocalTractLab("/nonexistent/libVocalTractLabApi.so")
