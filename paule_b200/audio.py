"""Host-side audio pipeline of PAULE: VocalTractLab synthesis and the log-mel front-end.

This is NOT on the GPU hot path -- BASELINE.json's north_star keeps "the VocalTractLab C++ synthesis and the librosa mel
recomputation host-side and off the hot path, called only at outer-loop boundaries".  It is what turns ``paule_b200.Paule``
into a complete drop-in for real (non-synthetic) use:

* ``VocalTractLab``      ctypes binding of the synthesiser's C API (the binary the reference ships as
                         ``paule/vocaltractlab_api/libVocalTractLabApi.so``; it is NOT redistributed here -- pass its path or set
                         ``PAULE_VTL_LIB`` / ``PAULE_VTL_SPEAKER``).  ``speak`` restates paule/util.py:175-249, ``speak_and_tube``
                         :317-433 (tube areas / incisor / tongue tip / velum per frame for the somatosensory branch).
* ``mel_spectrogram``    the reference's ``librosa_melspec`` (paule/util.py:115-120) without librosa: centred STFT (n_fft 1024,
                         hop 220, periodic Hann window), Slaney mel filterbank (60 bands, 10 Hz .. 12 kHz, area-normalised),
                         ``amplitude_to_db(ref=0.15, amin=1e-5, top_db=80)``; ``normalize_mel`` = util.py:137-146.
                         librosa is a third-party dependency that is absent from this image, so the restatement follows its
                         published algorithm (librosa 0.10 defaults) and is pinned only by the constant the reference quotes
                         (-83.52182518111363 dB for silence, util.py:136) and by its own unit tests: PARITY UNPINNED against
                         librosa itself.  Resampling (only needed for targets that are not 44.1 kHz; VocalTractLab emits
                         44.1 kHz) uses a polyphase filter, not librosa's ``kaiser_best``.
* ``make_synthesizer``   the callable ``Paule(synthesizer=...)`` expects: normalised cps -> (signal, rate, normalised log-mel),
                         i.e. ``speak(inv_normalize_cp(cp))`` -> ``librosa_melspec`` -> ``normalize_mel_librosa``
                         (paule/paule.py:1097-1104).
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Tuple

import numpy as np

SAMPLE_RATE = 44100
FRAME_STEPS = 110            # audio samples per cp frame (2.5 ms), paule/util.py:213
N_FFT, HOP, N_MELS, FMIN, FMAX = 1024, 220, 60, 10.0, 12000.0   # paule/util.py:118
DB_REF, DB_AMIN, TOP_DB = 0.15, 1e-5, 80.0                      # paule/util.py:119 + librosa defaults

# cp normalisation of the reference (paule/util.py:70-86): the models work on (cp - mean) / std, VocalTractLab on raw cps
CP_THEORETICAL_MEANS = np.array([5.0e-01, -4.75, -2.5e-01, -3.5, 0.0, 1.0, 5.0e-01, 4.5e-01, 5.0e-01, -1.0, 3.5, -2.5e-01, 5.0e-01,
                                 1.0, -1.0, -3.0, 5.0e-01, 5.0e-01, 0.0, 3.2e+02, 1.0e+04, 1.25e-01, 1.25e-01, 0.0, 1.57075, 0.0,
                                 5.0e-01, 0.0, 5.0e+01, -2.0e+01])
CP_THEORETICAL_STDS = np.array([5.0e-01, 1.25, 2.5e-01, 3.5, 1.0, 3.0, 5.0e-01, 5.5e-01, 3.5, 2.0, 2.0, 2.75, 3.5, 4.0, 3.0, 3.0,
                                5.0e-01, 5.0e-01, 1.0, 2.8e+02, 1.0e+04, 1.75e-01, 1.75e-01, 2.5e-01, 1.57075, 1.0, 5.0e-01,
                                5.0e-01, 5.0e+01, 2.0e+01])
# tube normalisation (paule/util.py:88-112): 7 section areas, incisor position, tongue tip, velum opening
_TUBE_MINS = np.array([0.0] * 7 + [14.0, -1.0, 0.0])
_TUBE_MAXS = np.array([15.0] * 7 + [18.0, 1.0, 1.0])
TUBE_THEORETICAL_MEANS = (_TUBE_MINS + _TUBE_MAXS) / 2.0
TUBE_THEORETICAL_STDS = np.std(np.stack([_TUBE_MINS, _TUBE_MAXS]), axis=0)


def normalize_cp(cp):
    return (cp - CP_THEORETICAL_MEANS) / CP_THEORETICAL_STDS


def inv_normalize_cp(norm_cp):
    return CP_THEORETICAL_STDS * norm_cp + CP_THEORETICAL_MEANS


# ---------------------------------------------------------------------------------------------------------------------------
# log-mel front-end
# ---------------------------------------------------------------------------------------------------------------------------
def _hz_to_mel(f):
    """Slaney's auditory-toolbox mel scale (librosa's default, htk=False): linear below 1 kHz, logarithmic above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, min_log_hz) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int = SAMPLE_RATE, n_fft: int = N_FFT, n_mels: int = N_MELS, fmin: float = FMIN, fmax: float = FMAX):
    """Triangular mel filters [n_mels, 1 + n_fft // 2], Slaney area normalisation (librosa.filters.mel defaults)."""
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    return weights * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]


_FILTERS = {}


def resample(wav: np.ndarray, orig_sr: int, target_sr: int = SAMPLE_RATE) -> np.ndarray:
    """Polyphase resampling to the front-end's rate (the reference uses librosa/resampy ``kaiser_best``: same pass band, a
    different anti-aliasing filter -- targets recorded at 44.1 kHz and everything VocalTractLab produces bypass this)."""
    if int(orig_sr) == int(target_sr):
        return np.asarray(wav, dtype=np.float64)
    from math import gcd
    from scipy.signal import resample_poly
    g = gcd(int(orig_sr), int(target_sr))
    return resample_poly(np.asarray(wav, dtype=np.float64), int(target_sr) // g, int(orig_sr) // g)


def mel_spectrogram(wav: np.ndarray, sample_rate: int, pad_mode: str = "constant") -> np.ndarray:
    """paule/util.py:115-120 ``librosa_melspec``: waveform -> log-mel [frames, 60] in dB (float64), frames = 1 + len(wav) // 220."""
    y = resample(np.asarray(wav, dtype=np.float64), sample_rate)
    key = (SAMPLE_RATE, N_FFT, N_MELS, FMIN, FMAX)
    if key not in _FILTERS:
        n = np.arange(N_FFT)
        _FILTERS[key] = (mel_filterbank(), 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT))      # periodic Hann (fftbins=True)
    fb, window = _FILTERS[key]
    y = np.pad(y, N_FFT // 2, mode=pad_mode)                                                   # center=True
    n_frames = 1 + (len(y) - N_FFT) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    spec = np.abs(np.fft.rfft(y[idx] * window[None, :], axis=1))                              # [frames, 513], power = 1
    mel = spec @ fb.T                                                                          # [frames, 60]
    # amplitude_to_db(S, ref, amin, top_db) = power_to_db(S**2, ref**2, amin**2, top_db)
    log_spec = 10.0 * np.log10(np.maximum(DB_AMIN ** 2, mel ** 2)) - 10.0 * np.log10(max(DB_AMIN ** 2, DB_REF ** 2))
    log_spec = np.maximum(log_spec, log_spec.max() - TOP_DB)
    return np.array(log_spec, order="C", dtype=np.float64)


MEL_MEAN = float(mel_spectrogram(np.zeros(5000), SAMPLE_RATE)[0, 0])     # -83.52182518111363 (paule/util.py:136-137)
MEL_STD = abs(MEL_MEAN)


def normalize_mel(mel):
    """paule/util.py:141-142 ``normalize_mel_librosa``."""
    return (mel - MEL_MEAN) / MEL_STD


def inv_normalize_mel(norm_mel):
    return MEL_STD * norm_mel + MEL_MEAN


def target_mel_from_audio(sig: np.ndarray, sr: int) -> np.ndarray:
    """The acoustic target of ``plan_resynth`` from a waveform (paule/paule.py:524-526): normalised log-mel, shifted to min 0."""
    if np.ndim(sig) == 2:                    # stereo -> mono
        sig = np.mean(sig, axis=1)
    mel = normalize_mel(mel_spectrogram(sig, sr))
    return mel - mel.min()


def read_audio(path: str) -> Tuple[np.ndarray, int]:
    """Waveform of an audio file: ``soundfile`` when it is installed (the reference's reader, paule/paule.py:487: flac / wav /
    ogg), else RIFF wav through scipy."""
    try:
        import soundfile as sf
        return sf.read(path)
    except ImportError:
        from scipy.io import wavfile
        if not path.lower().endswith(".wav"):
            raise ImportError("reading '%s' needs the soundfile package (only RIFF .wav can be read without it)" % path)
        sr, data = wavfile.read(path)
        if np.issubdtype(data.dtype, np.integer):
            data = data.astype(np.float64) / float(np.iinfo(data.dtype).max)
        return data.astype(np.float64), int(sr)


# ---------------------------------------------------------------------------------------------------------------------------
# VocalTractLab
# ---------------------------------------------------------------------------------------------------------------------------
class VocalTractLab:
    """ctypes binding of the VocalTractLab API 2.x the reference calls (paule/util.py:29-41: load + vtlInitialize; convention:
    ``int`` return, 0 = OK, anything else raises ``ValueError`` as in the reference).  The library keeps global synthesis
    state, so calls are serialised by a lock (a thread pool of synthesis jobs needs one process per worker to run in parallel)."""

    def __init__(self, library_path: Optional[str] = None, speaker_file: Optional[str] = None):
        library_path = library_path or os.environ.get("PAULE_VTL_LIB")
        if not library_path:
            here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vocaltractlab_api")
            library_path = os.path.join(here, "libVocalTractLabApi.so")
        speaker_file = speaker_file or os.environ.get("PAULE_VTL_SPEAKER") or os.path.join(os.path.dirname(library_path), "JD3.speaker")
        if not os.path.exists(library_path) or not os.path.exists(speaker_file):
            raise FileNotFoundError(f"VocalTractLab library '{library_path}' / speaker '{speaker_file}' not found: the synthesiser "
                                    "binary ships with the reference package (paule/vocaltractlab_api/), not with paule_b200; "
                                    "pass library_path= / speaker_file= or set PAULE_VTL_LIB / PAULE_VTL_SPEAKER")
        self._lib = ctypes.cdll.LoadLibrary(library_path)
        failure = self._lib.vtlInitialize(ctypes.c_char_p(speaker_file.encode()))
        if failure != 0:
            raise ValueError('Error in vtlInitialize! Errorcode: %i' % failure)
        version = ctypes.c_char_p(b' ' * 64)
        self._lib.vtlGetVersion(version)
        self.version = version.value.decode()
        self._lock = threading.Lock()
        c = [ctypes.c_int(0) for _ in range(5)] + [ctypes.c_double(0)]
        self._lib.vtlGetConstants(*[ctypes.byref(v) for v in c])
        self.audio_sampling_rate, self.n_tube_sections, self.n_tract, self.n_glottis = (v.value for v in c[:4])
        if (self.audio_sampling_rate, self.n_tract, self.n_glottis) != (SAMPLE_RATE, 19, 11):
            raise ValueError(f"unexpected VocalTractLab constants: rate {self.audio_sampling_rate}, {self.n_tract} tract / "
                             f"{self.n_glottis} glottis parameters (paule/util.py:209-211 expects 44100 / 19 / 11)")

    def speak(self, cp_param: np.ndarray) -> Tuple[np.ndarray, int]:
        """Raw (NOT normalised) cps [frames, 30] -> (signal [(frames - 1) * 110], 44100); paule/util.py:175-249."""
        cp = np.ascontiguousarray(cp_param, dtype=np.float64)
        n = cp.shape[0]
        audio = (ctypes.c_double * int((n - 1) * FRAME_STEPS + 2000))()       # 2000 samples of head room, as the reference
        tract = (ctypes.c_double * (n * 19))(*np.ascontiguousarray(cp[:, 0:19]).reshape(-1))
        glottis = (ctypes.c_double * (n * 11))(*np.ascontiguousarray(cp[:, 19:30]).reshape(-1))
        with self._lock:
            failure = self._lib.vtlSynthesisReset()
            if failure != 0:
                raise ValueError(f'Error in vtlSynthesisReset! Errorcode: {failure}')
            failure = self._lib.vtlSynthBlock(ctypes.byref(tract), ctypes.byref(glottis), n, FRAME_STEPS, ctypes.byref(audio), 0)
            if failure != 0:
                raise ValueError('Error in vtlSynthBlock! Errorcode: %i' % failure)
        return np.array(audio[:-2000]), SAMPLE_RATE

    def speak_and_tube(self, cp_param: np.ndarray):
        """(signal, 44100, tube) with tube [frames, 10] = 7 pooled section areas, incisor position, tongue-tip side elevation,
        velum opening, UN-normalised -- frame-by-frame synthesis with tube export (paule/util.py:317-433).  The reference
        returns the 40 raw sections; what its somatosensory models consume is the 10-channel reduction below."""
        cp = np.ascontiguousarray(cp_param, dtype=np.float64)
        n = cp.shape[0]
        audio = np.zeros((max(n - 1, 1), FRAME_STEPS))
        length = (ctypes.c_double * 40)(); area = (ctypes.c_double * 40)(); arti = (ctypes.c_int * 40)()
        incisor, tip, velum = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
        areas = np.zeros((n, 40)); lengths = np.zeros((n, 40)); extra = np.zeros((n, 3))
        with self._lock:
            failure = self._lib.vtlSynthesisReset()
            if failure != 0:
                raise ValueError(f'Error in vtlSynthesisReset! Errorcode: {failure}')
            for i in range(n):
                tract = (ctypes.c_double * 19)(*cp[i, 0:19])
                glottis = (ctypes.c_double * 11)(*cp[i, 19:30])
                buf = (ctypes.c_double * FRAME_STEPS)()
                failure = self._lib.vtlSynthesisAddTract(0 if i == 0 else FRAME_STEPS, ctypes.byref(buf), ctypes.byref(tract),
                                                         ctypes.byref(glottis))
                if failure != 0:
                    raise ValueError('Error in vtlSynthesisAddTract! Errorcode: %i' % failure)
                if i > 0:
                    audio[i - 1] = np.array(buf)
                failure = self._lib.vtlTractToTube(ctypes.byref(tract), ctypes.byref(length), ctypes.byref(area), ctypes.byref(arti),
                                                   ctypes.byref(incisor), ctypes.byref(tip), ctypes.byref(velum))
                if failure != 0:
                    raise ValueError('Error in vtlTractToTube! Errorcode: %i' % failure)
                areas[i], lengths[i] = np.array(area), np.array(length)
                extra[i] = (incisor.value, tip.value, velum.value)
        info = {"tube_length_cm": lengths, "tube_area_cm2": areas, "incisor_pos_cm": extra[:, 0],
                "tongue_tip_side_elevation": extra[:, 1], "velum_opening_cm2": extra[:, 2]}
        return audio.reshape(-1), SAMPLE_RATE, info


def make_synthesizer(vtl: VocalTractLab):
    """The callable ``Paule(synthesizer=...)`` expects: NORMALISED cps [T,30] -> (signal, 44100, normalised log-mel [T // 2, 60])
    -- ``speak(inv_normalize_cp(cp))`` -> ``librosa_melspec`` -> ``normalize_mel_librosa`` (paule/paule.py:1097-1104)."""
    def synthesize(cp_norm: np.ndarray):
        sig, sr = vtl.speak(inv_normalize_cp(np.asarray(cp_norm, dtype=np.float64)))
        mel = normalize_mel(mel_spectrogram(sig, sr))
        return sig, sr, mel[: cp_norm.shape[0] // 2].astype(np.float32)
    return synthesize
