"""CPU restatement of the reference's spectrogram -- TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's baseline
legs may import this package).

PARITY UNPINNED at the third-party boundary: the arithmetic lives in ``librosa.stft`` (librosa==0.7.2,
install/requirements.txt:6), which is not installed here and not vendored; the reference has no spectrogram fixtures.
This file restates the published algorithm of librosa 0.7.2 ``core/spectrum.py::stft`` as called from
core/dataset/dataset.py:486-495:

    fft_window = scipy.signal.get_window("hann", win_length, fftbins=True)    # periodic Hann
    fft_window = util.pad_center(fft_window, n_fft)                           # lpad = (n_fft - win_length) // 2
    y = np.pad(y, n_fft // 2, mode="constant")                                # center=True, pad_mode="constant"
    y_frames = util.frame(y, n_fft, hop_length)                               # 1 + (len(y) - n_fft) // hop frames
    stft = rfft(fft_window[:, None] * y_frames, axis=0)                       # (1 + n_fft // 2, frames)

followed by dataset.py:495 ``np.log(np.real(spec * np.conj(spec)) + eps)``.  Independent check available here:
``scipy.signal.stft`` with the same window / hop and zero boundary sees the same samples under the same window, so its
magnitudes (re-scaled by the window sum) must agree frame by frame (tests/test_audio_frontend.py).
"""
from __future__ import annotations

import numpy as np


def hann_periodic(n: int) -> np.ndarray:
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft(y: np.ndarray, n_fft: int = 511, hop: int = 120, win: int = 240) -> np.ndarray:
    """Complex STFT (1 + n_fft // 2, frames), float64."""
    w = np.zeros(n_fft)
    lpad = (n_fft - win) // 2
    w[lpad:lpad + win] = hann_periodic(win)
    yp = np.pad(np.asarray(y, np.float64), n_fft // 2, mode="constant")
    frames = 1 + (yp.shape[0] - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(frames)[None, :]
    return np.fft.rfft(w[:, None] * yp[idx], axis=0)


def log_power_spectrogram(sample: np.ndarray, sampling_rate: int = 24000, window_size: float = 10, step_size: float = 5,
                          eps: float = 1e-6) -> np.ndarray:
    """dataset.py:459-495 (_get_spectrogram, spec_type "stft")."""
    nperseg = int(round(window_size * sampling_rate / 1e3))
    noverlap = int(round(step_size * sampling_rate / 1e3))
    spec = stft(sample, n_fft=511, hop=noverlap, win=nperseg)
    return np.log(np.real(spec * np.conj(spec)) + eps)
