"""GPU audio front-end: the log-power STFT that produces the Audio input of ``TBNModel`` (SURVEY.md section 8f,
rank 3).  Mirrors ``Video_Dataset._get_audio_segment`` / ``_get_spectrogram`` (core/dataset/dataset.py:421-511,
``spec_type == "stft"``): a window of ``audio_length`` seconds around the frame, ``librosa.stft(n_fft=511,
hop_length=round(5 ms), win_length=round(10 ms), window="hann", pad_mode="constant")`` (centred), then
``log(|X|^2 + 1e-6)``.

The reference computes one spectrogram per segment on a data-loader worker with librosa; at B200 speed the model
consumes ~45,000 of them per second.  Here the transform is a tensor-core GEMM against a windowed DFT matrix
(csrc/stft.cu); host code only slices the windows.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .sampling import audio_window_start


def extract_segment(aud_sample: np.ndarray, frame_idx: int, vid_fps: int = 60, audio_length: float = 2.1,
                    sampling_rate: int = 24000) -> np.ndarray:
    """The 1-D window of dataset.py:439-451: zero-pad a recording shorter than the window, centre on ``frame_idx``,
    clamp to the recording."""
    min_len = int(audio_length * sampling_rate)
    start = audio_window_start(frame_idx, aud_sample.shape[0], vid_fps, audio_length, sampling_rate)
    if aud_sample.shape[0] < min_len:
        aud_sample = np.pad(aud_sample, (0, min_len - aud_sample.shape[0]))
    return aud_sample[start:start + min_len]


def dft_matrix(n_fft: int, win: int) -> torch.Tensor:
    """(2*bins, win) float64: rows [0, bins) = w[j] cos(2 pi k (j + off) / n_fft), rows [bins, 2 bins) = -w[j] sin(...),
    w = periodic Hann(win) (scipy.signal.get_window("hann", win, fftbins=True)), off = (n_fft - win) // 2
    (librosa.util.pad_center)."""
    bins = n_fft // 2 + 1
    j = torch.arange(win, dtype=torch.float64)
    w = 0.5 - 0.5 * torch.cos(2.0 * math.pi * j / win)
    off = (n_fft - win) // 2
    k = torch.arange(bins, dtype=torch.float64).reshape(-1, 1)
    # reduce k * m modulo n_fft in integers before the multiplication by 2 pi / n_fft: full float64 accuracy
    km = (torch.arange(bins).reshape(-1, 1) * (torch.arange(win) + off).reshape(1, -1)) % n_fft
    ang = 2.0 * math.pi * km.double() / n_fft
    del k
    return torch.cat([w * torch.cos(ang), -w * torch.sin(ang)], 0)


class SpectrogramFrontEnd:
    """``fe(samples)``: (n, seg_len) float32 CUDA tensor of audio windows -> (n, 1, n_fft // 2 + 1, frames) float32
    log-power spectrograms, ``frames = 1 + (seg_len - 1 + 2 * (n_fft // 2) - n_fft + 1) // hop`` (librosa.util.frame)."""

    def __init__(self, device, sampling_rate: int = 24000, window_size: float = 10, step_size: float = 5, n_fft: int = 511,
                 eps: float = 1e-6):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TbnKernelError("SpectrogramFrontEnd needs a CUDA device (there is no CPU fallback)")
        _lib.load()
        self.n_fft = n_fft
        self.win = int(round(window_size * sampling_rate / 1e3))   # nperseg, dataset.py:483
        self.hop = int(round(step_size * sampling_rate / 1e3))     # noverlap is used as the hop length, dataset.py:484-491
        self.eps = float(eps)
        self.bins = n_fft // 2 + 1
        self.kp = (self.win + 63) // 64 * 64
        d = dft_matrix(n_fft, self.win)
        dp = torch.zeros(2 * self.bins, self.kp, dtype=torch.float64)
        dp[:, : self.win] = d
        hi = dp.to(torch.bfloat16)
        lo = (dp - hi.double()).to(torch.bfloat16)
        self.dft3 = torch.cat([hi, lo, hi], 1).contiguous().to(self.device)  # pairs with samples split [hi | hi | lo]
        self.zero_bias = torch.zeros(2 * self.bins, dtype=torch.float32, device=self.device)
        self._ws = {}

    def num_frames(self, seg_len: int) -> int:
        return 1 + (seg_len + 2 * (self.n_fft // 2) - self.n_fft) // self.hop

    def _buf(self, name, numel, dtype):
        t = self._ws.get(name)
        if t is None or t.numel() < numel:
            t = self._ws[name] = torch.empty(numel, dtype=dtype, device=self.device)
        return t[:numel]

    def __call__(self, samples: torch.Tensor) -> torch.Tensor:
        if samples.dim() != 2 or samples.dtype != torch.float32 or not samples.is_cuda:
            raise ValueError("samples must be a (n, seg_len) float32 CUDA tensor")
        if samples.stride(1) != 1:
            samples = samples.contiguous()
        n, seg_len = samples.shape
        frames = self.num_frames(seg_len)
        rows = n * frames
        stream = torch.cuda.current_stream(self.device).cuda_stream
        a3 = self._buf("a3", rows * 3 * self.kp, torch.bfloat16)
        g = self._buf("g", rows * 2 * self.bins, torch.float32)
        out = torch.empty((n, 1, self.bins, frames), dtype=torch.float32, device=self.device)
        _lib.call("tbn_stft_frames", samples.data_ptr(), samples.stride(0), n, seg_len, self.n_fft, self.win, self.hop,
                  a3.data_ptr(), stream)
        _lib.call("tbn_conv_tc", a3.data_ptr(), self.dft3.data_ptr(), self.zero_bias.data_ptr(), g.data_ptr(), 1, 1, rows,
                  3 * self.kp, 3 * self.kp, 0, 2 * self.bins, 2 * self.bins, 0, 1, 0, _lib.TBN_F32, 1, stream)
        _lib.call("tbn_stft_logpower", g.data_ptr(), out.data_ptr(), n, frames, self.bins, self.eps, stream)
        return out

    def from_recordings(self, recordings: Sequence[np.ndarray], frame_indices: Sequence[Sequence[int]], vid_fps: int = 60,
                        audio_length: float = 2.1, sampling_rate: int = 24000) -> torch.Tensor:
        """Slice the windows on the host (dataset.py:439-451), one pinned copy, one transform: recordings[i] with the
        segment centre frames frame_indices[i] -> (sum_i len(frame_indices[i]), 1, bins, frames)."""
        segs = [extract_segment(rec, int(f), vid_fps, audio_length, sampling_rate)
                for rec, idx in zip(recordings, frame_indices) for f in idx]
        host = torch.from_numpy(np.stack(segs).astype(np.float32, copy=False)).pin_memory()
        return self(host.to(self.device, non_blocking=True))
