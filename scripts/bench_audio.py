"""Audio front-end throughput: 192 windows of 2.1 s (one bench step of cfg 2) -> log-power spectrograms.
GPU: tbn_stft_frames + tensor-core GEMM + tbn_stft_logpower (CUDA events); CPU: the numpy oracle on one core."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from attention_based_tbn_b200 import audio  # noqa: E402
from oracle.stft_oracle import log_power_spectrogram, stft  # noqa: E402

n, L = 192, 50400
rng = np.random.default_rng(0)
ys = (0.1 * rng.standard_normal((n, L))).astype(np.float32)
x = torch.from_numpy(ys).cuda()
fe = audio.SpectrogramFrontEnd("cuda")
for _ in range(3):
    out = fe(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
iters = 20
for _ in range(iters):
    out = fe(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
t0 = time.perf_counter()
for i in range(8):
    ref = log_power_spectrogram(ys[i])
cpu_ms = (time.perf_counter() - t0) / 8 * 1e3
p_ref = np.abs(stft(ys[7])) ** 2
p_got = np.exp(out[7, 0].double().cpu().numpy()) - 1e-6
in_b, out_b = n * L * 4, out.numel() * 4
print(f"GPU: {ms * 1e3:.0f} us per {n} spectrograms ({n / ms * 1e3:.0f} spectrograms/s; in+out {(in_b + out_b) / ms / 1e6:.0f} GB/s algorithmic); "
      f"CPU oracle (numpy rfft, 1 core): {cpu_ms:.2f} ms per spectrogram ({1e3 / cpu_ms:.0f}/s); "
      f"max |dP| / max P = {np.abs(p_got - p_ref).max() / p_ref.max():.2e}")
