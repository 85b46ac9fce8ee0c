"""GPU audio front-end: log-power STFT of the audio windows.

The reference's arithmetic is librosa.stft (absent here -> "parity unpinned", see oracle/stft_oracle.py); the oracle's
restatement is checked against scipy.signal.stft, an independent implementation that sees the same samples under the
same window, and the CUDA path against the oracle."""
import numpy as np
import pytest
import torch

from attention_based_tbn_b200 import audio, sampling


def _signal(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 24000.0
    y = 0.05 * rng.standard_normal(n)
    for f, a in ((220.0, 0.5), (3150.0, 0.05), (9000.0, 0.005)):
        y += a * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
    return y.astype(np.float32)


def test_oracle_magnitudes_match_scipy_stft():
    import scipy.signal as ss
    from oracle.stft_oracle import hann_periodic, stft

    for n in (30696, 50400, 1000, 511):
        y = _signal(n, n).astype(np.float64)
        X = stft(y)
        assert X.shape == (256, 1 + (n - 1) // 120)  # librosa.util.frame on the centre-padded signal
        _, _, Z = ss.stft(y, window=hann_periodic(240), nperseg=240, noverlap=120, nfft=511, boundary="zeros", padded=False)
        Z = Z * hann_periodic(240).sum()  # scipy divides by the window sum
        assert np.abs(np.abs(X) - np.abs(Z[:, : X.shape[1]])).max() < 1e-12


def test_frame_counts_of_the_baseline_configs():
    # 1.279 s -> the 256 x 256 input of the TBN paper (cfg 1); 2.1 s -> 256 x 420 under librosa's framing
    from oracle.stft_oracle import log_power_spectrogram

    assert log_power_spectrogram(np.zeros(int(1.279 * 24000), np.float32)).shape == (256, 256)
    assert log_power_spectrogram(np.zeros(int(2.1 * 24000), np.float32)).shape == (256, 420)
    assert np.allclose(log_power_spectrogram(np.zeros(1000, np.float32)), np.log(1e-6))


def test_dft_matrix_reproduces_the_complex_stft():
    from oracle.stft_oracle import stft

    y = _signal(5000, 3).astype(np.float64)
    X = stft(y)
    d = audio.dft_matrix(511, 240).numpy()
    for t in (0, 1, 7, X.shape[1] - 1):
        seg = np.zeros(240)
        for j in range(240):
            l = t * 120 - 120 + j
            if 0 <= l < y.shape[0]:
                seg[j] = y[l]
        r = d @ seg
        assert np.abs(r[:256] - X.real[:, t]).max() < 1e-11 and np.abs(r[256:] - X.imag[:, t]).max() < 1e-11


def test_extract_segment_follows_the_reference_window_rule():
    rec = np.arange(24000 * 10, dtype=np.float32)
    seg = audio.extract_segment(rec, 300)  # centre 5.0 s, window 2.1 s
    assert seg.shape == (50400,) and seg[0] == int((300 / 60 - 1.05) * 24000)
    assert audio.extract_segment(rec, 0)[0] == 0  # clamped at the start
    assert audio.extract_segment(rec, 600)[-1] == rec[-1]  # clamped at the end
    assert sampling.audio_window_start(300, rec.shape[0]) == int((300 / 60 - 1.05) * 24000)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seg_len", [(1, 30696), (5, 50400), (3, 1000), (2, 511), (4, 12345)])
def test_spectrogram_frontend_vs_oracle(n, seg_len):
    from oracle.stft_oracle import log_power_spectrogram, stft

    fe = audio.SpectrogramFrontEnd("cuda")
    ys = np.stack([_signal(seg_len, 17 * i + seg_len) for i in range(n)])
    base = torch.zeros(n, seg_len + 24, device="cuda")
    base[:, 8:8 + seg_len] = torch.from_numpy(ys).cuda()
    out = fe(base[:, 8:8 + seg_len])  # strided rows, unaligned start
    assert out.shape == (n, 1, 256, fe.num_frames(seg_len)) and out.dtype == torch.float32
    got = out[:, 0].double().cpu().numpy()
    for i in range(n):
        ref = log_power_spectrogram(ys[i])
        p_ref = np.abs(stft(ys[i])) ** 2
        p_got = np.exp(got[i]) - 1e-6
        # power relative to the loudest bin: the split-bf16 GEMM is accurate to ~2^-16 per product (fp32 class)
        assert np.abs(p_got - p_ref).max() <= 2e-5 * p_ref.max()
        loud = p_ref >= 1e-5 * p_ref.max()
        assert np.abs(got[i][loud] - ref[loud]).max() < 2e-2
        assert np.abs(got[i] - ref)[loud].mean() < 1e-3


@pytest.mark.gpu
def test_frontend_from_recordings_feeds_the_model_input_shape():
    fe = audio.SpectrogramFrontEnd("cuda")
    recs = [_signal(24000 * 6, 1), _signal(24000 * 4, 2)]
    spec = fe.from_recordings(recs, [[100, 180, 250], [10, 120, 230]], audio_length=1.279)
    assert spec.shape == (6, 1, 256, 256)
    from oracle.stft_oracle import log_power_spectrogram
    ref = log_power_spectrogram(audio.extract_segment(recs[1], 230, audio_length=1.279))
    loud = ref > ref.max() - 10
    assert np.abs(spec[5, 0].double().cpu().numpy() - ref)[loud].max() < 2e-2
