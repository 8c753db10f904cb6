"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference featuriser.

Reference: speech/loader.py:152-166 (`log_specgram`) and :65-69 (`Preprocessor.preprocess`:
`(log_specgram - mean) / std`).  The reference calls `scipy.signal.spectrogram(audio, fs,
window='hann', nperseg, noverlap, detrend=False)` whose defaults are scaling='density',
mode='psd', one-sided, no boundary extension, no padding.  This file restates that arithmetic
explicitly in float64 numpy (so that the CUDA kernel has a formula to be checked against) and is
pinned two ways: against scipy itself on random input and against the golden output of the
reference's own function on its own fixture `tests/test0.wav` (tests/golden/specgram.npz, made by
tests/golden/make_specgram_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import it.
"""
import numpy as np


def hann_periodic(n):
    """scipy.signal.get_window('hann', n) (fftbins=True => periodic)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def log_specgram(audio, sample_rate, window_size=20, step_size=10, eps=1e-10):
    """(frames, nperseg//2+1) float32 = log(PSD + eps); loader.py:156-166."""
    nperseg = int(window_size * sample_rate / 1e3)
    noverlap = int(step_size * sample_rate / 1e3)
    step = nperseg - noverlap
    x = np.asarray(audio).astype(np.float64)
    n_frames = (x.shape[0] - noverlap) // step if x.shape[0] >= nperseg else 0
    win = hann_periodic(nperseg)
    scale = 1.0 / (sample_rate * np.sum(win * win))
    idx = np.arange(nperseg)[None, :] + step * np.arange(n_frames)[:, None]
    frames = x[idx] * win[None, :]
    spec = np.abs(np.fft.rfft(frames, axis=1)) ** 2 * scale
    if nperseg % 2 == 0:
        spec[:, 1:-1] *= 2.0          # one-sided: every bin but DC and Nyquist counts twice
    else:
        spec[:, 1:] *= 2.0
    return np.log(spec.astype(np.float32) + eps)


def preprocess(audio, sample_rate, mean, std):
    """loader.py:65-67: normalised features."""
    return (log_specgram(audio, sample_rate) - mean) / std
