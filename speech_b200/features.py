"""GPU featuriser - host-side mirror of the reference's `log_specgram` + normalisation
(speech/loader.py:152-166 and :65-67), SURVEY.md section 8f rank 2.

    feats, n_frames = log_specgram_batch(audios, 16000, mean=preproc.mean, std=preproc.std)

`audios` is a list of int16 numpy arrays (what `speech.utils.wave.array_from_wave` returns); the
result is the (B, max frames, nperseg/2+1) float32 CUDA tensor the encoder consumes, zero padded
like `zero_pad_concat`, plus the true frame count of every utterance.  The arithmetic runs in
csrc/specgram.cu (direct DFT in float64); there is no CPU path.
"""
import numpy as np
import torch

from . import _lib, ops


def frame_count(n_samples, nperseg, step):
    """scipy.signal.spectrogram without boundary extension / padding (loader.py:159-164)."""
    noverlap = nperseg - step
    return (n_samples - noverlap) // step if n_samples >= nperseg else 0


def log_specgram_batch(audios, sample_rate, mean=None, std=None, window_size=20, step_size=10,
                       eps=1e-10, device=None):
    lib = _lib.load()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    nperseg = int(window_size * sample_rate / 1e3)       # loader.py:157-158
    noverlap = int(step_size * sample_rate / 1e3)
    step = nperseg - noverlap
    nbins = nperseg // 2 + 1
    B = len(audios)
    lens = [int(a.shape[0]) for a in audios]
    for a in audios:
        if a.dtype != np.int16 or a.ndim != 1:
            raise _lib.SpeechB200Error("log_specgram_batch expects 1-D int16 PCM arrays")
    n_frames = [frame_count(n, nperseg, step) for n in lens]
    max_frames = max(max(n_frames), 1)
    offs = np.zeros(B, dtype=np.int64)
    offs[1:] = np.cumsum(lens[:-1])
    total = int(sum(lens))
    host = torch.empty(max(total, 1), dtype=torch.int16).pin_memory()
    np.concatenate(audios, out=host.numpy()[:total]) if total else None
    pcm = host.to(device, non_blocking=True)
    d_off = torch.from_numpy(offs).to(device)
    d_len = torch.tensor(lens, dtype=torch.int32).to(device)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(nperseg) / nperseg)     # periodic Hann
    scale = 1.0 / (sample_rate * float(np.sum(win * win)))
    d_mean = d_std = None
    if mean is not None:
        d_mean = torch.as_tensor(np.asarray(mean, dtype=np.float32)).to(device)
        d_std = torch.as_tensor(np.asarray(std, dtype=np.float32)).to(device)
        if d_mean.numel() != nbins or d_std.numel() != nbins:
            raise _lib.SpeechB200Error("mean/std must have nperseg/2+1 = %d entries" % nbins)
    out = torch.empty(B, max_frames, nbins, dtype=torch.float32, device=device)
    sp = _lib.stream_ptr()
    ops._launch("log_specgram", 0.0,
                lambda: lib.sb_log_specgram(pcm.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), B,
                                            nperseg, step, scale, float(eps), _lib.ptr(d_mean),
                                            _lib.ptr(d_std), out.data_ptr(), max_frames, sp))
    return out, n_frames
