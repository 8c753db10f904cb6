"""Generates tests/golden/specgram.npz by running the REFERENCE's own `speech.loader.log_specgram`
(/root/reference/speech/loader.py:156-166, unmodified, imported in this container) on its own
fixtures tests/test0.wav and tests/test1.wav.  `soundfile` (the reference's wave reader) is not
installed here, so the int16 samples are read with scipy.io.wavfile and a stub module satisfies
the import; the function under test does not touch it.

    python tests/golden/make_specgram_golden.py
"""
import os
import sys
import types
import warnings

import numpy as np
import scipy.io.wavfile

REF = "/root/reference"
for _absent in ("soundfile", "editdistance"):      # un-installed deps the function never touches
    sys.modules.setdefault(_absent, types.ModuleType(_absent))
sys.path.insert(0, REF)
from speech import loader  # noqa: E402

out = {}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for name in ("test0", "test1"):
        sr, audio = scipy.io.wavfile.read(os.path.join(REF, "tests", name + ".wav"))
        assert audio.dtype == np.int16
        out[name + "_audio"] = audio
        out[name + "_sr"] = np.int64(sr)
        out[name + "_logspec"] = loader.log_specgram(audio, sr)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "specgram.npz"), **out)
for k, v in out.items():
    print(k, getattr(v, "shape", v), getattr(v, "dtype", ""))
