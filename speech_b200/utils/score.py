"""Scoring - host-side mirror of speech/utils/score.py:7-18 (reference), SURVEY.md section 8f
rank 4.  The reference depends on the `editdistance` C extension; the Levenshtein distance here
is the library's host function `sb_edit_distance` (csrc/api.cu)."""
import ctypes

from .. import _lib


def edit_distance(a, b):
    """Levenshtein distance of two sequences of hashable tokens (`editdistance.eval(a, b)`)."""
    lib = _lib.load()
    ids = {}
    ia = [ids.setdefault(t, len(ids)) for t in a]
    ib = [ids.setdefault(t, len(ids)) for t in b]
    ca = (ctypes.c_int * max(len(ia), 1))(*ia)
    cb = (ctypes.c_int * max(len(ib), 1))(*ib)
    d = lib.sb_edit_distance(ca, len(ia), cb, len(ib))
    if d < 0:
        raise _lib.SpeechB200Error("sb_edit_distance: invalid arguments")
    return int(d)


def compute_cer(results):
    """results: list of (ground truth, prediction) sequence pairs -> total edit distance divided
    by the total label length (score.py:7-18)."""
    dist = sum(edit_distance(label, pred) for label, pred in results)
    total = sum(len(label) for label, _ in results)
    return dist / total
