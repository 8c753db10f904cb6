"""Build the oracle's C restatement (oracle/ctc_ref.c) -> oracle/_build/liboracle.so (gcc).
TEST INFRASTRUCTURE ONLY; never linked or loaded by speech_b200/."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liboracle.so")


def build(force=False):
    src = os.path.join(HERE, "ctc_ref.c")
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", src, "-o", LIB, "-lm"])
    return LIB


def load():
    lib = ctypes.CDLL(build())
    lib.oracle_ctc.restype = ctypes.c_int
    return lib


def ctc(acts, labels, label_lens, act_lens, blank, need_grad=True):
    import numpy as np
    lib = load()
    acts = np.ascontiguousarray(acts, np.float32)
    B, T, V = acts.shape
    labels = np.ascontiguousarray(labels, np.int32)
    label_lens = np.ascontiguousarray(label_lens, np.int32)
    act_lens = np.ascontiguousarray(act_lens, np.int32)
    costs = np.zeros(B, np.float64)
    grads = np.zeros_like(acts) if need_grad else None
    P = ctypes.c_void_p
    lib.oracle_ctc(acts.ctypes.data_as(P), grads.ctypes.data_as(P) if need_grad else None,
                   labels.ctypes.data_as(P), label_lens.ctypes.data_as(P),
                   act_lens.ctypes.data_as(P), B, T, V, int(blank), costs.ctypes.data_as(P))
    return costs, grads


if __name__ == "__main__":
    print(build(force=True))
