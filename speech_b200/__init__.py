"""speech_b200 - B200-native (sm_100a) replacement for the data-parallel hot path of awni/speech.

Layout:
    csrc/        hand-written CUDA kernels + the C ABI (include/speech_b200.h), built in-tree
    _lib.py      ctypes binding of that ABI (no CPU fallback)
    functions/   operator-level drop-ins: functions.ctc.CTCLoss, transducer.* (reference imports)
    models/      host-side mirror of speech.models.{Model,CTC,Seq2Seq,Transducer}
"""
__version__ = "0.1.0"
