"""The model classes `train.py` / `eval.py` look up by name (`eval("models." + cfg["model"]["class"])`,
reference train.py:88-91): same names as `speech.models`."""
from . import ctc_model, model, seq2seq, transducer_model

Model = model.Model
CTC = ctc_model.CTC
Seq2Seq = seq2seq.Seq2Seq
Transducer = transducer_model.Transducer

__all__ = ["Model", "CTC", "Seq2Seq", "Transducer"]
