"""Host-side mirror of `speech.models` (reference speech/models/__init__.py:2-5)."""
from .model import Model
from .seq2seq import Seq2Seq
from .ctc_model import CTC
from .transducer_model import Transducer

__all__ = ["Model", "Seq2Seq", "CTC", "Transducer"]
