"""Host-side mirror of `speech.models` (reference speech/models/__init__.py:2-5)."""
from .model import Model
from .ctc_model import CTC

__all__ = ["Model", "CTC"]
