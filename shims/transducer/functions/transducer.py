"""Import-compatible stand-in for awni/transducer's `transducer.functions.transducer`
(reference import: speech/models/transducer_model.py:11)."""
from speech_b200.functions.transducer import TransducerLoss  # noqa: F401
