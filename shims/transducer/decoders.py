"""Import-compatible stand-in for awni/transducer's `transducer.decoders`
(reference import: speech/models/transducer_model.py:10)."""
from speech_b200.transducer_decoders import decode_static  # noqa: F401
