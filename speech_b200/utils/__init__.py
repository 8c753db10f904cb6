from .io import load, load_state, save, save_state   # noqa: F401
from .score import compute_cer, edit_distance        # noqa: F401
