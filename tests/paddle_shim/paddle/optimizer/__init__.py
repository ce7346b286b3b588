"""paddle.optimizer stand-in: only the lr namespace the reference's scheduler.py touches."""
from . import lr  # noqa: F401
