"""Mirror of reference models/model.py (LFAN, the shipped default: default_config.py:66)."""
from ..modules import LFAN  # noqa: F401
