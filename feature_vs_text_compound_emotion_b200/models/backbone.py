"""Mirror of reference models/backbone.py (VisualBackbone; VGGish is a later scope row)."""
from ..modules import Flatten, VisualBackbone  # noqa: F401
