"""Mirror of reference models/arcface_model.py (the symbols on the LFAN path)."""
from ..modules import Backbone, Flatten, bottleneck_IR, get_blocks  # noqa: F401
