"""Mirror of reference models/transformer.py (the four symbols LFAN uses)."""
from ..modules import (MultiModalEncoderBlock, MultimodalMultiheadAttention,  # noqa: F401
                       MultimodalTransformerEncoder)
