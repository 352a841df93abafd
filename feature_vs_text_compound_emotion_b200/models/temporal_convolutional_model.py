"""Mirror of reference models/temporal_convolutional_model.py."""
from ..modules import Chomp1d, TemporalBlock, TemporalConvNet  # noqa: F401
