"""Mirror of reference models/backbone.py (VGG / VGGish / VisualBackbone / AudioBackbone)."""
from ..modules import AudioBackbone, Flatten, VGG, VGGish, VisualBackbone, _vgg, make_layers  # noqa: F401
