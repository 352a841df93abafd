"""Import-path shim: the same module names as the reference's ``models`` package, so
``from models.model import LFAN`` becomes
``from feature_vs_text_compound_emotion_b200.models.model import LFAN``."""
