"""B200-native LFAN inference hot path (IR-50 -> TCN -> cross-modal attention -> classifier).

Drop-in mirrors of the reference's ``models/*`` nn.Modules whose ``forward`` runs
hand-written sm_100a CUDA through the C-ABI library ``libcer_b200.so`` (include/cer_b200.h).
There is no CPU fallback: importing the modules works anywhere (so state_dicts can be built
and inspected), but any compute call raises if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
