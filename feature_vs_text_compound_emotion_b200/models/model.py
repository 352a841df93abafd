"""Mirror of reference models/model.py: LFAN (the shipped default, default_config.py:66) and the
alternative heads CAN / JMT / MT with their fusion modules."""
from ..heads import (CAN, JMT, AttentionFusion, JMTFusion, MTFusion, SequentialEncoder,  # noqa: F401
                     TransformerEncoderBlock, TransformerEncoderLayer)
from ..modules import LFAN  # noqa: F401
