"""ocrl_b200 -- B200-native (sm_100a) slot-attention hot path of OCRL behind the reference's
own OCR interface.  ``SLATE`` / ``SLATE_Module`` are drop-ins for ``ocrs.SLATE`` /
``ocrs.SLATE_Module``."""
from .slot_attn import SlotAttention, SlotAttentionEncoder  # noqa: F401
from .feature_stage import PositionalEmbedding, SlotAttnCNNEncoder  # noqa: F401
from .slate_module import SLATE_Module  # noqa: F401
from .slate import SLATE, Base  # noqa: F401
from .graphed import GraphedEncoder, StreamedEncoder  # noqa: F401
from .pooling import RolloutExtractor, Transformer_Module  # noqa: F401

__all__ = ["SLATE", "SLATE_Module", "Base", "SlotAttention", "SlotAttentionEncoder", "SlotAttnCNNEncoder",
           "PositionalEmbedding", "GraphedEncoder", "StreamedEncoder", "Transformer_Module", "RolloutExtractor"]
