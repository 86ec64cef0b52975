"""Attention back-end selector kept for API compatibility
(/root/reference/src/UCF_VIT/utils/fused_attn.py:12-16).

On B200 every member routes to the same hand-written tcgen05 flash-attention kernel
(`ucf_attention_fwd/bwd`): there is no xFormers / ROCm-CK / SDPA dispatch in this package."""
from enum import Enum


class FusedAttn(Enum):
    FLASH = "FLASH"
    CK = "CK"
    DEFAULT = "DEFAULT"
    NONE = "NONE"
