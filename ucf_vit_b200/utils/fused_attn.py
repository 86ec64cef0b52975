"""Attention back-end selector kept for API compatibility
(/root/reference/src/UCF_VIT/utils/fused_attn.py:12-16).

On B200 every member routes to the same hand-written tcgen05 flash-attention kernel
(`ucf_attention_fwd/bwd`): there is no xFormers / ROCm-CK / SDPA dispatch in this package.  The
members exist so that configs and call sites written for the reference (`FusedAttn.FLASH`,
`FusedAttn["CK"]`, `FusedAttn("DEFAULT")`) keep working unchanged."""
import enum

_BACKENDS = ("FLASH", "CK", "DEFAULT", "NONE")

# name == value for every member, as the reference's YAML configs select them by string
FusedAttn = enum.Enum("FusedAttn", [(b, b) for b in _BACKENDS], module=__name__)
FusedAttn.__doc__ = "Which fused-attention back end the caller asked for (all map to the CUDA kernel here)."
