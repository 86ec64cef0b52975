"""`UCF_VIT.fsdp.arch` surface (/root/reference/src/UCF_VIT/fsdp/arch.py): the simple-mode
models with `tensor_par_size` / `tensor_par_group` accepted (must be 1 / ignored).  At
tensor_par_size == 1 both reference copies compute the same function; the fsdp
copy's A2/A4/A16 defects are not reproduced.  Blocks default to `fsdp.building_blocks.Block`
so `transformer_auto_wrap_policy({Block, Sequential})` and activation checkpointing find them."""
from ..simple import arch as _a
from ..simple.arch import feature_take_indices, global_pool_nlc, init_weights_vit_timm  # noqa: F401
from .building_blocks import Block, Mlp, PatchEmbed, _check_tp


def _strip_tp(kwargs):
    _check_tp(kwargs.pop('tensor_par_size', 1))
    kwargs.pop('tensor_par_group', None)
    kwargs.setdefault('block_fn', Block)
    kwargs.setdefault('embed_layer', PatchEmbed)
    kwargs.setdefault('mlp_layer', Mlp)
    return kwargs


class VIT(_a.VIT):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **_strip_tp(kwargs))


class SAP(_a.SAP):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **_strip_tp(kwargs))


class MAE(_a.MAE):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **_strip_tp(kwargs))


class UNETR(_a.UNETR):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **_strip_tp(kwargs))


class DiffusionVIT(_a.DiffusionVIT):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **_strip_tp(kwargs))
