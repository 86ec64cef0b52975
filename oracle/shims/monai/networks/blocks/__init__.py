import importlib.util as _u, os as _o, sys as _s
_p = _o.path.join(_o.path.dirname(__file__), *[".."] * 5, "ucf_vit_b200", "utils", "unetr_blocks.py")
_spec = _u.spec_from_file_location("_unetr_blocks_for_shim", _o.path.abspath(_p))
_m = _u.module_from_spec(_spec); _spec.loader.exec_module(_m)
UnetrBasicBlock, UnetrPrUpBlock, UnetrUpBlock = _m.UnetrBasicBlock, _m.UnetrPrUpBlock, _m.UnetrUpBlock
UnetResBlock, UnetBasicBlock = _m.UnetResBlock, _m.UnetBasicBlock
_impl = _m
