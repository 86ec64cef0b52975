from . import _impl
get_conv_layer, UnetOutBlock, UnetResBlock, UnetBasicBlock = (
    _impl.get_conv_layer, _impl.UnetOutBlock, _impl.UnetResBlock, _impl.UnetBasicBlock)
