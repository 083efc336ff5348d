def is_spconv_module(module):
    from . import SubMConv3d
    return isinstance(module, SubMConv3d)
