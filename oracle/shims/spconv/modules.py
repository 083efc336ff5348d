from .pytorch.modules import is_spconv_module  # noqa: F401
