"""``demcoreg``: the reference's import name (reference ``demcoreg/__init__.py:3``: ``__all__ = ['coreglib', 'modis_grid']``),
resolved to the B200-native modules so that ``from demcoreg import coreglib`` and ``import demcoreg.dem_align`` work on an
unedited caller.  Only the hot-path modules exist (``coreglib``, ``dem_align``, ``robust_stats``, ``compute_diff``, and the
raster pieces of ``dem_mask``)."""
import importlib
import sys

__all__ = ['coreglib', 'dem_align', 'robust_stats', 'compute_diff', 'dem_mask']

for _name in __all__:
    _mod = importlib.import_module('demcoreg_b200.' + _name)
    sys.modules[__name__ + '.' + _name] = _mod
    globals()[_name] = _mod
del _name, _mod
