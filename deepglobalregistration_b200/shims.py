"""Register the drop-in modules the reference imports by name.

    from deepglobalregistration_b200 import shims; shims.install()
    import MinkowskiEngine as ME          # -> deepglobalregistration_b200.me
    from easydict import EasyDict         # -> attribute dict (checkpoint configs unpickle)

After install() the reference's own ``model/resunet.py`` / ``model/residual_block.py`` /
``model/common.py`` import and run unchanged on top of libdgr_b200 (the plugin
boundary of SURVEY.md §8b).
"""
import sys
import types

from .synthetic import AttrDict


def install(force=False):
  from . import me
  if force or 'MinkowskiEngine' not in sys.modules:
    sys.modules['MinkowskiEngine'] = me
    sys.modules['MinkowskiEngine.MinkowskiFunctional'] = me.MinkowskiFunctional
    sys.modules['MinkowskiEngine.utils'] = me.utils
  if force or 'easydict' not in sys.modules:
    try:
      import easydict  # noqa: F401
    except ImportError:
      mod = types.ModuleType('easydict')
      mod.EasyDict = AttrDict
      sys.modules['easydict'] = mod
  return me
