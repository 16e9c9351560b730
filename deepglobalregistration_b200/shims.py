"""Register the drop-in modules the reference imports by name.

    from deepglobalregistration_b200 import shims; shims.install()
    import MinkowskiEngine as ME          # -> deepglobalregistration_b200.me
    from easydict import EasyDict         # -> attribute dict (checkpoint configs unpickle)

    import open3d as o3d                  # -> only if the real one is missing: the few names
                                          #    demo.py needs (io.read_point_cloud, geometry.PointCloud,
                                          #    utility.Vector3dVector, visualization.draw_geometries)

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
  if force or 'open3d' not in sys.modules:
    try:
      import open3d  # noqa: F401
    except ImportError:
      sys.modules['open3d'] = _open3d_stub()
  return me


def _open3d_stub():
  """The part of open3d demo.py:10-48 touches, backed by io.py.  Registration (ICP, RANSAC) is NOT
  routed through here: DeepGlobalRegistration calls libdgr_b200 for those."""
  import numpy as np

  from . import io as dio
  o3d = types.ModuleType('open3d')
  o3d.__dgr_stub__ = True
  o3d.io = types.ModuleType('open3d.io')
  o3d.io.read_point_cloud = dio.read_point_cloud
  o3d.io.write_point_cloud = lambda path, pcd, **kw: (dio.write_ply(path, pcd.points, dtype='double'), True)[1]
  o3d.geometry = types.ModuleType('open3d.geometry')
  o3d.geometry.PointCloud = dio.PointCloud
  o3d.utility = types.ModuleType('open3d.utility')
  o3d.utility.Vector3dVector = lambda a: np.asarray(a, dtype=np.float64).reshape(-1, 3)
  o3d.utility.VerbosityLevel = types.SimpleNamespace(Error=0, Warning=1, Info=2, Debug=3)
  o3d.utility.set_verbosity_level = lambda level: None
  o3d.visualization = types.ModuleType('open3d.visualization')

  def draw_geometries(geometries, *args, **kwargs):
    print('[open3d stub] draw_geometries: ' + ', '.join(repr(g) for g in geometries) + ' (no display)')
  o3d.visualization.draw_geometries = draw_geometries
  for sub in ('io', 'geometry', 'utility', 'visualization'):
    sys.modules['open3d.' + sub] = getattr(o3d, sub)
  return o3d
