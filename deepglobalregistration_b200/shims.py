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
  """The part of open3d the reference touches on the registration path: demo.py:10-48 (I/O, PointCloud, a no-op
  viewer, backed by io.py) and core/deep_global_registration.py:50-64,317-322 + util/pointcloud.py:15-23
  (pipelines.registration.registration_icp / registration_ransac_based_on_correspondence, utility vectors) backed
  by libdgr_b200 (o3d_registration.py) - so the reference's OWN DeepGlobalRegistration class and demo.py run on
  this stack unmodified.  This package's DeepGlobalRegistration does not go through here: it calls the library."""
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
  o3d.utility.Vector2iVector = lambda a: np.asarray(a, dtype=np.int32).reshape(-1, 2)
  from . import o3d_registration as reg
  o3d.pipelines = types.ModuleType('open3d.pipelines')
  o3d.pipelines.registration = types.ModuleType('open3d.pipelines.registration')
  for name in ('TransformationEstimationPointToPoint', 'ICPConvergenceCriteria', 'RANSACConvergenceCriteria',
               'CorrespondenceCheckerBasedOnDistance', 'RegistrationResult', 'registration_icp',
               'registration_ransac_based_on_correspondence'):
    setattr(o3d.pipelines.registration, name, getattr(reg, name))
  o3d.registration = o3d.pipelines.registration          # the pre-0.12 module path
  sys.modules['open3d.pipelines'] = o3d.pipelines
  sys.modules['open3d.pipelines.registration'] = o3d.pipelines.registration
  sys.modules['open3d.registration'] = o3d.pipelines.registration
  o3d.utility.VerbosityLevel = types.SimpleNamespace(Error=0, Warning=1, Info=2, Debug=3)
  o3d.utility.set_verbosity_level = lambda level: None
  o3d.visualization = types.ModuleType('open3d.visualization')

  def draw_geometries(geometries, *args, **kwargs):
    print('[open3d stub] draw_geometries: ' + ', '.join(repr(g) for g in geometries) + ' (no display)')
  o3d.visualization.draw_geometries = draw_geometries
  for sub in ('io', 'geometry', 'utility', 'visualization'):
    sys.modules['open3d.' + sub] = getattr(o3d, sub)
  return o3d
