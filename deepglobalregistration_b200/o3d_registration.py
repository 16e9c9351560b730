"""``open3d.pipelines.registration`` stand-in backed by libdgr_b200 - the two calls the reference's
``core/deep_global_registration.py`` makes into open3d, with open3d's signatures and result objects:

  registration_icp(source, target, max_correspondence_distance, init, ...)             (:317-322)
      -> dgr_icp_point_to_point: nearest target point through a voxel hash of the target, fp64 Kabsch
         update, open3d's default stopping rule (relative fitness / RMSE 1e-6, 30 iterations);
  registration_ransac_based_on_correspondence(source, target, corres, ...)              (:50-64)
      -> dgr_ransac_correspondence: criteria.max_iteration four-point hypotheses, each scored on all
         correspondences (the reference passes its 80000 into the confidence slot, so open3d >= 0.12 never
         exits early; under 0.10 / 0.11 the same argument bounded the VALIDATED hypotheses instead - this
         stand-in follows the >= 0.12 reading).

With ``shims.install()`` these are reachable as ``open3d.pipelines.registration`` (and the pre-0.12 alias
``open3d.registration``) whenever the real open3d is absent, so the reference's own class runs on this stack
without a line changed.  There is no CPU path: the functions raise without an sm_100 device.
"""
import numpy as np
import torch

from . import _abi


class TransformationEstimationPointToPoint:
  def __init__(self, with_scaling=False):
    if with_scaling:
      raise NotImplementedError('with_scaling=True is not used by DGR and not built')
    self.with_scaling = False


class ICPConvergenceCriteria:
  def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
    self.relative_fitness, self.relative_rmse, self.max_iteration = relative_fitness, relative_rmse, max_iteration


class RANSACConvergenceCriteria:
  def __init__(self, max_iteration=100000, confidence=0.999):
    self.max_iteration = int(max_iteration)
    self.confidence = min(float(confidence), 1.0)       # open3d clamps; DGR passes 80000 here


class CorrespondenceCheckerBasedOnDistance:
  def __init__(self, distance_threshold):
    self.distance_threshold = distance_threshold


class RegistrationResult:
  def __init__(self, transformation, fitness=0.0, inlier_rmse=0.0, n_corr=0):
    self.transformation = np.asarray(transformation, dtype=np.float64).reshape(4, 4).copy()
    self.fitness, self.inlier_rmse = float(fitness), float(inlier_rmse)
    self.correspondence_set = np.zeros((int(n_corr), 2), dtype=np.int32)   # count only; pairs stay on the device

  def __repr__(self):
    return (f'RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e}, and '
            f'correspondence_set size of {len(self.correspondence_set)}')


def _points(pcd, device):
  pts = np.asarray(getattr(pcd, 'points', pcd), dtype=np.float64).reshape(-1, 3)
  return torch.from_numpy(np.ascontiguousarray(pts)).to(device)


def _target_hash(tgt64, max_dist):
  """Voxel hash of the target with at most one point per cell (what the ICP kernel searches): cell =
  max_dist / 2 as in DGR (voxelised clouds, radius 2 voxels); a cloud with several points per cell gets finer
  cells up to the kernel's reach of 4."""
  from .me.coords import KEY_MARGIN
  for div in (2.0, 3.0, 4.0):
    cell = max_dist / div
    raw, minmax = _abi.quantize_points(tgt64, cell)
    spec = _abi.keyspec_build(minmax, 4, KEY_MARGIN)
    table, _, _, cnt = _abi.unique_first(raw, spec)
    if _abi.read_count(cnt) == tgt64.shape[0]:
      return cell, spec, table
  raise NotImplementedError('target has several points within max_correspondence_distance / 4 of each other: '
                            'voxel-downsample it first (DGR always passes voxelised clouds)')


def registration_icp(source, target, max_correspondence_distance, init=None, estimation_method=None, criteria=None):
  dev = _abi.require_device('cuda')
  _abi.refresh_stream()
  criteria = criteria or ICPConvergenceCriteria()
  if isinstance(estimation_method, TransformationEstimationPointToPoint) or estimation_method is None:
    pass
  else:
    raise NotImplementedError('only point-to-point ICP is built (what DGR calls)')
  src64, tgt64 = _points(source, dev), _points(target, dev)
  T0 = np.eye(4) if init is None else np.asarray(init, dtype=np.float64).reshape(4, 4)
  if len(src64) == 0 or len(tgt64) == 0:
    return RegistrationResult(T0)
  cell, spec, table = _target_hash(tgt64, float(max_correspondence_distance))
  src, tgt = src64.float().contiguous(), tgt64.float().contiguous()
  T12 = torch.from_numpy(np.ascontiguousarray(T0[:3])).to(dev)
  state = torch.empty(64, dtype=torch.float64, device=dev)
  res = torch.empty(20, dtype=torch.float64, device=dev)
  _abi.call('dgr_icp_point_to_point', _abi.ptr(src), src.shape[0], _abi.ptr(tgt), _abi.ptr(spec), _abi.ptr(table.keys),
            _abi.ptr(table.vals), table.cap, 0, float(cell), float(max_correspondence_distance), _abi.ptr(T12),
            int(criteria.max_iteration), float(criteria.relative_fitness), float(criteria.relative_rmse),
            _abi.ptr(state), _abi.ptr(res), _abi.stream())
  r = res.cpu().numpy()
  return RegistrationResult(r[:16], r[16], r[17], r[19])


def registration_ransac_based_on_correspondence(source, target, corres, max_correspondence_distance,
                                                estimation_method=None, ransac_n=3, checkers=None, criteria=None,
                                                seed=0):
  dev = _abi.require_device('cuda')
  _abi.refresh_stream()
  if int(ransac_n) != 4:
    raise NotImplementedError('ransac_n = 4 (what DGR passes) is the built sample size')
  criteria = criteria or RANSACConvergenceCriteria()
  corres = np.asarray(corres).reshape(-1, 2)
  src, tgt = _points(source, dev).float().contiguous(), _points(target, dev).float().contiguous()
  if len(corres) < 4:
    return RegistrationResult(np.eye(4))
  idx0 = torch.from_numpy(np.ascontiguousarray(corres[:, 0], dtype=np.int32)).to(dev)
  idx1 = torch.from_numpy(np.ascontiguousarray(corres[:, 1], dtype=np.int32)).to(dev)
  r = _abi.ransac_correspondence(src, tgt, idx0, idx1, float(max_correspondence_distance),
                                 num_hyp=criteria.max_iteration, seed=seed).cpu().numpy()
  return RegistrationResult(r[:16], r[16], r[17], r[19])
