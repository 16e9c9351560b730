"""ORACLE (test infrastructure only) - point-to-point ICP as the reference calls it:
``o3d.pipelines.registration.registration_icp(source, target, max_correspondence_distance=2*voxel,
init=T)`` (core/deep_global_registration.py:317-322) with open3d's defaults
(TransformationEstimationPointToPoint(with_scaling=False),
ICPConvergenceCriteria(relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30)).

PARITY UNPINNED: open3d (requirements.txt:42) is not installed in the build container and
not vendored, so this restates open3d's published RegistrationICP loop: evaluate
correspondences (nearest target point within the radius, KD-tree), then up to 30 times
{Kabsch update on the current correspondences, re-evaluate, stop when both fitness and
inlier RMSE change by less than 1e-6}.  All arithmetic in float64, as open3d does."""
import numpy as np
from scipy.spatial import cKDTree


def kabsch(P, Q):
  """Rigid (R, t) minimising sum |R p + t - q|^2 (Umeyama without scaling), float64."""
  mp, mq = P.mean(0), Q.mean(0)
  S = (Q - mq).T @ (P - mp) / len(P)
  U, _, Vt = np.linalg.svd(S)
  d = np.ones(3)
  if np.linalg.det(U) * np.linalg.det(Vt) < 0:
    d[2] = -1.0
  R = U @ np.diag(d) @ Vt
  return R, mq - R @ mp


def icp_point_to_point(src, tgt, max_dist, T_init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
  src, tgt = np.asarray(src, np.float64), np.asarray(tgt, np.float64)
  T = np.eye(4) if T_init is None else np.array(T_init, np.float64)
  pts = src @ T[:3, :3].T + T[:3, 3]
  tree = cKDTree(tgt)

  def evaluate(p):
    d, j = tree.query(p, k=1, distance_upper_bound=max_dist)
    m = np.isfinite(d)
    n = int(m.sum())
    return m, j, (n / len(p) if len(p) else 0.0), (float(np.sqrt((d[m] ** 2).sum() / n)) if n else 0.0)

  m, j, fit, rmse = evaluate(pts)
  it = 0
  for it in range(1, max_iter + 1):
    if m.any():
      R, t = kabsch(pts[m], tgt[j[m]])
    else:
      R, t = np.eye(3), np.zeros(3)
    U = np.eye(4)
    U[:3, :3], U[:3, 3] = R, t
    T = U @ T
    pts = pts @ R.T + t
    pf, pr = fit, rmse
    m, j, fit, rmse = evaluate(pts)
    if abs(pf - fit) < rel_fitness and abs(pr - rmse) < rel_rmse:
      break
  return T, dict(fitness=fit, inlier_rmse=rmse, iterations=it, n_corr=int(m.sum()))
