"""ORACLE (test infrastructure only) - the safeguard branch of the reference:
``o3d.pipelines.registration.registration_ransac_based_on_correspondence(source, target, corres,
max_correspondence_distance=2*voxel, TransformationEstimationPointToPoint(False), ransac_n=4,
criteria=RANSACConvergenceCriteria(4000000, num_iterations))``
(core/deep_global_registration.py:50-64, called from :302-315 with num_iterations=80000).

PARITY UNPINNED: open3d (requirements.txt:42) is not installed in the build container and not
vendored, so this restates open3d's published RANSAC loop: per iteration draw ransac_n
correspondences uniformly (with replacement), estimate the rigid pose from them (Umeyama without
scaling), score it on ALL correspondences (inlier: |T p - q| < max_distance; fitness = inliers /
correspondences, inlier_rmse = sqrt(sum d^2 / inliers)), keep it if fitness is higher or equal
with a lower RMSE.  open3d stops early once ``log(1 - confidence) / log(1 - fitness^ransac_n)``
iterations have run; the reference passes 80000 in the confidence slot, which open3d clamps to
1.0, so the bound is infinite and all max_iteration hypotheses are evaluated.

open3d seeds one mt19937 per OpenMP thread from std::random_device, i.e. the reference's draw is
not reproducible; parity with it can only be statistical.  What CAN be exact is the search given
the draws, so the sampler here is the counter hash the CUDA kernel uses (csrc/ransac.cu
ransac_pick): same (seed, hypothesis, slot) -> same correspondence, on both sides.  All
arithmetic in float64, as open3d does."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
  x = x.astype(np.uint64).copy()
  x ^= x >> np.uint64(33)
  x *= np.uint64(0xff51afd7ed558ccd)
  x ^= x >> np.uint64(33)
  x *= np.uint64(0xc4ceb9fe1a85ec53)
  x ^= x >> np.uint64(33)
  return x


def sample_indices(seed, hyp, n):
  """[len(hyp), 4] correspondence numbers of the given hypotheses."""
  hyp = np.asarray(hyp, np.uint64)[:, None]
  slot = np.arange(4, dtype=np.uint64)[None, :]
  with np.errstate(over='ignore'):
    z = _mix64(np.uint64(seed) + (hyp * np.uint64(4) + slot + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
  return (((z >> np.uint64(32)) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def kabsch_batch(P, Q):
  """P, Q [B, m, 3] -> R [B, 3, 3], t [B, 3] minimising sum |R p + t - q|^2 per batch entry."""
  mp, mq = P.mean(1), Q.mean(1)
  S = np.einsum('bmr,bmc->brc', Q - mq[:, None], P - mp[:, None]) / P.shape[1]
  U, _, Vt = np.linalg.svd(S)
  d = np.ones((len(P), 3))
  d[:, 2] = np.where(np.linalg.det(U) * np.linalg.det(Vt) < 0, -1.0, 1.0)
  R = np.einsum('bij,bj,bjk->bik', U, d, Vt)
  return R, mq - np.einsum('bij,bj->bi', R, mp)


def score(R, t, P, Q, max_dist):
  """inliers [B], sum of squared inlier distances [B] of poses (R, t) on correspondences P -> Q."""
  d2 = ((np.einsum('bij,nj->bni', R, P) + t[:, None] - Q[None]) ** 2).sum(-1)
  m = np.sqrt(d2) < max_dist
  return m.sum(1), (d2 * m).sum(1)


def ransac_correspondence(src, tgt, idx0, idx1, max_dist, max_iteration, seed=0, chunk=512):
  """-> (T 4x4 float64, info).  Identity when no hypothesis has an inlier (open3d's initial
  best result)."""
  P = np.asarray(src, np.float64)[np.asarray(idx0)]
  Q = np.asarray(tgt, np.float64)[np.asarray(idx1)]
  n = len(P)
  best = (0, 0.0, -1, np.eye(3), np.zeros(3))      # inliers, err2, hypothesis, R, t
  for lo in range(0, max_iteration, chunk):
    hyp = np.arange(lo, min(lo + chunk, max_iteration))
    s = sample_indices(seed, hyp, n)
    R, t = kabsch_batch(P[s], Q[s])
    cnt, err = score(R, t, P, Q, max_dist)
    for b in np.flatnonzero(cnt >= max(best[0], 1)):
      # IsBetterRANSACThan: higher fitness, or the same with lower RMSE (strict, so the earliest
      # hypothesis keeps a tie)
      if cnt[b] > best[0] or err[b] < best[1]:
        best = (int(cnt[b]), float(err[b]), int(hyp[b]), R[b], t[b])
  T = np.eye(4)
  T[:3, :3], T[:3, 3] = best[3], best[4]
  info = dict(fitness=best[0] / n if n else 0.0, inlier_rmse=float(np.sqrt(best[1] / best[0])) if best[0] else 0.0,
              hypothesis=best[2], inliers=best[0])
  return T, info


def count_inliers(T, src, tgt, idx0, idx1, max_dist):
  P = np.asarray(src, np.float64)[np.asarray(idx0)]
  Q = np.asarray(tgt, np.float64)[np.asarray(idx1)]
  d = np.linalg.norm(P @ T[:3, :3].T + T[:3, 3] - Q, axis=1)
  return int((d < max_dist).sum())
