"""oracle/ransac.py (open3d registration_ransac_based_on_correspondence restated; parity
unpinned - open3d is not installable here): known-answer recovery, the best-hypothesis rule,
the counter-hash sampler, and Kabsch against the ICP oracle's."""
import numpy as np

from deepglobalregistration_b200 import synthetic as syn
from oracle import icp as oicp
from oracle import ransac as orn


make_corr = syn.correspondence_set


def test_sampler_is_uniform_in_range_and_reproducible():
  s = orn.sample_indices(3, np.arange(20000), 777)
  assert s.shape == (20000, 4) and s.min() >= 0 and s.max() < 777
  assert np.array_equal(s, orn.sample_indices(3, np.arange(20000), 777))
  assert not np.array_equal(s, orn.sample_indices(4, np.arange(20000), 777))
  hist = np.bincount(s.reshape(-1), minlength=777)
  assert abs(hist.mean() - 80000 / 777) < 1e-9 and hist.min() > 50 and hist.max() < 170
  # a hypothesis draws the same correspondences wherever it sits in the batch
  assert np.array_equal(orn.sample_indices(3, [4242], 777)[0], s[4242])


def test_kabsch_batch_matches_the_icp_oracle():
  g = np.random.default_rng(0)
  P, Q = g.normal(size=(5, 4, 3)), g.normal(size=(5, 4, 3))
  R, t = orn.kabsch_batch(P, Q)
  for b in range(5):
    R1, t1 = oicp.kabsch(P[b], Q[b])
    np.testing.assert_allclose(R[b], R1, atol=1e-12)
    np.testing.assert_allclose(t[b], t1, atol=1e-12)
    assert abs(np.linalg.det(R[b]) - 1) < 1e-12


def test_known_answer_and_best_rule():
  P, tgt, i0, i1, T_gt, inl = make_corr(1)
  T, info = orn.ransac_correspondence(P, tgt, i0, i1, 0.05, 4096, seed=5)
  te, re = syn.rte_rre(T, T_gt)
  assert te < 0.03 and re < 0.03, (te, re, info)
  assert info['inliers'] == orn.count_inliers(T, P, tgt, i0, i1, 0.05) >= 0.8 * inl.sum()
  # more hypotheses can only improve (inliers, then rmse) - the search is a prefix of the longer one
  T2, info2 = orn.ransac_correspondence(P, tgt, i0, i1, 0.05, 8192, seed=5)
  assert (info2['inliers'], -info2['inlier_rmse']) >= (info['inliers'], -info['inlier_rmse'])
  # the winner really is the pose of its four draws
  s = orn.sample_indices(5, [info['hypothesis']], len(i0))
  R, t = orn.kabsch_batch(P[i0][s].astype(np.float64), tgt[i1][s].astype(np.float64))
  np.testing.assert_allclose(T[:3, :3], R[0], atol=1e-12)


def test_no_inlier_returns_identity():
  P, tgt, i0, i1, _, _ = make_corr(2, n=200, inlier_frac=0.0)
  T, info = orn.ransac_correspondence(P, tgt, i0, i1, 1e-9, 256)
  assert np.array_equal(T, np.eye(4)) and info['hypothesis'] == -1 and info['fitness'] == 0.0
