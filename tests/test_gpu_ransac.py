"""dgr_ransac_correspondence (the reference's safeguard, core/deep_global_registration.py:50-64,
:302-315) against oracle/ransac.py.  Both sides draw the same hypotheses (counter-hash sampler),
so the search itself is comparable exactly: same winner, same pose to fp64 round-off; the
inlier test runs in fp32 on the GPU and fp64 in the oracle, so counts may differ by the few
correspondences that sit within ~1e-6 of the threshold."""
import numpy as np
import pytest
import torch

from deepglobalregistration_b200 import synthetic as syn
from oracle import ransac as orn

pytestmark = pytest.mark.gpu
make_corr = syn.correspondence_set


def run(P, tgt, i0, i1, max_dist, num_hyp, seed):
  from deepglobalregistration_b200 import _abi
  dev = torch.device('cuda')
  t = lambda a, dt: None if a is None else torch.as_tensor(a).to(dev, dt).contiguous()
  return _abi.ransac_correspondence(t(P, torch.float32), t(tgt, torch.float32), t(i0, torch.int32),
                                    t(i1, torch.int32), max_dist, num_hyp=num_hyp, seed=seed).cpu().numpy()


@pytest.mark.parametrize('seed,n,frac', [(1, 1500, 0.3), (2, 4000, 0.1), (3, 700, 0.6)])
def test_same_winner_as_oracle(seed, n, frac):
  P, tgt, i0, i1, T_gt, inl = make_corr(seed, n=n, inlier_frac=frac)
  num_hyp = 6000 if frac >= 0.3 else 40000
  res = run(P, tgt, i0, i1, 0.05, num_hyp, seed + 10)
  T = res[:16].reshape(4, 4)
  T_o, info = orn.ransac_correspondence(P, tgt, i0, i1, 0.05, num_hyp, seed=seed + 10)
  # the GPU's winner is the pose of ITS four draws under the oracle's sampler (checks the hash)
  hyp = int(res[18])
  s = orn.sample_indices(seed + 10, [hyp], n)
  R, t = orn.kabsch_batch(P[i0][s].astype(np.float64), tgt[i1][s].astype(np.float64))
  np.testing.assert_allclose(T[:3, :3], R[0], atol=1e-9)
  np.testing.assert_allclose(T[:3, 3], t[0], atol=1e-9)
  assert np.array_equal(T[3], [0, 0, 0, 1])
  # and it is as good as the oracle's best: identical hypothesis unless fp32 / fp64 disagree on a
  # threshold-straddling correspondence
  mine = orn.count_inliers(T, P, tgt, i0, i1, 0.05)
  assert abs(mine - int(res[19])) <= 2 and abs(res[16] - mine / n) <= 1e-12
  assert hyp == info['hypothesis'] or mine >= info['inliers'] - 2, (hyp, info, mine)
  if hyp == info['hypothesis']:
    assert abs(res[17] - info['inlier_rmse']) <= 1e-6
  te, re = syn.rte_rre(T, T_gt)
  assert te < 0.05 and re < 0.05, (te, re)


def test_reproducible_and_seeded():
  P, tgt, i0, i1, _, _ = make_corr(5)
  a = run(P, tgt, i0, i1, 0.05, 5000, 1)
  b = run(P, tgt, i0, i1, 0.05, 5000, 1)
  c = run(P, tgt, i0, i1, 0.05, 5000, 2)
  assert np.array_equal(a, b)
  assert a[18] != c[18]
  # a null idx0 means arange
  assert np.array_equal(a, run(P, tgt, None, i1, 0.05, 5000, 1))


def test_hypothesis_counts_not_multiple_of_the_block():
  """1, 1023, 1025 hypotheses: the tail threads of the last block must not win."""
  P, tgt, i0, i1, _, _ = make_corr(6, n=300, inlier_frac=0.5)
  for num_hyp in (1, 1023, 1025):
    res = run(P, tgt, i0, i1, 0.05, num_hyp, 3)
    T_o, info = orn.ransac_correspondence(P, tgt, i0, i1, 0.05, num_hyp, seed=3)
    assert -1 <= int(res[18]) < num_hyp
    assert int(res[18]) == info['hypothesis'] or abs(int(res[19]) - info['inliers']) <= 2


def test_degenerate_inputs():
  # nothing can be an inlier -> identity, hypothesis -1 (open3d's initial best result)
  P, tgt, i0, i1, _, _ = make_corr(7, n=200, inlier_frac=0.0)
  res = run(P, tgt, i0, i1, 1e-9, 2048, 0)
  assert np.array_equal(res[:16].reshape(4, 4), np.eye(4)) and res[18] == -1 and res[16] == 0 and res[17] == 0
  # one correspondence: every draw is the same point -> pure translation onto it
  res = run(P[:1], tgt, None, i1[:1], 0.05, 64, 0)
  T = res[:16].reshape(4, 4)
  np.testing.assert_allclose(T[:3, :3], np.eye(3), atol=1e-12)
  np.testing.assert_allclose(T[:3, 3], tgt[i1[0]].astype(np.float64) - P[0].astype(np.float64), atol=1e-6)
  assert res[16] == 1.0 and int(res[19]) == 1
  # argument checks come back as errors, not crashes
  from deepglobalregistration_b200 import _abi
  with pytest.raises(_abi.DgrError):
    run(P, tgt, i0, i1, -1.0, 16, 0)
  with pytest.raises(_abi.DgrError):
    run(P, tgt, i0, i1, 0.05, 0, 0)


def test_reference_size_search_time():
  """The reference's effective setting: 4 M hypotheses over ~50k correspondences."""
  P, tgt, i0, i1, T_gt, _ = make_corr(8, n=50000, inlier_frac=0.05)
  run(P, tgt, i0, i1, 0.1, 4096, 0)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  res = run(P, tgt, i0, i1, 0.1, 4000000, 0)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  print(f'\n[ransac] 4e6 hypotheses x 50000 correspondences: {ms:.1f} ms = '
        f'{4e6 * 5e4 / ms / 1e9:.1f} G evaluations/ms... inliers {int(res[19])}')
  te, re = syn.rte_rre(res[:16].reshape(4, 4), T_gt)
  assert te < 0.05 and re < 0.05
  assert ms < 2000
