"""The oracle's non-ME stages against golden vectors produced by the reference's own
code (tests/golden/make_golden.py imports core/knn.py, core/registration.py,
core/loss.py from /root/reference).  This pins oracle/registration.py."""
import numpy as np
import pytest
import torch

from oracle import registration as oreg


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_knn_matches_reference(golden, tag):
  idx = oreg.feature_knn(golden[f'knn_{tag}_F0'], golden[f'knn_{tag}_F1'], nn_max_n=250)
  assert np.array_equal(idx.numpy(), golden[f'knn_{tag}_idx'])      # bit-exact indices


def test_knn_duplicate_rows_pick_lowest_index(golden):
  idx = oreg.feature_knn(golden['knn_a_F0'], golden['knn_a_F1'])
  assert idx[5].item() == 3                                         # F1[3] == F1[7] == F0[5]


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_weighted_procrustes_matches_reference(golden, tag):
  R, t = oreg.weighted_procrustes(golden[f'reg_{tag}_X'], golden[f'reg_{tag}_Y'], golden[f'reg_{tag}_w'])
  np.testing.assert_allclose(R.numpy(), golden[f'reg_{tag}_R_proc'], atol=2e-6)
  np.testing.assert_allclose(t.numpy(), golden[f'reg_{tag}_t_proc'], atol=2e-6)
  X, Y, w = (torch.from_numpy(golden[f'reg_{tag}_{k}']) for k in 'XYw')
  loss = oreg.robust_loss(X @ R.t() + t, Y, w, 0.1).item()
  assert abs(loss - float(golden[f'reg_{tag}_loss_proc'])) <= 1e-5 * max(1.0, abs(loss))


def test_rot6d_matches_reference(golden):
  for p, want in zip(golden['rot6d_in'], golden['rot6d_out']):
    got = oreg.rot6d_to_matrix(torch.from_numpy(p)).numpy()
    np.testing.assert_allclose(got, want, atol=1e-6)


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_refine_matches_reference(golden, tag):
  R, t, info = oreg.se3_refine(golden[f'reg_{tag}_X'], golden[f'reg_{tag}_Y'], golden[f'reg_{tag}_w'],
                               quantization_size=0.1)
  # tolerance of north_star: 1e-3 rad / 1e-3 m; the restatement is far tighter
  np.testing.assert_allclose(R.numpy(), golden[f'reg_{tag}_R_ref'], atol=1e-4)
  np.testing.assert_allclose(t.numpy(), golden[f'reg_{tag}_t_ref'], atol=1e-4)
  ref_it = int(golden[f'reg_{tag}_iters'])
  assert abs(info['iterations'] - ref_it) <= max(5, ref_it // 10)   # stop rule is fp-order sensitive
  assert abs(info['loss'] - float(golden[f'reg_{tag}_loss'])) <= 1e-5


def test_known_answer_rigid_copy():
  g = np.random.default_rng(3)
  X = g.normal(size=(500, 3)).astype(np.float32)
  a = 0.4
  Rg = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]], np.float32)
  tg = np.array([0.1, 0.2, -0.3], np.float32)
  R, t = oreg.weighted_procrustes(X, X @ Rg.T + tg, np.ones((500, 1), np.float32))
  np.testing.assert_allclose(R.numpy(), Rg, atol=1e-6)
  np.testing.assert_allclose(t.numpy(), tg, atol=1e-6)


def test_icp_oracle_known_answer():
  """oracle/icp.py (open3d RegistrationICP restatement): a slightly perturbed rigid copy converges
  back to the exact transform, with open3d's stop rule (both fitness and RMSE changes < 1e-6)."""
  from deepglobalregistration_b200 import synthetic as syn
  from oracle.icp import icp_point_to_point, kabsch
  g = np.random.default_rng(0)
  src = g.uniform(0, 2, size=(4000, 3))
  T_gt = syn.random_se3(g, 30.0, 0.4)
  tgt = syn.apply_se3(T_gt, src)
  R, t = kabsch(src, tgt)
  np.testing.assert_allclose(R, T_gt[:3, :3], atol=1e-10)
  np.testing.assert_allclose(t, T_gt[:3, 3], atol=1e-10)
  T_init = syn.random_se3(np.random.default_rng(1), 1.0, 0.01) @ T_gt
  T, info = icp_point_to_point(src, tgt, 0.1, T_init)
  te, re = syn.rte_rre(T, T_gt)
  assert te < 1e-9 and re < 1e-7 and info['fitness'] == 1.0 and info['iterations'] <= 30
  # no correspondence within the radius: the pose is left alone, fitness 0
  T2, info2 = icp_point_to_point(src, tgt + 100.0, 0.1, np.eye(4))
  assert np.array_equal(T2, np.eye(4)) and info2['fitness'] == 0.0


def test_lidar_pair_is_one_scene_seen_from_two_poses():
  from deepglobalregistration_b200 import synthetic as syn
  from scipy.spatial import cKDTree
  a, b, T = syn.lidar_pair(0)
  assert 100_000 < len(a) < 130_000 and abs(len(a) - len(b)) < 5000 and not np.array_equal(a[:100], b[:100])
  d, _ = cKDTree(b).query(syn.apply_se3(T, a[::50]))
  assert np.median(d) < 0.3          # static structure lines up under the ground-truth motion
