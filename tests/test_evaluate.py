"""The pair-list evaluation driver (evaluate.py: what scripts/test_3dmatch.py / test_kitti.py do
around register()) on CPU: metric, 3DMatch tree and pair-list parsing, file-backed pairs through
the sharded loop at world size 1 and 2 (gloo), summary numbers."""
import datetime
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepglobalregistration_b200 import evaluate as ev
from deepglobalregistration_b200 import io as dio
from deepglobalregistration_b200 import synthetic as syn


class _CentroidRegistrar:
  """Stands in for DeepGlobalRegistration: translation between centroids (exact for the pure
  translations used here, wrong on purpose for the rotated pair)."""
  last_branch = 'procrustes'

  def register(self, a, b):
    a, b = np.asarray(getattr(a, 'points', a), np.float64), np.asarray(getattr(b, 'points', b), np.float64)
    T = np.eye(4)
    T[:3, 3] = b.mean(0) - a.mean(0)
    self.last_info = dict(wsum=float(len(a)), iterations=1)
    return T


def _tree(root):
  """Two scenes; fragment j = fragment i moved by a known pose; the last pair is rotated by 90 deg so the
  centroid registrar fails on it."""
  g = np.random.default_rng(0)
  expect = []
  for scene, n_pairs in (('kitchen', 3), ('lab', 2)):
    os.makedirs(os.path.join(root, scene))
    os.makedirs(os.path.join(root, scene + '-evaluation'))
    traj = []
    for k in range(n_pairs):
      p = g.normal(size=(200, 3))
      T01 = np.eye(4)
      T01[:3, 3] = g.normal(size=3)
      if scene == 'lab' and k == 1:
        T01[:3, :3] = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]
      dio.write_ply(os.path.join(root, scene, f'cloud_bin_{2 * k}.ply'), p, dtype='double')
      dio.write_ply(os.path.join(root, scene, f'cloud_bin_{2 * k + 1}.ply'), syn.apply_se3(T01, p), dtype='double')
      traj.append(([2 * k, 2 * k + 1, 2 * n_pairs], np.linalg.inv(T01)))     # gt.log: pose of j in i's frame
      expect.append(T01)
    dio.write_trajectory(os.path.join(root, scene + '-evaluation', 'gt.log'), traj)
  return expect


def test_metric_matches_the_reference_definition():
  T = syn.random_se3(np.random.default_rng(0), 10.0, 0.2)
  ok, rte, rre = ev.rte_rre(T, np.eye(4), 0.3, 15)
  te, re = syn.rte_rre(T, np.eye(4))
  assert abs(rte - te) < 1e-12 and abs(rre - math.degrees(re)) < 1e-9 and ok == 1
  assert ev.rte_rre(T, np.eye(4), 0.3, 1.0)[0] == 0
  assert ev.rte_rre(np.eye(4), np.eye(4), 0.3, 15)[2] < 1e-5          # the eps clip floors RRE just above 0
  assert np.array_equal(ev.rte_rre(None, np.eye(4), 0.3, 15), [0, np.inf, np.inf])


def test_threedmatch_tree_single_process(tmp_path):
  expect = _tree(str(tmp_path))
  pairs = ev.threedmatch_pairs(str(tmp_path))
  assert [p.group for p in pairs] == ['kitchen'] * 3 + ['lab'] * 2
  assert pairs[1].file0.endswith(os.path.join('kitchen', 'cloud_bin_2.ply'))
  for p, T in zip(pairs, expect):
    np.testing.assert_allclose(p.T_gt, T, atol=1e-12)
  res = ev.evaluate(_CentroidRegistrar(), pairs, 0.3, 15)
  assert res['stats'].shape == (5, 5) and list(res['stats'][:, 0]) == [1, 1, 1, 1, 0]
  assert list(res['stats'][:, 4]) == [0, 0, 0, 1, 1] and res['stats'][4, 2] > 80
  s = ev.summarize(res)
  assert s['recall'] == 0.8 and s['recall_per_group'] == {'kitchen': 1.0, 'lab': 0.5}
  assert s['recall_group_average'] == 0.75 and s['rre_success'] < 1e-3 and s['pairs'] == 5


def test_pair_list_formats(tmp_path):
  g = np.random.default_rng(1)
  a = g.normal(size=(50, 3)).astype(np.float32)
  np.concatenate([a, np.zeros((50, 1), np.float32)], 1).tofile(tmp_path / 'a.bin')
  np.savez(tmp_path / 'b.npz', pcd=a + np.float32(1.0))
  T = np.eye(4)
  T[:3, 3] = 1.0
  (tmp_path / 'pairs.txt').write_text(
      '# KITTI-style pair with ground truth and a drive id\n'
      f'a.bin b.npz {" ".join(repr(float(x)) for x in T.reshape(-1))} drive8\n'
      '\n'
      f'{tmp_path / "a.bin"} b.npz   # no ground truth\n')
  pairs = ev.read_pair_list(str(tmp_path / 'pairs.txt'))
  assert len(pairs) == 2 and pairs[0].group == 'drive8' and pairs[1].T_gt is None
  assert os.path.isabs(pairs[0].file0) and np.array_equal(pairs[0].T_gt, T)
  res = ev.evaluate(_CentroidRegistrar(), pairs, 0.6, 5)
  assert res['stats'][0, 0] == 1 and np.isnan(res['stats'][1, 0])
  np.testing.assert_allclose(res['poses'][1][:3, 3], 1.0, atol=1e-6)
  assert ev.summarize(res)['with_ground_truth'] == 1
  (tmp_path / 'bad.txt').write_text('a.bin b.npz 1 2 3\n')
  with pytest.raises(ValueError, match='expected'):
    ev.read_pair_list(str(tmp_path / 'bad.txt'))


def _worker(rank, world, port, root):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world, timeout=datetime.timedelta(seconds=90))
  res = ev.evaluate(_CentroidRegistrar(), ev.threedmatch_pairs(root), 0.3, 15)
  torch.save(res['stats'], os.path.join(root, f'stats{rank}.pt'))
  dist.destroy_process_group()


def test_two_ranks_agree_with_one(tmp_path):
  _tree(str(tmp_path))
  one = ev.evaluate(_CentroidRegistrar(), ev.threedmatch_pairs(str(tmp_path)), 0.3, 15)['stats']
  for attempt in range(3):                      # a lost rendezvous (port taken in between) is retried, not waited on
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    try:
      mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
      break
    except Exception:
      if attempt == 2:
        raise
  r0, r1 = torch.load(tmp_path / 'stats0.pt', weights_only=False), torch.load(tmp_path / 'stats1.pt', weights_only=False)
  assert np.array_equal(r0[:, [0, 4]], one[:, [0, 4]]) and np.array_equal(r0[:, :3], r1[:, :3])
  np.testing.assert_allclose(r0[:, 1:3], one[:, 1:3], atol=1e-9)


def test_metric_and_trajectory_reader_match_reference_golden():
  """tests/golden/reference_io.npz was produced by the reference's own util/file.py::read_trajectory
  and scripts/test_3dmatch.py::rte_rre (tests/golden/make_golden_io.py)."""
  here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
  gold = np.load(os.path.join(here, 'reference_io.npz'))
  traj = dio.read_trajectory(os.path.join(here, 'sample_gt.log'))
  assert np.array_equal(np.array([t.metadata for t in traj]), gold['traj_meta'])
  assert np.array_equal(np.stack([t.pose for t in traj]), gold['traj_pose'])           # bit-exact parse
  for Tp, Tg, want in zip(gold['metric_pred'], gold['metric_gt'], gold['metric_out']):
    got = ev.rte_rre(Tp, Tg, 0.3, 15)
    assert got[0] == want[0]
    np.testing.assert_allclose(got[1:], want[1:], rtol=1e-12, atol=1e-12)
  assert np.array_equal(ev.rte_rre(None, gold['metric_gt'][0], 0.3, 15), gold['metric_none'])
  assert 0 < gold['metric_out'][:, 0].sum() < len(gold['metric_out'])                  # both outcomes covered


def test_kitti_pairs_and_ground_truth_match_reference_golden(tmp_path):
  """tests/golden/reference_kitti.npz: file list and ground-truth poses produced by the reference's own
  KITTINMPairDataset on a synthetic odometry tree (tests/golden/make_golden_kitti.py)."""
  gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_kitti.npz'))
  os.makedirs(tmp_path / 'dataset' / 'poses')
  for drive in (8, 9, 10):
    P = gold[f'poses_{drive:02d}']
    np.savetxt(tmp_path / 'dataset' / 'poses' / f'{drive:02d}.txt', P, fmt='%.9e')
    vel = tmp_path / 'dataset' / 'sequences' / f'{drive:02d}' / 'velodyne'
    os.makedirs(vel)
    for k in range(len(P) - 1):                       # one pose more than scans (see the generator)
      (vel / f'{k:06d}.bin').write_bytes(b'')
  np.testing.assert_allclose(ev.KITTI_VELO2CAM, gold['velo2cam_T'].T, rtol=0, atol=0)
  pairs = ev.kitti_pairs(str(tmp_path))
  got = [(int(p.group[-2:]), int(os.path.basename(p.file0)[:-4]), int(os.path.basename(p.file1)[:-4])) for p in pairs]
  assert got == [tuple(int(x) for x in f) for f in gold['files']]
  for p, M in zip(pairs, gold['M']):
    np.testing.assert_allclose(p.T_gt, M, atol=1e-10)
    assert np.allclose(p.T_gt[:3, :3] @ p.T_gt[:3, :3].T, np.eye(3), atol=1e-6)
  # every pair is 10..~11.5 m apart (the frame BEFORE the first one farther than 10 m), same drive
  d = [np.linalg.norm(p.T_gt[:3, 3]) for p in pairs]
  assert 8.0 < min(d) and max(d) <= 10.0 + 1e-9
  assert {p.group for p in pairs} == {'drive08', 'drive09', 'drive10'}
  # a drive with no scans is an error, not an empty result
  os.makedirs(tmp_path / 'empty' / 'dataset' / 'sequences' / '08' / 'velodyne')
  with pytest.raises(FileNotFoundError):
    ev.kitti_pairs(str(tmp_path / 'empty'), drives=(8,))
