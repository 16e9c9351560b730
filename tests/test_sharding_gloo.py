"""World-size-2 test of the pair-sharding host logic on CPU (gloo): ownership, result
packing and the all-gather return every pair's result, in order, on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepglobalregistration_b200 import sharding


class _FakeDgr:
  """Stands in for DeepGlobalRegistration: the collective plumbing is what is under test."""

  def register(self, a, b):
    T = np.eye(4)
    T[:3, 3] = a[0] + 2 * b[0]
    self.last_info = dict(wsum=float(a.sum()), iterations=int(b[0, 0]))
    self.last_branch = 'procrustes' if a[0, 0] % 2 == 0 else 'safeguard'
    return T


def _pairs(n):
  return [(np.full((3, 3), float(i)), np.full((3, 3), float(10 + i))) for i in range(n)]


def _worker(rank, world, port, n_pairs, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  res = sharding.register_pairs(_FakeDgr(), _pairs(n_pairs))
  torch.save(res, os.path.join(out_dir, f'r{rank}.pt'))
  dist.destroy_process_group()


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def test_shard_indices_cover_everything_once():
  for n, w in ((0, 2), (1, 2), (7, 2), (256, 8), (5, 8)):
    got = sorted(i for r in range(w) for i in sharding.shard_indices(n, r, w))
    assert got == list(range(n))


def test_single_process_path():
  res = sharding.register_pairs(_FakeDgr(), _pairs(3))
  assert res.shape == (3, 20) and res[2, 3].item() == 2 + 2 * 12


def test_two_rank_gloo_all_gather(tmp_path):
  n_pairs, world = 7, 2
  mp.spawn(_worker, args=(world, _free_port(), n_pairs, str(tmp_path)), nprocs=world, join=True)
  r0, r1 = torch.load(tmp_path / 'r0.pt'), torch.load(tmp_path / 'r1.pt')
  assert torch.equal(r0, r1) and r0.shape == (n_pairs, 20)
  for i in range(n_pairs):
    assert r0[i, 3].item() == i + 2 * (10 + i)           # pose translation of pair i
    assert r0[i, 16].item() == 9.0 * i                    # wsum
    assert r0[i, 17].item() == 10 + i                     # iterations
    assert r0[i, 18].item() == (0.0 if i % 2 == 0 else 1.0)
