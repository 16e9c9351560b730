"""World-size-2 test of the pair-sharding host logic on CPU (gloo): ownership, result
packing and the all-gather return every pair's result, in order, on every rank."""
import datetime
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
  dist.init_process_group('gloo', rank=rank, world_size=world, timeout=datetime.timedelta(seconds=90))
  res = sharding.register_pairs(_FakeDgr(), _pairs(n_pairs))
  torch.save(res, os.path.join(out_dir, f'r{rank}.pt'))
  dist.destroy_process_group()


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _spawn(fn, world, *args):
  """mp.spawn on a fresh port; a lost rendezvous (the port was taken between _free_port() and the bind) times out
  in the workers and is retried instead of hanging the suite."""
  for attempt in range(3):
    try:
      mp.spawn(fn, args=(world, _free_port(), *args), nprocs=world, join=True)
      return
    except Exception:
      if attempt == 2:
        raise


def test_shard_indices_cover_everything_once():
  for n, w in ((0, 2), (1, 2), (7, 2), (256, 8), (5, 8)):
    got = sorted(i for r in range(w) for i in sharding.shard_indices(n, r, w))
    assert got == list(range(n))


def test_single_process_path():
  res = sharding.register_pairs(_FakeDgr(), _pairs(3))
  assert res.shape == (3, 20) and res[2, 3].item() == 2 + 2 * 12


def test_two_rank_gloo_all_gather(tmp_path):
  n_pairs, world = 7, 2
  _spawn(_worker, world, n_pairs, str(tmp_path))
  r0, r1 = torch.load(tmp_path / 'r0.pt'), torch.load(tmp_path / 'r1.pt')
  assert torch.equal(r0, r1) and r0.shape == (n_pairs, 20)
  for i in range(n_pairs):
    assert r0[i, 3].item() == i + 2 * (10 + i)           # pose translation of pair i
    assert r0[i, 16].item() == 9.0 * i                    # wsum
    assert r0[i, 17].item() == 10 + i                     # iterations
    assert r0[i, 18].item() == (0.0 if i % 2 == 0 else 1.0)


class _FakeBatchDgr(_FakeDgr):
  """A registrar with the round-2 batch API: register_pairs hands it lazily produced pairs (callables) and keeps
  the results in input order whatever order the in-flight pairs complete in."""

  def register_batch(self, pairs, inflight=4):
    import threading
    import time
    out = [None] * len(pairs)

    def work(k):
      for i in range(k, len(pairs), inflight):
        a, b = pairs[i]() if callable(pairs[i]) else pairs[i]
        time.sleep(0.001 * ((7 * i) % 5))                  # completion order differs from input order
        d = _FakeDgr()
        T = d.register(a, b)
        out[i] = (T, d.last_branch, dict(d.last_info, t_done=time.perf_counter()))
    threads = [threading.Thread(target=work, args=(k,)) for k in range(inflight)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    return out


def _worker_batch(rank, world, port, n_pairs, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world, timeout=datetime.timedelta(seconds=90))
  res = sharding.register_pairs(_FakeBatchDgr(), _pairs(n_pairs), inflight=3)
  torch.save(res, os.path.join(out_dir, f'b{rank}.pt'))
  dist.destroy_process_group()


def test_batch_api_two_ranks_keep_pair_order(tmp_path):
  n_pairs, world = 11, 2
  _spawn(_worker_batch, world, n_pairs, str(tmp_path))
  r0, r1 = torch.load(tmp_path / 'b0.pt'), torch.load(tmp_path / 'b1.pt')
  assert torch.equal(r0[:, :19], r1[:, :19]) and r0.shape == (n_pairs, 20)
  for i in range(n_pairs):
    assert r0[i, 3].item() == i + 2 * (10 + i) and r0[i, 16].item() == 9.0 * i and r0[i, 17].item() == 10 + i
  assert bool((r0[:, 19] >= 0).all())                      # milliseconds between completions on the owning rank
