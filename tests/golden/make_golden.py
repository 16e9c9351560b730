"""Generate tests/golden/reference_stages.npz by importing the UNMODIFIED reference
modules from /root/reference (core/knn.py, core/registration.py, core/loss.py,
core/metrics.py).  /root/reference exists only in the build container, so the
vectors are committed and this script documents how they were made:

    python tests/golden/make_golden.py

MinkowskiEngine and open3d are not importable here, hence only the stages that do
not touch them can be pinned this way.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from core.knn import find_knn_gpu                       # noqa: E402
from core.loss import HighDimSmoothL1Loss               # noqa: E402
from core.registration import (GlobalRegistration, ortho2rotation,  # noqa: E402
                               weighted_procrustes)


def correspondences(seed, n, outlier_frac, angle=0.7, noise=0.0):
  """BASELINE.json configs[0] / SURVEY §8d config 1 generator."""
  g = np.random.default_rng(seed)
  X = g.normal(size=(n, 3)).astype(np.float32)
  R = np.array([[np.cos(angle), -np.sin(angle), 0], [np.sin(angle), np.cos(angle), 0], [0, 0, 1]],
               np.float32)
  t = np.array([0.3, -0.2, 0.1], np.float32)
  Y = X @ R.T + t + g.normal(scale=noise, size=(n, 3)).astype(np.float32)
  out = g.random(n) < outlier_frac
  Y[out] = g.normal(scale=2.0, size=(int(out.sum()), 3)).astype(np.float32)
  w = np.where(out, g.uniform(0, 0.1, n), g.uniform(0.5, 1.0, n)).astype(np.float32)[:, None]
  return X, Y.astype(np.float32), w, R, t


def main():
  torch.manual_seed(0)
  out = {}
  # --- kNN -------------------------------------------------------------------
  for tag, (n0, n1, c) in dict(a=(700, 900, 32), b=(257, 1, 32), c=(64, 300, 16)).items():
    g = np.random.default_rng(ord(tag))
    F0 = g.normal(size=(n0, c)).astype(np.float32)
    F1 = g.normal(size=(n1, c)).astype(np.float32)
    F0 /= np.linalg.norm(F0, axis=1, keepdims=True)
    F1 /= np.linalg.norm(F1, axis=1, keepdims=True)
    if n1 > 10:
      F1[7] = F1[3]                      # exact duplicate rows: lowest index must win
      F0[5] = F1[3]
    nn = find_knn_gpu(torch.from_numpy(F0), torch.from_numpy(F1), nn_max_n=250, knn=1)
    out[f'knn_{tag}_F0'], out[f'knn_{tag}_F1'] = F0, F1
    out[f'knn_{tag}_idx'] = nn.long().reshape(-1).numpy()
  # --- Procrustes / loss / rot6d / refine ------------------------------------------
  for tag, (seed, n, frac, noise) in dict(a=(0, 1000, 0.3, 0.0), b=(1, 300, 0.0, 0.01),
                                           c=(2, 5000, 0.5, 0.02)).items():
    X, Y, w, R, t = correspondences(seed, n, frac, noise=noise)
    Xt, Yt, wt = torch.from_numpy(X), torch.from_numpy(Y), torch.from_numpy(w)
    Rp, tp = weighted_procrustes(Xt, Yt, wt, np.finfo(np.float32).eps)
    loss = HighDimSmoothL1Loss(wt, 0.1)
    l0 = loss(Xt @ Rp.t() + tp, Yt).item()
    Rr, tr, info = GlobalRegistration(Xt, Yt, weights=wt.clone(), break_threshold_ratio=1e-4,
                                      quantization_size=0.1, verbose=False)
    out.update({f'reg_{tag}_X': X, f'reg_{tag}_Y': Y, f'reg_{tag}_w': w,
                f'reg_{tag}_R_gt': R, f'reg_{tag}_t_gt': t,
                f'reg_{tag}_R_proc': Rp.numpy(), f'reg_{tag}_t_proc': tp.numpy(),
                f'reg_{tag}_loss_proc': np.float64(l0),
                f'reg_{tag}_R_ref': Rr.numpy(), f'reg_{tag}_t_ref': tr.numpy(),
                f'reg_{tag}_iters': np.int64(info['iterations']),
                f'reg_{tag}_loss': np.float64(info['loss']),
                f'reg_{tag}_breaks': np.int64(info['break_count'])})
  p = torch.tensor([[0.9, 0.1, -0.2, 0.3, 1.1, 0.05], [1e-9, 0, 0, 0, 1e-9, 0]])
  out['rot6d_in'] = p.numpy()
  out['rot6d_out'] = ortho2rotation(p).numpy()
  np.savez_compressed(os.path.join(HERE, 'reference_stages.npz'), **out)
  print('wrote', len(out), 'arrays')


if __name__ == '__main__':
  main()
