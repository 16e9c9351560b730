"""Weighted Procrustes and robust SE(3) refinement with the reference's function
signatures (core/registration.py:91-113,135-194), executed by the single-launch cluster
kernel ``dgr_se3_register``."""
import numpy as np
import torch

from .. import _abi

F32_EPS = float(np.finfo(np.float32).eps)


def _prep(X, Y, w):
  if isinstance(X, np.ndarray):
    X = torch.from_numpy(X)
  if isinstance(Y, np.ndarray):
    Y = torch.from_numpy(Y)
  dev = X.device if X.is_cuda else torch.device('cuda')
  dev = _abi.require_device(dev)
  _abi.refresh_stream()
  X = X.to(dev, torch.float32).contiguous()
  Y = Y.to(dev, torch.float32).contiguous()
  if w is None:
    w = torch.ones(len(X), 1, device=dev)
  w = w.detach().to(dev, torch.float32).reshape(-1).contiguous()
  assert len(X) == len(Y) == len(w)
  return X, Y, w


def weighted_procrustes(X, Y, w, eps=F32_EPS):
  """-> (R [3,3], t [3]) float32 on the CPU, as the reference returns them."""
  X, Y, w = _prep(X, Y, w)
  res = _abi.se3_register(X, Y, w, max_iter=0).cpu()
  return res[:9].reshape(3, 3).clone(), res[9:12].clone()


def weighted_procrustes_autograd(X, Y, w, eps=F32_EPS):
  """Differentiable weighted Procrustes for training (core/registration.py:91-113 as called from
  core/trainer.py:580-600): the same arithmetic in torch ops on the inputs' device - fp32 moments, the 3x3 SVD in
  float64 with the det(U) det(V) reflection fix - so gradients reach the weights through torch's SVD backward.
  -> (R [3,3], t [3]) on the inputs' device.  Inference uses weighted_procrustes (the cluster kernel)."""
  X, Y = X.float(), Y.float()
  w = w.reshape(-1, 1).float()
  wn = w / (w.abs().sum() + eps)
  mx = (wn * X).sum(0, keepdim=True)
  my = (wn * Y).sum(0, keepdim=True)
  S = ((Y - my).t() @ (wn * (X - mx))).double()
  U, _, Vh = torch.linalg.svd(S)
  d = torch.ones(3, dtype=torch.float64, device=S.device)
  if float(torch.det(U) * torch.det(Vh)) < 0:
    d = d * torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64, device=S.device)
  R = (U @ torch.diag(d) @ Vh).float()
  t = my.reshape(3) - (R @ mx.reshape(3, 1)).reshape(3)
  return R, t


def GlobalRegistration(points, trans_points, weights=None, max_iter=1000, verbose=False, stat_freq=20,
                       max_break_count=20, break_threshold_ratio=1e-5, loss_fn=None, quantization_size=1):
  """-> (R [3,3], t [1,3], dict(iterations, loss, break_count)); tensors on the input device."""
  if loss_fn is not None:
    raise NotImplementedError('custom loss functions are not supported; the built path is '
                              'HighDimSmoothL1Loss (core/loss.py:42-61)')
  X, Y, w = _prep(points, trans_points, weights)
  res = _abi.se3_register(X, Y, w, quantization_size=quantization_size, max_iter=max_iter,
                          max_break_count=max_break_count, break_threshold_ratio=break_threshold_ratio)
  host = res.cpu()
  info = {'iterations': int(host[12]), 'loss': float(host[13]), 'break_count': int(host[14])}
  return res[:9].reshape(3, 3), res[9:12].reshape(1, 3), info
