"""ORACLE (test infrastructure only) - CPU restatement of the non-MinkowskiEngine
stages of DGR's hot path.  PINNED: every function here is checked in
tests/test_oracle_registration.py against golden vectors produced by importing
the reference's own modules (tests/golden/make_golden.py).

  feature_knn            core/knn.py:23-74 (find_knn_gpu, chunked branch) with
                         core/metrics.py:62-65 (pdist 'L2')
  weighted_procrustes    core/registration.py:91-113
  rot6d_to_matrix        core/registration.py:16-64 (ortho2rotation)
  robust_loss            core/loss.py:42-61 (HighDimSmoothL1Loss)
  se3_refine             core/registration.py:135-194 (GlobalRegistration) with
                         core/registration.py:116-132 (Transformation)
  inlier_weights         core/deep_global_registration.py:269-272
"""
import numpy as np
import torch

F32_EPS = float(np.finfo(np.float32).eps)      # HighDimSmoothL1Loss.eps, core/loss.py:44


def feature_knn(F0, F1, nn_max_n=250, return_ambiguous=False, band=1e-6):
  """Top-1 neighbour in F1 of every row of F0: argmin_j sqrt(sum_c (a-b)^2 + 1e-7),
  lowest j on ties.  With return_ambiguous also returns a bool mask of rows whose
  float64 top-2 distance gap is below ``band`` (relative): rows where any fp32
  summation order may legitimately pick another index."""
  F0 = torch.as_tensor(F0, dtype=torch.float32)
  F1 = torch.as_tensor(F1, dtype=torch.float32)
  step = nn_max_n if nn_max_n > 1 else len(F0)
  idx, amb = [], []
  for s in range(0, len(F0), max(step, 1)):
    a = F0[s:s + step]
    diff = a[:, None, :] - F1[None, :, :]
    d2 = (diff * diff).sum(2)
    d = torch.sqrt(d2 + 1e-7) if nn_max_n > 1 else d2
    idx.append(d.min(dim=1).indices)
    if return_ambiguous:
      d64 = ((a.double()[:, None, :] - F1.double()[None, :, :]) ** 2).sum(2)
      k = min(2, d64.shape[1])
      top = torch.topk(d64, k, dim=1, largest=False).values
      if k == 2:
        gap = (top[:, 1] - top[:, 0])
        amb.append(gap <= band * (top[:, 1] + 1e-7))
      else:
        amb.append(torch.zeros(len(a), dtype=torch.bool))
  idx = torch.cat(idx) if idx else torch.zeros(0, dtype=torch.long)
  if return_ambiguous:
    return idx, (torch.cat(amb) if amb else torch.zeros(0, dtype=torch.bool))
  return idx


def inlier_weights(logit, clip_weight_thresh):
  w = torch.sigmoid(torch.as_tensor(logit, dtype=torch.float32))
  if clip_weight_thresh > 0:
    w = torch.where(w < clip_weight_thresh, torch.zeros_like(w), w)
  return w


def weighted_procrustes(X, Y, w, eps=F32_EPS):
  """Closed-form weighted Kabsch.  X, Y [N,3] fp32, w [N,1] fp32.  Moments in
  fp32, the 3x3 SVD in float64; returns (R [3,3], t [3]) fp32 with Y ~ R X + t."""
  X, Y, w = (torch.as_tensor(a, dtype=torch.float32) for a in (X, Y, w))
  w = w.reshape(-1, 1)
  wn = w / (w.abs().sum() + eps)
  mx = (wn * X).sum(0, keepdim=True)
  my = (wn * Y).sum(0, keepdim=True)
  S = ((Y - my).t() @ (wn * (X - mx))).double()
  U, _, Vh = torch.linalg.svd(S)
  d = torch.ones(3, dtype=torch.float64)
  if torch.det(U) * torch.det(Vh) < 0:
    d[2] = -1.0
  R = (U @ torch.diag(d) @ Vh).float()
  t = (my.reshape(3) - (R @ mx.reshape(3, 1)).reshape(3)).float()
  return R, t


def rot6d_to_matrix(p):
  """Gram-Schmidt on the two 3-vectors of p [6]; they become columns 0 and 1."""
  a, b = p[0:3], p[3:6]
  x = a / torch.clamp(torch.sqrt((a * a).sum()), min=1e-8)
  proj = (x * b).sum() / torch.clamp((x * x).sum(), min=1e-8) * x
  y = b - proj
  y = y / torch.clamp(torch.sqrt((y * y).sum()), min=1e-8)
  z = torch.linalg.cross(x, y)
  return torch.stack([x, y, z], dim=1)


def robust_loss(P, Y, w, q, eps=F32_EPS):
  s = (((P - Y) / q) ** 2).sum(1, keepdim=True)
  near = (s < 1).float()
  rho = 0.5 * near * s + 0.5 * (1 - near) * (torch.sqrt(s + eps) - 0.5)
  return (rho * w).sum() / w.sum()


def se3_refine(X, Y, w, quantization_size, max_iter=1000, max_break_count=20,
               break_threshold_ratio=1e-4, lr=0.1, gamma=0.999):
  """Weighted-Procrustes initialisation followed by Adam on (rot6d, trans) with an
  exponentially decayed learning rate and the reference's stopping rule: stop when
  loss < 1e-7, or after ``max_break_count`` (cumulative, never reset) iterations
  whose |loss_prev - loss| < loss_prev * ratio.  Returns R [3,3], t [1,3], info."""
  X, Y, w = (torch.as_tensor(a, dtype=torch.float32) for a in (X, Y, w))
  w = w.reshape(-1, 1).detach()
  R0, t0 = weighted_procrustes(X, Y, w)
  rot6d = torch.cat([R0[:, 0], R0[:, 1]]).clone().requires_grad_(True)
  trans = t0.reshape(1, 3).clone().requires_grad_(True)
  opt = torch.optim.Adam([rot6d, trans], lr=lr)

  def loss_fn():
    return robust_loss(X @ rot6d_to_matrix(rot6d).t() + trans, Y, w, quantization_size)

  with torch.no_grad():
    loss_prev = loss_fn().item()
  breaks, it, loss_val = 0, 0, loss_prev
  for it in range(max_iter):
    loss = loss_fn()
    loss_val = loss.item()
    if loss_val < 1e-7:
      break
    opt.zero_grad()
    loss.backward()
    opt.step()
    for g in opt.param_groups:
      g['lr'] *= gamma
    if abs(loss_prev - loss_val) < loss_prev * break_threshold_ratio:
      breaks += 1
      if breaks >= max_break_count:
        break
    loss_prev = loss_val
  with torch.no_grad():
    R = rot6d_to_matrix(rot6d.detach())
  return R, trans.detach(), dict(iterations=it, loss=loss_val, break_count=breaks)
