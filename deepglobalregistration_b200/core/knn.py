"""Feature nearest neighbour with the reference's ``find_knn_gpu`` signature and return
shapes (core/knn.py:23-74): [N, 1] int64 in the chunked branch (nn_max_n > 1), [N]
otherwise.  ``nn_max_n`` only selects the compared quantity (sqrt(d2 + 1e-7) vs d2) - the
CUDA kernel never materialises a distance matrix, so no chunking is needed."""
import torch

from .. import _abi


def find_knn_gpu(F0, F1, nn_max_n=-1, knn=1, return_distance=False):
  if knn != 1:
    raise NotImplementedError('the DGR inference path uses knn=1 (core/deep_global_registration.py:178)')
  dev = _abi.require_device(F0.device)
  _abi.refresh_stream()
  F0 = F0.to(dev, torch.float32).contiguous()
  F1 = F1.to(dev, torch.float32).contiguous()
  idx, dist = _abi.knn_top1(F0, F1, return_distance=True)
  inds = idx.long()
  if nn_max_n > 1:
    inds, dists = inds.unsqueeze(1), dist.unsqueeze(1)
  else:
    # the un-chunked branch of the reference compares squared distances
    dists = (dist * dist - 1e-7).clamp_min(0).unsqueeze(1)
  return (inds, dists) if return_distance else inds
