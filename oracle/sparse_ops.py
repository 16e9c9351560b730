"""ORACLE (test infrastructure only) - CPU restatement of the MinkowskiEngine 0.5.4
semantics the DGR hot path relies on.

PARITY UNPINNED for this file: MinkowskiEngine (reference requirements.txt:24) is
a third-party dependency that is neither vendored under /root/reference nor
installable offline, and the reference ships no tests or golden vectors
(SURVEY.md §4, §8c).  The semantics below are restated from MinkowskiEngine's
published behaviour and anchored on the reference's call sites:

  ME.utils.sparse_quantize        core/deep_global_registration.py:152
  ME.utils.batched_coordinates    core/deep_global_registration.py:158
  ME.SparseTensor                 core/deep_global_registration.py:167,214
  ME.MinkowskiConvolution         model/residual_block.py:38-44, model/resunet.py:589-596
  ME.MinkowskiConvolutionTranspose  model/residual_block.py:72-80
  ME.MinkowskiBatchNorm           model/common.py:13
  ME.cat / MEF.relu / +=          model/resunet.py:598-649, model/residual_block.py:118-134

They are cross-checked, independently of ME, against dense torch convolutions
(tests/test_oracle_sparse.py).  Frozen choices (SURVEY.md §8a list (1)-(9)):
  * kernel offset index kappa <-> offset vector: axis 0 fastest, centred, scaled by
    the input tensor stride;
  * stride-2 coordinates: per-axis floor(c / 2s) * 2s (floor toward -inf);
  * bucket kappa of a kernel map holds (in_row i, out_row j) with
    C_in[i] == C_out[j] + offset_kappa;
  * transposed convolution re-uses the matching down-convolution's buckets with the
    roles of i and j swapped and the same kappa -> weight index;
  * rows of a strided map are ordered by first occurrence among the finer rows
    (unobservable in ME; frozen here so intermediate tensors compare row by row).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product never does.
"""
import numpy as np
import torch

BN_EPS = 1e-5


# --------------------------------------------------------------------------- #
# integer coordinate work (exact)
# --------------------------------------------------------------------------- #
def _first_occurrence(rows):
  """Indices (ascending) of the first occurrence of every distinct row, and the
  inverse map row -> rank of its first occurrence."""
  rows = np.ascontiguousarray(rows)
  _, first, inv = np.unique(rows, axis=0, return_index=True, return_inverse=True)
  inv = np.asarray(inv).reshape(-1)
  order = np.argsort(first, kind='stable')          # unique-id -> rank by first occurrence
  rank = np.empty_like(order)
  rank[order] = np.arange(len(order))
  return first[order].astype(np.int64), rank[inv].astype(np.int64)


def quantize_first(xyz, voxel_size):
  """sparse_quantize(xyz / voxel, return_index=True) followed by the re-flooring of
  preprocess() (core/deep_global_registration.py:152-158).  The division happens in
  the input dtype (float64 for PLY/open3d inputs, float32 for float32 arrays).
  Returns (coords int32 [N,3] of the kept points, sel int64 [N] ascending)."""
  q = np.floor(xyz / voxel_size).astype(np.int32)
  sel, _ = _first_occurrence(q)
  return q[sel], sel


def batched_coordinates(coords_list):
  out = []
  for b, c in enumerate(coords_list):
    c = np.asarray(c, dtype=np.int32)
    out.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
  return np.concatenate(out, 0)


def stride_coords(coords, out_stride):
  """Coarse map of a stride-2 convolution: floor each spatial coordinate to a
  multiple of out_stride, keep the batch column, merge duplicates.
  Returns (coarse coords [M, D+1], inverse int64 [N] fine row -> coarse row)."""
  c = coords.astype(np.int64).copy()
  c[:, 1:] = np.floor_divide(c[:, 1:], out_stride) * out_stride
  sel, inv = _first_occurrence(c)
  return c[sel].astype(np.int32), inv


def kernel_offsets(kernel_size, D, tensor_stride=1):
  """[K, D] int offsets, kappa = sum_i digit_i * k**i (axis 0 fastest)."""
  k = int(kernel_size)
  K = k ** D
  kap = np.arange(K)
  offs = np.empty((K, D), np.int64)
  for ax in range(D):
    offs[:, ax] = ((kap // (k ** ax)) % k - k // 2) * tensor_stride
  return offs


def _linear_keys(coords, lo, span):
  c = coords.astype(np.int64) - lo
  key = np.zeros(len(c), np.int64)
  for ax in range(c.shape[1]):
    key = key * span[ax] + c[:, ax]
  return key


def kernel_map(in_coords, out_coords, offsets):
  """List over kappa of (in_rows, out_rows) int64 arrays, out_rows ascending.
  Bucket kappa: C_in[i] == C_out[j] + (0, offset_kappa)."""
  pad = int(np.abs(offsets).max()) + 1 if len(offsets) else 1
  both = np.concatenate([in_coords, out_coords], 0).astype(np.int64)
  lo = both.min(0) - pad
  span = both.max(0) + pad - lo + 1
  assert float(np.prod(span.astype(np.float64))) < 2.0 ** 62, 'coordinate range too large for oracle keys'
  kin = _linear_keys(in_coords, lo, span)
  order = np.argsort(kin, kind='stable')
  kin_sorted = kin[order]
  oc = out_coords.astype(np.int64)
  buckets = []
  for off in offsets:
    q = oc.copy()
    q[:, 1:] += off
    kq = _linear_keys(q, lo, span)
    pos = np.searchsorted(kin_sorted, kq)
    pos_c = np.minimum(pos, len(kin_sorted) - 1)
    hit = kin_sorted[pos_c] == kq if len(kin_sorted) else np.zeros(len(kq), bool)
    j = np.nonzero(hit)[0]
    buckets.append((order[pos_c[j]].astype(np.int64), j.astype(np.int64)))
  return buckets


def swap_map(buckets):
  return [(j, i) for (i, j) in buckets]


# --------------------------------------------------------------------------- #
# floating point layers
# --------------------------------------------------------------------------- #
def conv_forward(feat, weight, buckets, n_out, bias=None, dtype=torch.float32):
  """out[j] = sum_kappa sum_{(i,j) in bucket kappa} feat[i] @ W[kappa]  (+ bias),
  accumulated in ascending kappa (gather -> mm -> scatter-add per offset)."""
  feat = feat.to(dtype)
  w = weight.to(dtype)
  if w.dim() == 2:
    w = w[None]
  out = torch.zeros(n_out, w.shape[2], dtype=dtype)
  for kap, (i, j) in enumerate(buckets):
    if len(i) == 0:
      continue
    out.index_add_(0, torch.from_numpy(j), feat[torch.from_numpy(i)] @ w[kap])
  if bias is not None:
    out += bias.to(dtype).reshape(1, -1)
  return out


def linear_forward(feat, weight, bias=None, dtype=torch.float32):
  """kernel_size == 1 convolution: F @ W[Cin, Cout] (+ bias [1, Cout])."""
  out = feat.to(dtype) @ weight.to(dtype).reshape(feat.shape[1], -1)
  if bias is not None:
    out = out + bias.to(dtype).reshape(1, -1)
  return out


def batchnorm_eval(feat, sd, prefix, dtype=torch.float32, calibrate=False):
  if calibrate:      # record the input statistics as running statistics (util/calibrate.py twin)
    x = feat.float()
    sd[prefix + '.bn.running_mean'] = x.mean(0)
    sd[prefix + '.bn.running_var'] = x.var(0, unbiased=False).clamp_min(1e-6)
    sd[prefix + '.bn.weight'] = torch.ones(x.shape[1])
    sd[prefix + '.bn.bias'] = torch.zeros(x.shape[1])
  w = sd[prefix + '.bn.weight'].to(dtype)
  b = sd[prefix + '.bn.bias'].to(dtype)
  m = sd[prefix + '.bn.running_mean'].to(dtype)
  v = sd[prefix + '.bn.running_var'].to(dtype)
  return (feat.to(dtype) - m) / torch.sqrt(v + BN_EPS) * w + b


class CoordinateMaps:
  """Per-tensor-stride coordinate maps and kernel maps of one sparse tensor family
  (what ME's CoordinateManager caches)."""

  def __init__(self, coords):
    coords = np.asarray(coords, np.int32)
    assert len(np.unique(coords, axis=0)) == len(coords), 'oracle expects unique coordinates'
    self.D = coords.shape[1] - 1
    self.coords = {1: coords}
    self.down = {}      # stride s -> buckets of the (s -> 2s) convolution, kernel 3
    self.same = {}      # (stride, kernel_size) -> buckets

  def coords_at(self, s):
    if s not in self.coords:
      self.coords[s], _ = stride_coords(self.coords_at(s // 2), s)
    return self.coords[s]

  def same_map(self, s, kernel_size):
    key = (s, kernel_size)
    if key not in self.same:
      c = self.coords_at(s)
      self.same[key] = kernel_map(c, c, kernel_offsets(kernel_size, self.D, s))
    return self.same[key]

  def down_map(self, s, kernel_size=3):
    if s not in self.down:
      self.down[s] = kernel_map(self.coords_at(s), self.coords_at(2 * s),
                                kernel_offsets(kernel_size, self.D, s))
    return self.down[s]
