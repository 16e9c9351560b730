"""Self-consistency of the MinkowskiEngine restatement (oracle/sparse_ops.py): exact
integer semantics on hand-checkable cases, and equivalence of the sparse
convolution / transposed convolution with dense torch convolutions evaluated at
the active sites - a check of offset order and transpose semantics that does not
depend on MinkowskiEngine (SURVEY.md §8c golden-vector list (i))."""
import numpy as np
import torch
import torch.nn.functional as tF

from oracle import sparse_ops as so


def test_quantize_first_keeps_first_point_per_voxel_ascending():
  xyz = np.array([[0.26, 0.0, 0.0], [0.01, 0.0, 0.0], [0.29, 0.04, 0.0], [-0.01, 0.0, 0.0],
                  [0.02, 0.01, 0.03]])
  coords, sel = so.quantize_first(xyz, 0.05)
  assert sel.tolist() == [0, 1, 3]
  assert coords.tolist() == [[5, 0, 0], [0, 0, 0], [-1, 0, 0]]      # floor, not truncation


def test_quantize_divides_in_input_dtype():
  v = 0.05
  x64 = np.array([[0.15, 0.3, 3 * 0.05]], np.float64)
  c64, _ = so.quantize_first(x64, v)
  assert c64.tolist() == np.floor(x64 / v).astype(np.int32).tolist()
  x32 = x64.astype(np.float32)
  c32, _ = so.quantize_first(x32, v)
  assert c32.tolist() == np.floor(x32 / v).astype(np.int32).tolist()
  assert (x32 / v).dtype == np.float32


def test_stride_coords_floor_for_negatives():
  c = np.array([[0, -1, 0, 3], [0, -2, 1, 2], [0, 1, -3, -4], [0, 0, 0, 0]], np.int32)
  coarse, inv = so.stride_coords(c, 2)
  assert coarse.tolist() == [[0, -2, 0, 2], [0, 0, -4, -4], [0, 0, 0, 0]]
  assert inv.tolist() == [0, 0, 1, 2]
  coarse4, _ = so.stride_coords(coarse, 4)
  assert coarse4.tolist() == [[0, -4, 0, 0], [0, 0, -4, -4], [0, 0, 0, 0]]


def test_kernel_offsets_axis0_fastest():
  o = so.kernel_offsets(3, 3, 2)
  assert o[0].tolist() == [-2, -2, -2]
  assert o[1].tolist() == [0, -2, -2]
  assert o[3].tolist() == [-2, 0, -2]
  assert o[13].tolist() == [0, 0, 0]
  assert o[26].tolist() == [2, 2, 2]
  assert so.kernel_offsets(3, 6).shape == (729, 6)
  assert so.kernel_offsets(7, 3)[171].tolist() == [0, 0, 0]


def _random_cloud(rng, n, lo, hi, D=3):
  c = np.unique(rng.integers(lo, hi, size=(n, D)), axis=0)
  c = c[rng.permutation(len(c))]
  return np.concatenate([np.zeros((len(c), 1), np.int64), c], 1).astype(np.int32)


def _dense(coords, feat, lo, size):
  """[1, C, Z, Y, X] grid: axis 0 (x) of the sparse coords is the LAST dense axis."""
  C = feat.shape[1]
  g = torch.zeros(1, C, size, size, size, dtype=feat.dtype)
  x, y, z = (coords[:, 1 + a].astype(np.int64) - lo for a in range(3))
  g[0, :, z, y, x] = feat.t()
  return g


def _conv3d_weight(W, k):
  """W [K, Cin, Cout] with kappa = dx + k*dy + k*k*dz  ->  conv3d weight [Cout, Cin, kz, ky, kx]."""
  return W.reshape(k, k, k, W.shape[1], W.shape[2]).permute(4, 3, 0, 1, 2).contiguous()


def test_conv_stride1_equals_dense_conv3d():
  rng = np.random.default_rng(0)
  for k in (3, 5):
    coords = _random_cloud(rng, 60, -4, 5)
    n = len(coords)
    feat = torch.randn(n, 4, dtype=torch.float64)
    W = torch.randn(k ** 3, 4, 6, dtype=torch.float64)
    buckets = so.kernel_map(coords, coords, so.kernel_offsets(k, 3, 1))
    got = so.conv_forward(feat, W, buckets, n, dtype=torch.float64)
    lo, size = -4 - k, 9 + 2 * k
    dense = tF.conv3d(_dense(coords, feat, lo, size), _conv3d_weight(W, k), padding=k // 2)
    x, y, z = (coords[:, 1 + a].astype(np.int64) - lo for a in range(3))
    want = dense[0][:, z, y, x].t()
    torch.testing.assert_close(got, want, atol=1e-10, rtol=0)


def test_conv_stride2_and_transpose_equal_dense():
  rng = np.random.default_rng(1)
  fine = _random_cloud(rng, 80, -6, 6)
  coarse, _ = so.stride_coords(fine, 2)
  nf, nc = len(fine), len(coarse)
  down = so.kernel_map(fine, coarse, so.kernel_offsets(3, 3, 1))
  # --- strided convolution ------------------------------------------------------
  feat = torch.randn(nf, 3, dtype=torch.float64)
  W = torch.randn(27, 3, 5, dtype=torch.float64)
  got = so.conv_forward(feat, W, down, nc, dtype=torch.float64)
  lo, size = -8, 16                      # even lo keeps the stride-2 lattice aligned
  dense = tF.conv3d(_dense(fine, feat, lo, size), _conv3d_weight(W, 3), padding=1, stride=1)
  x, y, z = (coarse[:, 1 + a].astype(np.int64) - lo for a in range(3))
  torch.testing.assert_close(got, dense[0][:, z, y, x].t(), atol=1e-10, rtol=0)
  # --- transposed convolution: same buckets, roles swapped, same kappa ----------
  cfeat = torch.randn(nc, 5, dtype=torch.float64)
  Wt = torch.randn(27, 5, 2, dtype=torch.float64)
  got_t = so.conv_forward(cfeat, Wt, so.swap_map(down), nf, dtype=torch.float64)
  # dense restatement: out[c_coarse - off_k] += in[c_coarse] @ Wt[k] on the full grid
  grid = torch.zeros(size, size, size, 2, dtype=torch.float64)
  offs = so.kernel_offsets(3, 3, 1)
  for r in range(nc):
    for kap in range(27):
      p = coarse[r, 1:].astype(np.int64) + offs[kap] - lo
      grid[p[2], p[1], p[0]] += cfeat[r] @ Wt[kap]
  x, y, z = (fine[:, 1 + a].astype(np.int64) - lo for a in range(3))
  torch.testing.assert_close(got_t, grid[z, y, x], atol=1e-10, rtol=0)


def test_transpose_is_adjoint_of_conv():
  rng = np.random.default_rng(2)
  fine = _random_cloud(rng, 50, -5, 5)
  coarse, _ = so.stride_coords(fine, 2)
  down = so.kernel_map(fine, coarse, so.kernel_offsets(3, 3, 1))
  W = torch.randn(27, 3, 4, dtype=torch.float64)
  x = torch.randn(len(fine), 3, dtype=torch.float64)
  y = torch.randn(len(coarse), 4, dtype=torch.float64)
  lhs = (so.conv_forward(x, W, down, len(coarse), dtype=torch.float64) * y).sum()
  rhs = (x * so.conv_forward(y, W.transpose(1, 2), so.swap_map(down), len(fine), dtype=torch.float64)).sum()
  assert abs(lhs - rhs) < 1e-9


def test_kernel_map_6d_buckets():
  rng = np.random.default_rng(3)
  c = _random_cloud(rng, 40, -2, 3, D=6)
  offs = so.kernel_offsets(3, 6, 1)
  buckets = so.kernel_map(c, c, offs)
  assert len(buckets) == 729
  centre = 364
  assert buckets[centre][0].tolist() == list(range(len(c)))
  total = 0
  for kap, (i, j) in enumerate(buckets):
    assert np.array_equal(c[i][:, 1:], c[j][:, 1:] + offs[kap])
    assert len(np.unique(j)) == len(j) and len(np.unique(i)) == len(i)
    total += len(i)
  # brute force pair count
  d = c[:, None, 1:].astype(np.int64) - c[None, :, 1:]
  assert total == int((np.abs(d).max(2) <= 1).sum())
