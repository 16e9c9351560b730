"""Property tests (hypothesis) of the oracle's integer coordinate work - the invariants SURVEY.md
§8c lists for hash / unique / stride / kernel-map code, checked on random ragged inputs including
negative coordinates, duplicates, single points and empty neighbourhoods."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import sparse_ops as so

SET = settings(max_examples=60, deadline=None)


@st.composite
def clouds(draw, max_n=60, dims=(1, 2, 3, 6)):
  D = draw(st.sampled_from(dims))
  n = draw(st.integers(1, max_n))
  span = draw(st.integers(1, 6))
  rows = draw(st.lists(st.lists(st.integers(-span, span), min_size=D, max_size=D), min_size=n, max_size=n))
  return np.array(rows, dtype=np.int32).reshape(n, D)


def _unique_batched(c):
  u = np.unique(c, axis=0)
  return np.concatenate([np.zeros((len(u), 1), np.int32), u.astype(np.int32)], 1)


@SET
@given(clouds(dims=(3,)), st.sampled_from([0.05, 0.3, 1.0, 2.5]))
def test_quantize_first_invariants(c, voxel):
  xyz = c.astype(np.float64) * 0.37 + 0.011                 # arbitrary reals, negatives included
  q, sel = so.quantize_first(xyz, voxel)
  assert np.all(np.diff(sel) > 0)                            # ascending indices
  assert len(np.unique(q, axis=0)) == len(q)                 # one point per voxel
  full = np.floor(xyz / voxel).astype(np.int32)
  assert np.array_equal(q, full[sel])
  for row, v in zip(sel, q):                                 # and it is the FIRST point of its voxel
    assert row == np.flatnonzero((full == v).all(1))[0]
  assert len(q) == len(np.unique(full, axis=0))
  q2, sel2 = so.quantize_first(xyz[sel], voxel)              # idempotent on its own output
  assert np.array_equal(q2, q) and np.array_equal(sel2, np.arange(len(sel)))


@SET
@given(clouds(), st.sampled_from([2, 4, 8]))
def test_stride_coords_floor_and_cover(c, stride):
  coords = _unique_batched(c)
  coarse, inv = so.stride_coords(coords, stride)
  assert len(np.unique(coarse, axis=0)) == len(coarse)
  assert np.all(coarse[:, 1:] % stride == 0) and np.all(coarse[:, 0] == 0)
  parent = coarse[inv]                                       # every fine row has exactly its floor parent
  assert np.array_equal(parent[:, 1:], (coords[:, 1:] // stride) * stride)     # numpy // floors toward -inf
  assert set(inv.tolist()) == set(range(len(coarse)))        # no orphan coarse rows
  again, _ = so.stride_coords(coarse, stride)                # already on the lattice: unchanged as a set
  assert {tuple(r) for r in again} == {tuple(r) for r in coarse}


@SET
@given(clouds(dims=(1, 2, 3)), st.sampled_from([3, 5]))
def test_kernel_map_buckets(c, k):
  coords = _unique_batched(c)
  D = coords.shape[1] - 1
  offs = so.kernel_offsets(k, D, 1)
  assert offs.shape == (k ** D, D) and np.array_equal(offs[(k ** D) // 2], np.zeros(D, np.int32))
  buckets = so.kernel_map(coords, coords, offs)
  total = 0
  for kap, (i, j) in enumerate(buckets):
    assert len(np.unique(i)) == len(i) and len(np.unique(j)) == len(j)         # each row at most once per bucket
    assert np.array_equal(coords[i, 1:], coords[j, 1:] + offs[kap])           # C_in[i] = C_out[j] + offset
    assert np.all(np.diff(j) > 0) if len(j) > 1 else True                      # sorted by output row
    total += len(i)
  centre = buckets[(k ** D) // 2]
  assert np.array_equal(centre[0], np.arange(len(coords))) and np.array_equal(centre[1], centre[0])
  # offset -o holds the swapped pairs of offset o
  for kap in range(k ** D):
    i, j = buckets[kap]
    mi, mj = buckets[k ** D - 1 - kap]
    assert {(a, b) for a, b in zip(i, j)} == {(b, a) for a, b in zip(mi, mj)}
  # brute force count
  keys = {tuple(r) for r in coords[:, 1:]}
  want = sum(tuple(r + o) in keys for r in coords[:, 1:] for o in offs)
  assert total == want


@SET
@given(clouds(dims=(2, 3)), st.integers(1, 3), st.integers(1, 3))
def test_conv_is_linear_and_transpose_is_adjoint(c, cin, cout):
  coords = _unique_batched(c)
  coarse, _ = so.stride_coords(coords, 2)
  D = coords.shape[1] - 1
  down = so.kernel_map(coords, coarse, so.kernel_offsets(3, D, 1))
  g = torch.Generator().manual_seed(len(coords) * 7 + cin)
  W = torch.randn(3 ** D, cin, cout, generator=g, dtype=torch.float64)
  x, y = (torch.randn(len(coords), cin, generator=g, dtype=torch.float64) for _ in range(2))
  z = torch.randn(len(coarse), cout, generator=g, dtype=torch.float64)
  f = lambda v: so.conv_forward(v, W, down, len(coarse), dtype=torch.float64)
  assert torch.allclose(f(2.0 * x - 3.0 * y), 2.0 * f(x) - 3.0 * f(y), atol=1e-12)
  # <conv(x), z> == <x, conv_transpose(z)> with W^T per offset: the transposed map is the swapped one
  Wt = W.transpose(1, 2).contiguous()
  back = so.conv_forward(z, Wt, so.swap_map(down), len(coords), dtype=torch.float64)
  assert abs(float((f(x) * z).sum() - (x * back).sum())) <= 1e-9 * (1 + float(f(x).abs().sum()))
