"""The oracle's ResUNetBN2C restatement (oracle/resunet.py) against the reference's OWN model code:
model/resunet.py + model/residual_block.py + model/common.py are imported unmodified from
/root/reference and run on the CPU over oracle/me_cpu.py (a MinkowskiEngine-shaped module backed by
oracle/sparse_ops.py).  Same sparse operators on both sides, so this isolates - and pins - the GRAPH:
layer order, strides, transposed-convolution pairing, skip concatenation order, norm placement, final
bias and normalisation.  Needs /root/reference (build container only): skipped elsewhere."""
import os
import sys

import numpy as np
import pytest
import torch

from deepglobalregistration_b200 import synthetic as syn
from oracle import me_cpu
from oracle.resunet import resunet_forward

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'model')), reason='reference tree not present')


@pytest.fixture
def ref_models():
  restore = me_cpu.install()
  saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'model' or k.startswith('model.')}
  sys.path.insert(0, REF)
  try:
    from model import load_model
    yield load_model
  finally:
    sys.path.remove(REF)
    for k in [k for k in sys.modules if k == 'model' or k.startswith('model.')]:
      del sys.modules[k]
    sys.modules.update(saved)
    restore()


def _cloud(seed, n, D, extent):
  g = np.random.default_rng(seed)
  c = np.unique(g.integers(-extent, extent, size=(n, D)), axis=0)
  return np.concatenate([np.zeros((len(c), 1), np.int64), c], 1).astype(np.int32)


@pytest.mark.parametrize('D,cin,cout,k1,normalize,n,extent', [
    (3, 1, 32, 7, True, 600, 7),        # FCGF, 3DMatch setting (scripts/train_3dmatch.sh:19)
    (3, 1, 32, 5, True, 500, 9),        # FCGF, KITTI setting
    (6, 1, 1, 3, False, 300, 2),        # inlier network, 'ones' features
    (6, 6, 1, 3, False, 250, 2),        # inlier network, 'coords' features
])
def test_reference_graph_equals_oracle_restatement(ref_models, D, cin, cout, k1, normalize, n, extent):
  sd = syn.resunet_state_dict(D + k1, cin, cout, k1, D)
  g = torch.Generator().manual_seed(1)
  for k in sd:                                  # non-trivial BN statistics so a misplaced norm shows
    if k.endswith('running_mean'):
      sd[k] = 0.1 * torch.randn(sd[k].shape, generator=g)
    if k.endswith('bn.bias'):
      sd[k] = 0.1 * torch.randn(sd[k].shape, generator=g)
  net = ref_models('ResUNetBN2C')(cin, cout, bn_momentum=0.05, conv1_kernel_size=k1, normalize_feature=normalize, D=D)
  missing = net.load_state_dict(sd, strict=True)
  net.eval()
  coords = _cloud(D, n, D, extent)
  feats = torch.ones(len(coords), cin) if cin == 1 else torch.randn(len(coords), cin, generator=g)
  import MinkowskiEngine as ME
  assert getattr(ME, '__oracle_stand_in__', False)
  with torch.no_grad():
    got = net(ME.SparseTensor(feats, coordinates=coords)).F
  want = resunet_forward(sd, coords, feats, k1, normalize)
  assert got.shape == want.shape == (len(coords), cout)
  err = float((got - want).abs().max() / (1 + want.abs().max()))
  assert err <= 1e-6, err
  assert float(want.abs().max()) > 1e-4           # the comparison is not vacuous
