"""ORACLE (test infrastructure only) - CPU forward pass of the reference's ResUNet2
family (ResUNetBN2C) straight from a state dict.

Follows the graph of model/resunet.py:598-649 (ResUNet2.forward), the block of
model/residual_block.py:118-134 (BasicBlockBase.forward: conv-BN-ReLU-conv-BN,
add residual, ReLU) and the eval-mode norm of model/common.py:11-13.  Layer
hyper-parameters are those of ResUNet2.__init__ (model/resunet.py:428-596):
all 3^D convolutions except conv1 (conv1_kernel_size) and the two 1x1 layers;
``conv()`` never passes a bias (model/residual_block.py:38-44), only ``final``
has one (model/resunet.py:589-596).

The GRAPH is pinned: tests/test_oracle_graph_vs_reference.py runs the reference's unmodified
model/resunet.py on the CPU over oracle/me_cpu.py and requires identical outputs.
PARITY UNPINNED for the sparse operators themselves, see oracle/sparse_ops.py.
"""
import torch

from . import sparse_ops as so


def _block(x, sd, name, buckets, n, dtype, cal=False):
  out = so.conv_forward(x, sd[f'{name}.conv1.kernel'], buckets, n, dtype=dtype)
  out = torch.relu(so.batchnorm_eval(out, sd, f'{name}.norm1', dtype, cal))
  out = so.conv_forward(out, sd[f'{name}.conv2.kernel'], buckets, n, dtype=dtype)
  out = so.batchnorm_eval(out, sd, f'{name}.norm2', dtype, cal)
  return torch.relu(out + x)


def resunet_forward(sd, coords, feats, conv1_kernel_size, normalize_feature, dtype=torch.float32,
                    taps=None, calibrate=False):
  """coords int32 [N, D+1] (unique), feats [N, Cin] -> [N, Cout].
  ``taps`` (optional dict) receives named intermediate tensors.  ``calibrate`` rewrites the
  BatchNorm statistics in ``sd`` from the activations of this pass (synthetic checkpoints)."""
  cal = calibrate
  maps = so.CoordinateMaps(coords)
  n = {s: len(maps.coords_at(s)) for s in (1, 2, 4, 8)}
  tap = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)
  x = feats.to(dtype)

  out = so.conv_forward(x, sd['conv1.kernel'], maps.same_map(1, conv1_kernel_size), n[1], dtype=dtype)
  tap('conv1', out)
  out = so.batchnorm_eval(out, sd, 'norm1', dtype, cal)
  out_s1 = _block(out, sd, 'block1', maps.same_map(1, 3), n[1], dtype, cal)
  tap('out_s1', out_s1)
  skips = {1: out_s1}
  out = torch.relu(out_s1)
  for lvl, s in ((2, 1), (3, 2), (4, 4)):
    out = so.conv_forward(out, sd[f'conv{lvl}.kernel'], maps.down_map(s), n[2 * s], dtype=dtype)
    out = so.batchnorm_eval(out, sd, f'norm{lvl}', dtype, cal)
    out = _block(out, sd, f'block{lvl}', maps.same_map(2 * s, 3), n[2 * s], dtype, cal)
    tap(f'out_s{2 * s}', out)
    skips[2 * s] = out
    out = torch.relu(out)
  for lvl, s in ((4, 4), (3, 2), (2, 1)):   # transposed convs: stride 2s -> s
    out = so.conv_forward(out, sd[f'conv{lvl}_tr.kernel'], so.swap_map(maps.down_map(s)), n[s],
                          dtype=dtype)
    out = so.batchnorm_eval(out, sd, f'norm{lvl}_tr', dtype, cal)
    out = torch.relu(_block(out, sd, f'block{lvl}_tr', maps.same_map(s, 3), n[s], dtype, cal))
    tap(f'out_s{s}_tr', out)
    out = torch.cat([out, skips[s]], 1)          # ME.cat(upsampled, skip)
  out = torch.relu(so.linear_forward(out, sd['conv1_tr.kernel'], dtype=dtype))
  out = so.linear_forward(out, sd['final.kernel'], sd['final.bias'], dtype=dtype)
  tap('final', out)
  if normalize_feature:
    out = out / (torch.norm(out, p=2, dim=1, keepdim=True) + 1e-8)
  return out
