"""ORACLE (test infrastructure only) - the whole pairwise-registration hot path on
the CPU, stage by stage, following DeepGlobalRegistration.register()
(core/deep_global_registration.py:238-324): Procrustes + SE(3) refinement, or
the RANSAC safeguard (:302-315, oracle/ransac.py) when the weight sum is below
the gate, then optionally the ICP fine-tune (:317-322, oracle/icp.py).

PINNED against the reference's own register() run on the CPU over oracle-backed stand-ins
for MinkowskiEngine / open3d (tests/test_oracle_pipeline_vs_reference.py); the sparse operators
and ICP underneath remain unpinned restatements (see their headers).

Each stage returns its tensors so that GPU parity tests can tap in anywhere and
feed the oracle's outputs of stage k into the CUDA stage k+1 (stage-isolated
parity), see tests/test_gpu_pipeline.py.
"""
import numpy as np
import torch

from . import sparse_ops as so
from .registration import feature_knn, inlier_weights, se3_refine
from .resunet import resunet_forward


def preprocess(xyz, voxel_size):
  """core/deep_global_registration.py:134-161 -> (xyz fp32 [N,3], coords int32
  [N,4] with batch column 0, sel)."""
  coords, sel = so.quantize_first(xyz, voxel_size)
  return xyz[sel].astype(np.float32), so.batched_coordinates([coords]), sel


def fcgf(state, coords):
  cfg = state['config']
  feats = torch.ones(len(coords), 1)
  return resunet_forward(state['state_dict'], coords, feats, cfg['feat_conv1_kernel_size'],
                         cfg['normalize_feature'])


def inlier_coords(coords0, coords1, idx1):
  """:261-262  cat(coords0[idx0], coords1[idx1, 1:]) with idx0 = arange."""
  return np.concatenate([coords0, coords1[idx1, 1:]], 1).astype(np.int32)


def inlier_features(feat_type, xyz0, xyz1, idx1):
  """:185-208"""
  if feat_type == 'ones':
    return torch.ones(len(idx1), 1)
  if feat_type == 'coords':
    return torch.cat([torch.cos(torch.from_numpy(xyz0)), torch.cos(torch.from_numpy(xyz1[idx1]))], 1)
  raise ValueError(feat_type)


def inlier_logits(state, coords6, feats):
  cfg = state['config']
  return resunet_forward(state['state_dict_inlier'], coords6, feats, cfg['inlier_conv1_kernel_size'],
                         False)


def register(state, xyz0, xyz1, clip_weight_thresh=0.05, use_icp=False, safeguard_max_iteration=0,
             safeguard_seed=0):
  """Returns (T 4x4 float64, taps dict); taps['branch'] says which branch the weight-sum gate
  (:276-281) took.  The safeguard branch runs safeguard_max_iteration RANSAC hypotheses (the
  reference: 4 M; 0 = skip and return identity, for tests that only look at the gate)."""
  cfg = state['config']
  vs = cfg['voxel_size']
  p0, c0, sel0 = preprocess(xyz0, vs)
  p1, c1, sel1 = preprocess(xyz1, vs)
  f0, f1 = fcgf(state, c0), fcgf(state, c1)
  idx1 = feature_knn(f0, f1, cfg['nn_max_n']).numpy()
  c6 = inlier_coords(c0, c1, idx1)
  logit = inlier_logits(state, c6, inlier_features(cfg['inlier_feature_type'], p0, p1, idx1))
  w = inlier_weights(logit, clip_weight_thresh)
  wsum = float(w.sum())
  taps = dict(xyz0=p0, xyz1=p1, coords0=c0, coords1=c1, sel0=sel0, sel1=sel1, feat0=f0, feat1=f1,
              idx1=idx1, coords6=c6, logit=logit, weights=w, wsum=wsum)
  T = np.eye(4)
  if wsum >= max(200, len(w) * 0.05):
    R, t, info = se3_refine(p0, p1[idx1], w, 2 * vs)
    T[:3, :3] = R.numpy()
    T[:3, 3] = t.numpy().reshape(3)
    taps.update(branch='procrustes', refine=info, T_refined=T.copy())
  else:
    taps.update(branch='safeguard')
    if safeguard_max_iteration > 0:      # :302-315
      from .ransac import ransac_correspondence
      T, info = ransac_correspondence(p0, p1, np.arange(len(idx1)), idx1, 2 * vs, safeguard_max_iteration,
                                      safeguard_seed)
      taps.update(ransac=info, T_ransac=T.copy())
  if use_icp and (taps['branch'] == 'procrustes' or safeguard_max_iteration > 0):      # :317-322
    from .icp import icp_point_to_point
    T, icp_info = icp_point_to_point(p0, p1, 2 * vs, T)
    taps.update(icp=icp_info)
  return T, taps
