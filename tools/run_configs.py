"""BASELINE.json configs 3 and 5 on one B200 (development record; bench.py is the contract).

  config 3  KITTI-shape pair (64-beam synthetic LiDAR, ~120k returns -> ~18k voxels at 0.3 m,
            FCGF conv1 k=5): full register(), per-stage CUDA-event times.
  config 5  stress: room-type cloud scaled to ~1M voxels at 0.02 m, FCGF out = 64: voxelisation,
            hash / kernel-map construction, FCGF forward and feature kNN, sweep over N.

Round 2: config 3 goes through the native executor (one C call per pair); config 5 additionally runs the FCGF
forward through dgr_net_forward (device-side counts, bit-mask kernel maps) at every size and compares it with the
operator path.  Writes gpurun_out/configs_r02.json."""
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepglobalregistration_b200 import _abi, me as ME, native, synthetic as syn
from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration
from deepglobalregistration_b200.model import load_model


def ev():
  e = torch.cuda.Event(enable_timing=True)
  e.record()
  return e


def config3():
  st = syn.make_checkpoint(3, voxel_size=0.3, feat_conv1_kernel_size=5)
  d = DeepGlobalRegistration(types.SimpleNamespace(weights=st, clip_weight_thresh=0.05, verbose=False))
  d.use_icp = False
  xyz0, xyz1, T = syn.lidar_pair(0)
  xyz0, xyz1 = xyz0.astype(np.float32), xyz1.astype(np.float32)      # KITTI .bin is float32
  for _ in range(5):
    d.register(xyz0, xyz1)
  torch.cuda.synchronize()
  t = time.perf_counter()
  n = 20
  for _ in range(n):
    Tm = d.register(xyz0, xyz1)
  torch.cuda.synchronize()
  ms = (time.perf_counter() - t) / n * 1e3
  return dict(n_raw=[len(xyz0), len(xyz1)], n0=d.last_info['n0'], n1=d.last_info['n1'], branch=d.last_branch,
              refine_iterations=d.last_info.get('iterations'), ms_per_pair=ms, pairs_per_s=1e3 / ms)


def config5():
  out = []
  model = load_model('ResUNetBN2C')(1, 64, bn_momentum=0.05, conv1_kernel_size=7, normalize_feature=True)
  model.load_state_dict(syn.resunet_state_dict(0, 1, 64, 7, 3))
  model = model.cuda().eval()
  net, ctx = native.Net(model, 'cuda'), native.Context('cuda')
  full = syn.room_scan(0, n_raw=3_000_000, extent=(9.0, 7.5, 3.0))
  for frac in (1 / 16, 1 / 4, 1.0):
    xyz = full[:int(len(full) * frac)]
    d = torch.from_numpy(xyz).cuda()
    rec, native_ms, native_err = {}, None, None
    for rep in range(3):
      torch.cuda.synchronize()
      e0 = ev()
      coords_raw, minmax = _abi.quantize_points(d, 0.02)
      spec = _abi.keyspec_build(minmax, 4, 32)
      table, sel, _, cnt = _abi.unique_first(coords_raw, spec)
      n = _abi.read_count(cnt)
      coords = _abi.gather_rows_i32(coords_raw, sel[:n], n)
      e1 = ev()
      with torch.no_grad():
        coords._dgr_manager = ME.CoordinateManager(_parts=(coords, spec, table))
        x = ME.SparseTensor(torch.ones(n, 1, device='cuda'), coordinates=coords, device='cuda')
        layers, total, _ = model._plan(x.coordinate_manager, x.coordinate_map_key)
        e2 = ev()
        F = model.forward_fused(x).F
        e3 = ev()
      torch.cuda.synchronize()
      pairs = sum(km.n_pairs for _, _, km, _ in layers)
      if rep == 2:       # the same forward pass through the native executor (one C call, one host read)
        Fn = net.forward(ctx, coords)
        t0 = time.perf_counter()
        Fn = net.forward(ctx, coords)
        native_ms = (time.perf_counter() - t0) * 1e3
        native_err = float((Fn - F).abs().max())
        del Fn
      rec = dict(n_raw=len(xyz), n_voxels=n, voxelise_ms=e0.elapsed_time(e1),
                 voxelise_GBps=len(xyz) * (24 + 16 + 12) / e0.elapsed_time(e1) / 1e6,
                 kernel_maps_ms=e1.elapsed_time(e2), fcgf_convs_ms=e2.elapsed_time(e3),
                 kernel_map_pairs_total=int(pairs), native_forward_ms_incl_maps=native_ms,
                 native_vs_operator_path_max_abs=native_err, native_arena_bytes=ctx.stats()['arena_high_water'])
    # kNN of the cloud against itself shifted (same size): N x N x 64
    if n <= 300_000:
      F1 = F.roll(1, 0).contiguous()
      for _ in range(2):
        _abi.knn_top1(F, F1)
      a = ev()
      _abi.knn_top1(F, F1)
      b = ev()
      torch.cuda.synchronize()
      rec.update(knn_ms=a.elapsed_time(b), knn_Tpairs_per_s=n * n / a.elapsed_time(b) / 1e9)
    out.append(rec)
    del x, F
    torch.cuda.empty_cache()
  # the full 1M x 1M x 64 kNN once
  n = out[-1]['n_voxels']
  Fa = torch.nn.functional.normalize(torch.randn(n, 64, device='cuda'), dim=1)
  Fb = torch.nn.functional.normalize(torch.randn(n, 64, device='cuda'), dim=1)
  a = ev()
  _abi.knn_top1(Fa, Fb)
  b = ev()
  torch.cuda.synchronize()
  out[-1].update(knn_ms=a.elapsed_time(b), knn_Tpairs_per_s=n * n / a.elapsed_time(b) / 1e9)
  return out


if __name__ == '__main__':
  res = dict(config3_kitti_shape_register=config3(), config5_stress_sweep=config5())
  os.makedirs('gpurun_out', exist_ok=True)
  json.dump(res, open('gpurun_out/configs_r02.json', 'w'), indent=1)
  print(json.dumps(res, indent=1))
