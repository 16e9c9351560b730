"""Generate tests/golden/reference_kitti.npz by running the UNMODIFIED reference
dataloader/kitti_loader.py::KITTINMPairDataset (pair selection, velo2cam, odometry parsing) on a
synthetic KITTI odometry tree (poses + empty scan files).  The loader imports MinkowskiEngine and
open3d at module level; deepglobalregistration_b200.shims provides import-level stand-ins (nothing
of them executes here).  /root/reference exists only in the build container, hence committed vectors.

    python tests/golden/make_golden_kitti.py

The scan files of a drive end exactly at the partner frame of its last pair (its pose file has one
more line): the reference's loop otherwise evaluates `empty_array in list`, an error under the numpy
installed here (a warning in the reference's own environment, where it simply ends the drive).
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')


def trajectory(g, n):
  """Camera-0 poses of a car: forward = camera z, yawing slowly about camera y, 0.7-1.4 m per frame."""
  P = np.tile(np.eye(4), (n, 1, 1))
  yaw, pos = 0.0, np.zeros(3)
  for k in range(n):
    yaw += g.normal(0, 0.02)
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    pos = pos + R @ np.array([0.0, g.normal(0, 0.01), g.uniform(0.7, 1.4)])
    P[k, :3, :3], P[k, :3, 3] = R, pos
  return P


def main():
  from deepglobalregistration_b200 import evaluate as ev
  from deepglobalregistration_b200 import shims
  shims.install()
  os.chdir('/root/reference')                           # DATA_FILES are relative paths
  import dataloader.kitti_loader as kl
  g = np.random.default_rng(0)
  tmp = tempfile.mkdtemp()
  out = {}
  for drive, n in ((8, 170), (9, 120), (10, 90)):
    P = trajectory(g, n)
    # cut the drive at the partner of its last pair (see the module docstring)
    pos, cur, end = P[:, :3, 3], 0, None
    while True:
      far = np.flatnonzero(np.linalg.norm(pos[cur:cur + 100] - pos[cur], axis=1) > 10)
      if len(far) == 0:
        break
      end = int(far[0]) + cur - 1
      cur = end + 1
    P = P[:end + 2]                                     # poses: one frame more than there are scans
    vel = os.path.join(tmp, 'dataset', 'sequences', f'{drive:02d}', 'velodyne')
    os.makedirs(vel)
    os.makedirs(os.path.join(tmp, 'dataset', 'poses'), exist_ok=True)
    for k in range(len(P) - 1):
      open(os.path.join(vel, f'{k:06d}.bin'), 'wb').close()
    np.savetxt(os.path.join(tmp, 'dataset', 'poses', f'{drive:02d}.txt'), P[:, :3, :].reshape(len(P), 12), fmt='%.9e')
    out[f'poses_{drive:02d}'] = np.loadtxt(os.path.join(tmp, 'dataset', 'poses', f'{drive:02d}.txt'))
  cfg = types.SimpleNamespace(kitti_dir=tmp, icp_cache_path='icp', voxel_size=0.3,
                              positive_pair_search_voxel_size_multiplier=1.5, min_scale=0.8, max_scale=1.2,
                              rotation_range=360, kitti_max_time_diff=3)
  ds = kl.KITTINMPairDataset('test', transform=None, random_rotation=False, random_scale=False, config=cfg)
  files = np.array(ds.files, dtype=np.int64)
  Ms = []
  for drive, t0, t1 in ds.files:
    odo = ds.get_video_odometry(drive, [t0, t1])
    positions = [ds.odometry_to_positions(o) for o in odo]
    # the expression of dataloader/kitti_loader.py:147-148, on the reference's own velo2cam / positions
    M = (ds.velo2cam @ positions[0].T @ np.linalg.inv(positions[1].T) @ np.linalg.inv(ds.velo2cam)).T
    Ms.append(M)
  out.update(files=files, M=np.stack(Ms), velo2cam_T=np.asarray(ds.velo2cam))
  np.savez(os.path.join(HERE, 'reference_kitti.npz'), **out)
  print('pairs per drive:', {d: int((files[:, 0] == d).sum()) for d in (8, 9, 10)}, 'frames:',
        {k: len(v) for k, v in out.items() if k.startswith('poses')})
  mine = ev.kitti_pairs(tmp)
  print('own selection identical:', [(int(p.group[-2:]), int(os.path.basename(p.file0)[:-4]), int(os.path.basename(p.file1)[:-4]))
                                     for p in mine] == [tuple(int(x) for x in f) for f in files])


if __name__ == '__main__':
  main()
