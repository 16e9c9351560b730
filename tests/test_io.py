"""Readers of the formats either side of register() and the open3d
stand-in demo.py needs."""
import sys

import numpy as np
import pytest

from deepglobalregistration_b200 import io as dio


@pytest.fixture
def cloud():
  g = np.random.default_rng(0)
  return g.normal(size=(257, 3)) * [3.0, 2.0, 1.0]


@pytest.mark.parametrize('fmt', ['ascii', 'binary_little_endian', 'binary_big_endian'])
@pytest.mark.parametrize('dtype', ['float', 'double'])
def test_ply_round_trip(tmp_path, cloud, fmt, dtype):
  path = tmp_path / 'c.ply'
  red = np.arange(len(cloud)) % 256
  dio.write_ply(path, cloud, fmt=fmt, dtype=dtype, red=red.astype(np.uint8), intensity=cloud[:, 0].astype(np.float32))
  pts, extra = dio.read_ply(path)
  want = cloud.astype(np.float32).astype(np.float64) if dtype == 'float' else cloud
  assert pts.dtype == np.float64 and np.array_equal(pts, want)      # bit-exact through the file
  assert np.array_equal(extra['red'], red)
  pcd = dio.read_point_cloud(str(path))
  assert len(pcd) == len(cloud) and pcd.has_points() and 'red' in pcd.attributes


def test_ply_with_faces_comments_and_property_order(tmp_path):
  """x/y/z found by name among other properties; a face element with a list property before and
  after the vertices is skipped in both encodings."""
  head = ('ply\nformat {f} 1.0\ncomment made by hand\nelement face 2\nproperty list uchar int vertex_indices\n'
          'element vertex 3\nproperty uchar red\nproperty float z\nproperty float x\nproperty double y\n'
          'element edge 1\nproperty int a\nproperty int b\nend_header\n')
  want = np.array([[1.5, 2.5, 0.5], [4.0, -5.0, 3.0], [7.0, 8.0, 6.0]])
  a = tmp_path / 'a.ply'
  a.write_text(head.format(f='ascii') + '3 0 1 2\n4 0 1 2 2\n' +
               '9 0.5 1.5 2.5\n8 3 4 -5\n7 6 7 8\n' + '0 1\n')
  assert np.array_equal(dio.read_ply(a)[0], want)
  b = tmp_path / 'b.ply'
  body = b''
  for face in ([0, 1, 2], [0, 1, 2, 2]):
    body += np.uint8(len(face)).tobytes() + np.array(face, '<i4').tobytes()
  rec = np.zeros(3, dtype=[('red', 'u1'), ('z', '<f4'), ('x', '<f4'), ('y', '<f8')])
  rec['red'], rec['z'], rec['x'], rec['y'] = [9, 8, 7], want[:, 2], want[:, 0], want[:, 1]
  b.write_bytes(head.format(f='binary_little_endian').encode() + body + rec.tobytes() + np.array([0, 1], '<i4').tobytes())
  pts, extra = dio.read_ply(b)
  assert np.array_equal(pts, want) and np.array_equal(extra['red'], [9, 8, 7])


def test_ply_errors(tmp_path):
  p = tmp_path / 'x.ply'
  p.write_text('plyx\n')
  with pytest.raises(ValueError, match='not a PLY'):
    dio.read_ply(p)
  p.write_text('ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nend_header\n0 0\n1 1\n')
  with pytest.raises(ValueError, match='x / y / z'):
    dio.read_ply(p)
  p.write_bytes(b'ply\nformat binary_little_endian 1.0\nelement vertex 5\nproperty float x\nproperty float y\n'
                b'property float z\nend_header\n' + b'\0' * 20)
  with pytest.raises(ValueError, match='truncated'):
    dio.read_ply(p)
  with pytest.raises(ValueError, match='unsupported'):
    dio.read_points(str(tmp_path / 'x.obj'))


def test_kitti_3dmatch_npy_text(tmp_path, cloud):
  xyzr = np.concatenate([cloud, np.ones((len(cloud), 1))], 1).astype(np.float32)
  xyzr.tofile(tmp_path / '000000.bin')
  pts, refl = dio.read_kitti_bin(tmp_path / '000000.bin')
  assert pts.dtype == np.float32 and np.array_equal(pts, xyzr[:, :3]) and np.array_equal(refl, xyzr[:, 3])
  assert dio.read_points(str(tmp_path / '000000.bin')).dtype == np.float32      # the caller's dtype is kept
  (tmp_path / 'bad.bin').write_bytes(b'\0' * 10)
  with pytest.raises(ValueError):
    dio.read_kitti_bin(tmp_path / 'bad.bin')
  np.savez(tmp_path / 'frag.npz', pcd=cloud.astype(np.float32), color=np.zeros((len(cloud), 3)))
  assert np.array_equal(dio.read_points(str(tmp_path / 'frag.npz')), cloud.astype(np.float32))
  np.savez(tmp_path / 'nopcd.npz', xyz=cloud)
  with pytest.raises(ValueError, match='pcd'):
    dio.read_points(str(tmp_path / 'nopcd.npz'))
  np.save(tmp_path / 'c.npy', cloud)
  assert np.array_equal(dio.read_points(str(tmp_path / 'c.npy')), cloud)
  np.savetxt(tmp_path / 'c.xyz', cloud, fmt='%.17g')
  assert np.array_equal(dio.read_points(str(tmp_path / 'c.xyz')), cloud)
  assert isinstance(dio.read_point_cloud(str(tmp_path / 'c.npy')), dio.PointCloud)


def test_point_cloud_object(cloud):
  pcd = dio.PointCloud(cloud)
  T = np.eye(4)
  T[:3, :3] = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]
  T[:3, 3] = [1, 2, 3]
  moved = pcd.estimate_normals().transform(T).points
  assert np.allclose(moved, cloud @ T[:3, :3].T + T[:3, 3])
  with pytest.raises(ValueError):
    dio.PointCloud(np.zeros((4, 2)))
  assert len(dio.PointCloud()) == 0 and not dio.PointCloud().has_points()


def test_trajectory_round_trip(tmp_path):
  g = np.random.default_rng(1)
  poses = [([k, k + 1, 37], np.vstack([g.normal(size=(3, 4)), [0, 0, 0, 1]])) for k in range(5)]
  dio.write_trajectory(tmp_path / 'gt.log', poses)
  got = dio.read_trajectory(tmp_path / 'gt.log')
  assert len(got) == 5
  for (meta, mat), cp in zip(poses, got):
    assert cp.metadata == meta and np.array_equal(cp.pose, mat)        # %.17g round-trips float64
  # tabs and blank lines, as found in the 3DMatch gt.log files
  (tmp_path / 't.log').write_text('0\t1\t60\n1 0\t0 0\n0 1 0 0\n0 0 1 0\n0 0 0 1\n\n')
  cp, = dio.read_trajectory(tmp_path / 't.log')
  assert cp.metadata == [0, 1, 60] and np.array_equal(cp.pose, np.eye(4))
  (tmp_path / 'bad.log').write_text('0 1 2\n1 0 0 0\n')
  with pytest.raises(ValueError):
    dio.read_trajectory(tmp_path / 'bad.log')


def torch_cuda_available():
  import torch
  return torch.cuda.is_available()


def test_open3d_stand_in(tmp_path, cloud, monkeypatch):
  from deepglobalregistration_b200 import shims
  try:
    import open3d
    if not getattr(open3d, '__dgr_stub__', False):
      pytest.skip('real open3d present: the stand-in is not installed')
  except ImportError:
    pass
  monkeypatch.delitem(sys.modules, 'open3d', raising=False)
  shims.install()
  import open3d as o3d
  assert getattr(o3d, '__dgr_stub__', False)
  dio.write_ply(tmp_path / 'a.ply', cloud, dtype='double')
  pcd = o3d.io.read_point_cloud(str(tmp_path / 'a.ply'))
  pcd.estimate_normals()
  assert np.array_equal(np.asarray(pcd.points), cloud)
  q = o3d.geometry.PointCloud()
  q.points = o3d.utility.Vector3dVector(cloud[:10])
  assert len(q) == 10
  o3d.visualization.draw_geometries([pcd, q])
  assert o3d.io.write_point_cloud(str(tmp_path / 'b.ply'), q)
  assert np.array_equal(dio.read_ply(tmp_path / 'b.ply')[0], cloud[:10])
  # round 2: the registration pipeline the reference's own class calls is part of the stand-in (GPU-backed; the
  # calls themselves are exercised in tests/test_gpu_reference_on_shim.py)
  reg = o3d.pipelines.registration
  assert o3d.registration is reg and callable(reg.registration_icp) and \
      callable(reg.registration_ransac_based_on_correspondence)
  crit = reg.RANSACConvergenceCriteria(4000000, 80000)           # core/deep_global_registration.py:62
  assert crit.max_iteration == 4000000 and crit.confidence == 1.0
  assert reg.ICPConvergenceCriteria().max_iteration == 30
  assert o3d.utility.Vector2iVector(np.zeros((5, 2))).shape == (5, 2)
  if not torch_cuda_available():
    with pytest.raises(Exception):                                 # no CPU fallback
      reg.registration_icp(pcd, q, 0.1)
