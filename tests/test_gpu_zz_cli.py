"""The file-level callers on a GPU: demo CLI (the reference's demo.py flow:
read two point-cloud files, register, print the pose) and the pair-list evaluation driver, both
through checkpoint and point-cloud FILES.  Runs last (file name) - everything it calls is covered
stage by stage in the other GPU test files."""
import json
import types

import numpy as np
import pytest
import torch

from deepglobalregistration_b200 import io as dio
from deepglobalregistration_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
EXTENT = (1.8, 1.5, 1.25)


@pytest.fixture(scope='module')
def files(tmp_path_factory):
  root = tmp_path_factory.mktemp('cli')
  state = syn.make_checkpoint(0)
  torch.save(state, root / 'ckpt.pth')
  xyz0, xyz1, T_gt = syn.room_pair(2, n_raw=20000, extent=EXTENT)
  dio.write_ply(root / 'a.ply', xyz0, dtype='double')
  dio.write_ply(root / 'b.ply', xyz1, dtype='double')
  np.savez(root / 'b.npz', pcd=xyz1)
  return types.SimpleNamespace(root=root, ckpt=str(root / 'ckpt.pth'), state=state, xyz0=xyz0, xyz1=xyz1, T_gt=T_gt)


def test_demo_cli_matches_the_class(files, capsys):
  from deepglobalregistration_b200 import demo
  from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration
  out = files.root / 'moved.ply'
  T = demo.main(['--pcd0', str(files.root / 'a.ply'), '--pcd1', str(files.root / 'b.ply'), '--weights', files.ckpt,
                 '--json', '--out', str(out)])
  line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
  assert np.allclose(np.array(line['T']), T) and line['branch'] in ('procrustes', 'safeguard')
  d = DeepGlobalRegistration(types.SimpleNamespace(weights=files.state, clip_weight_thresh=0.05, verbose=False))
  T_direct = d.register(files.xyz0, files.xyz1)          # float64 arrays == what the PLY files hold
  te, re = syn.rte_rre(T, T_direct)
  assert te <= 1e-3 and re <= 1e-3, (te, re)       # two GPU runs differ only by atomic summation order
  moved, _ = dio.read_ply(out)
  assert np.allclose(moved, syn.apply_se3(T, files.xyz0), atol=1e-9)


def test_evaluation_driver_on_files(files, capsys):
  from deepglobalregistration_b200 import evaluate as ev
  T = files.T_gt
  (files.root / 'pairs.txt').write_text(
      f'a.ply b.npz {" ".join(repr(float(x)) for x in T.reshape(-1))} room\n'
      'a.ply b.ply\n')
  ev.main(['--pair_list', str(files.root / 'pairs.txt'), '--weights', files.ckpt, '--out_dir', str(files.root)])
  summary = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
  assert summary['pairs'] == 2 and summary['with_ground_truth'] == 1 and summary['world_size'] == 1
  saved = np.load(files.root / 'dgr-b200-stats.npz', allow_pickle=True)
  assert saved['stats'].shape == (1, 2, 5) and saved['poses'].shape == (2, 4, 4)
  # both lines are the same pair through different file formats (.npz / .ply, float64 both)
  te, re = syn.rte_rre(saved['poses'][0], saved['poses'][1])
  assert te <= 1e-3 and re <= 1e-3
  assert saved['stats'][0, 0, 3] > 0           # seconds per pair were recorded
