"""Generate tests/golden/reference_io.npz + tests/golden/sample_gt.log by running the UNMODIFIED
reference code from /root/reference (only present in the build container, hence committed vectors):

* util/file.py::read_trajectory on a hand-written gt.log (tabs, spaces, exponents, as in the
  3DMatch evaluation files) -> pins deepglobalregistration_b200.io.read_trajectory;
* scripts/test_3dmatch.py::rte_rre (the evaluation criterion) -> pins evaluate.rte_rre.  That
  script imports open3d / MinkowskiEngine at module level, so only the function definition is
  compiled out of the file's syntax tree - its body runs unchanged.

    python tests/golden/make_golden_io.py
"""
import ast
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, REF)
from util.file import read_trajectory                    # noqa: E402


def reference_function(path, name, namespace):
  tree = ast.parse(open(path).read())
  node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
  exec(compile(ast.Module([node], []), path, 'exec'), namespace)
  return namespace[name]


def main():
  g = np.random.default_rng(0)
  log = os.path.join(HERE, 'sample_gt.log')
  blocks = []
  for k in range(4):
    T = np.eye(4)
    A = g.normal(size=(3, 3))
    Q, _ = np.linalg.qr(A)
    T[:3, :3] = Q * np.sign(np.linalg.det(Q))
    T[:3, 3] = g.normal(size=3) * (10.0 ** (k - 1))
    sep = '\t' if k % 2 == 0 else ' '
    fmt = '%.8e' if k % 2 == 0 else '%.10f'
    blocks.append(sep.join(str(v) for v in (k, k + 2, 60)) + '\n' +
                  ''.join(sep.join(fmt % x for x in row) + '\n' for row in T))
  with open(log, 'w') as fh:
    fh.write(''.join(blocks))
  traj = read_trajectory(log)
  out = dict(traj_meta=np.array([t.metadata for t in traj]), traj_pose=np.stack([t.pose for t in traj]))

  rte_rre = reference_function(os.path.join(REF, 'scripts', 'test_3dmatch.py'), 'rte_rre', dict(np=np, math=math))
  preds, gts, res = [], [], []
  for k in range(12):
    def pose(scale_deg, scale_m):
      ang = np.deg2rad(scale_deg) * g.uniform(0.2, 1.0)
      ax = g.normal(size=3)
      ax /= np.linalg.norm(ax)
      K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
      T = np.eye(4)
      T[:3, :3] = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
      T[:3, 3] = g.normal(size=3) * scale_m
      return T
    Tg = pose(60, 1.0)
    Tp = Tg.copy() if k == 0 else pose(0.1 if k < 4 else 30, 0.001 if k < 4 else 0.4) @ Tg
    preds.append(Tp)
    gts.append(Tg)
    res.append(np.asarray(rte_rre(Tp, Tg, 0.3, 15), dtype=np.float64))
  out.update(metric_pred=np.stack(preds), metric_gt=np.stack(gts), metric_out=np.stack(res),
             metric_none=np.asarray(rte_rre(None, gts[0], 0.3, 15), dtype=np.float64))
  np.savez(os.path.join(HERE, 'reference_io.npz'), **out)
  print({k: v.shape for k, v in out.items()}, '\nsuccesses:', int(np.stack(res)[:, 0].sum()), 'of', len(res))


if __name__ == '__main__':
  main()
