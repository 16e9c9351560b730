"""Per-stage CUDA-event timing of one 3DMatch-shape pair (development aid; bench.py is the
contract).  Usage: python tools/stage_probe.py [n_raw] [reps]"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepglobalregistration_b200 import _abi, synthetic as syn
from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration


def main():
  n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
  reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
  state = syn.make_checkpoint(0)
  cfg = types.SimpleNamespace(weights=state, clip_weight_thresh=0.05, verbose=False)
  dgr = DeepGlobalRegistration(cfg)
  xyz0, xyz1, T = syn.room_pair(0, n_raw=n_raw)

  def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e

  for rep in range(reps):
    torch.cuda.synchronize()
    marks = [('start', ev())]
    with torch.no_grad():
      p0, c0, f0 = dgr.preprocess(xyz0, 0)
      p1, c1, f1 = dgr.preprocess(xyz1, 1)
      marks.append(('preprocess x2', ev()))
      F0 = dgr.fcgf_feature_extraction(f0, c0)
      marks.append(('fcgf 0', ev()))
      F1 = dgr.fcgf_feature_extraction(f1, c1)
      marks.append(('fcgf 1', ev()))
      idx1 = _abi.knn_top1(F0, F1)
      marks.append(('knn', ev()))
      c6 = _abi.inlier_coords(c0, c1, idx1)
      logit = dgr.inlier_prediction(torch.ones(len(idx1), 1, device='cuda'), c6)
      marks.append(('inlier net', ev()))
      w, ws = _abi.sigmoid_clip_sum(logit, 0.05)
      res = _abi.se3_register(p0, p1, w.reshape(-1), idx1=idx1, quantization_size=0.1,
                              break_threshold_ratio=1e-4)
      marks.append(('procrustes+refine', ev()))
    torch.cuda.synchronize()
    if rep == reps - 1:
      print(f'N0={len(c0)} N1={len(c1)} wsum={float(ws.item()):.1f} refine={res.cpu().numpy()[12:].tolist()}')
      tot = marks[0][1].elapsed_time(marks[-1][1])
      for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
        print(f'  {name:20s} {a.elapsed_time(b):9.3f} ms')
      print(f'  {"total":20s} {tot:9.3f} ms   ({1000.0 / tot:.1f} pairs/s)')
  import time
  torch.cuda.synchronize()
  t = time.time()
  for _ in range(3):
    dgr.register(xyz0, xyz1)
  torch.cuda.synchronize()
  print(f'register() host-to-host: {(time.time() - t) / 3 * 1e3:.2f} ms')


if __name__ == '__main__':
  main()
