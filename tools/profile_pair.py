"""One register() of a 3DMatch-shape pair between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
Usage: python tools/profile_pair.py [n_raw]"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepglobalregistration_b200 import synthetic as syn
from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration

n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
state = syn.make_checkpoint(0)
dgr = DeepGlobalRegistration(types.SimpleNamespace(weights=state, clip_weight_thresh=0.05, verbose=False))
dgr.use_icp = False   # profile the benchmarked unit (through the refinement)
xyz0, xyz1, T = syn.room_pair(0, n_raw=n_raw)
dgr.register(xyz0, xyz1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
dgr.register(xyz0, xyz1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done', dgr.last_info)
