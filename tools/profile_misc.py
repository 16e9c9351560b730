"""One register() with ICP and one safeguard (RANSAC, 400k hypotheses) between cudaProfilerStart/Stop: the kernels
of the widened rows for `ncu --profile-from-start off`."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepglobalregistration_b200 import synthetic as syn
from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration

state = syn.make_checkpoint(0)
dgr = DeepGlobalRegistration(types.SimpleNamespace(weights=state, clip_weight_thresh=0.05, verbose=False))
dgr.use_icp = True
xyz0, xyz1, T = syn.room_pair(0, n_raw=250_000)
dgr.register(xyz0, xyz1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
dgr.register(xyz0, xyz1)
dgr.clip_weight_thresh, dgr.safeguard_max_iteration = 0.999999, 400000      # gate closed -> RANSAC
dgr.register(xyz0, xyz1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done', dgr.last_branch, dgr.last_info)
