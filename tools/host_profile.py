import cProfile, pstats, os, sys, types, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepglobalregistration_b200 import synthetic as syn
from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration
state = syn.make_checkpoint(0)
dgr = DeepGlobalRegistration(types.SimpleNamespace(weights=state, clip_weight_thresh=0.05, verbose=False))
dgr.use_icp = False   # profile the benchmarked unit (through the refinement)
xyz0, xyz1, T = syn.room_pair(0, n_raw=250000)
for _ in range(6): dgr.register(xyz0, xyz1)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10): dgr.register(xyz0, xyz1)
torch.cuda.synchronize()
print('ms/pair', (time.perf_counter() - t) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): dgr.register(xyz0, xyz1)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(28)
