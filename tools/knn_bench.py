import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepglobalregistration_b200 import _abi
torch.manual_seed(0)
for n0, n1, c in ((51381, 39881, 32), (100000, 100000, 32), (50000, 50000, 64)):
    F0 = torch.nn.functional.normalize(torch.randn(n0, c, device='cuda'), dim=1)
    F1 = torch.nn.functional.normalize(torch.randn(n1, c, device='cuda'), dim=1)
    for _ in range(3): idx = _abi.knn_top1(F0, F1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): idx = _abi.knn_top1(F0, F1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f'knn {n0}x{n1}x{c}: {ms:.3f} ms  {n0*n1*c/ms/1e9:.2f} T pair-terms/s  checksum {int(idx.long().sum())}')
print('--- tc vs simt')
for n0, n1, c in ((51381, 39881, 32), (100000, 100000, 32), (50000, 50000, 64)):
    F0 = torch.nn.functional.normalize(torch.randn(n0, c, device='cuda'), dim=1)
    F1 = torch.nn.functional.normalize(torch.randn(n1, c, device='cuda'), dim=1)
    for mode in ('tc', 'simt'):
        for _ in range(2): idx = _abi.knn_top1(F0, F1, mode=mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): idx = _abi.knn_top1(F0, F1, mode=mode)
        e1.record(); torch.cuda.synchronize()
        print(f'{mode:5s} knn {n0}x{n1}x{c}: {e0.elapsed_time(e1)/5:.3f} ms checksum {int(idx.long().sum())}')
