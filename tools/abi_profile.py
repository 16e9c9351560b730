"""GPU time per C-ABI entry point over one warm register() of a 3DMatch-shape pair (CUDA events around
every call, in context and warm - complements the cold-cache, serialised ncu launch list).

    python tools/abi_profile.py [n_raw] [--icp]

Prints a table sorted by total time: calls, total ms, share of the sum, mean us per call; then the step time
and the part of it not covered by any ABI call (torch glue kernels, host gaps)."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepglobalregistration_b200 import _abi
from deepglobalregistration_b200 import synthetic as syn
from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration


def main():
  args = [a for a in sys.argv[1:] if not a.startswith('--')]
  n_raw = int(args[0]) if args else 250_000
  state = syn.make_checkpoint(0)
  dgr = DeepGlobalRegistration(types.SimpleNamespace(weights=state, clip_weight_thresh=0.05, verbose=False))
  dgr.use_icp = '--icp' in sys.argv
  xyz0, xyz1, _ = syn.room_pair(0, n_raw=n_raw)
  for _ in range(5):
    dgr.register(xyz0, xyz1)
  torch.cuda.synchronize()
  # un-instrumented step time first
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    dgr.register(xyz0, xyz1)
  e1.record()
  torch.cuda.synchronize()
  plain = e0.elapsed_time(e1) / 10
  _abi.CALL_PROFILE = []
  e0.record()
  dgr.register(xyz0, xyz1)
  e1.record()
  torch.cuda.synchronize()
  prof, _abi.CALL_PROFILE = _abi.CALL_PROFILE, None
  inst = e0.elapsed_time(e1)
  agg = {}
  for name, a, b in prof:
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += a.elapsed_time(b)
  total = sum(v[1] for v in agg.values())
  print(f'{"entry point":34s} {"calls":>6s} {"ms":>8s} {"share":>7s} {"us/call":>9s}')
  for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{name:34s} {n:6d} {ms:8.3f} {100 * ms / total:6.1f}% {1e3 * ms / n:9.1f}')
  print(f'{"sum over ABI calls":34s} {sum(v[0] for v in agg.values()):6d} {total:8.3f}')
  print(f'step: {plain:.3f} ms un-instrumented, {inst:.3f} ms instrumented; outside ABI calls (torch glue, gaps): '
        f'{inst - total:.3f} ms; N0={dgr.last_info.get("n0")} N1={dgr.last_info.get("n1")} branch={dgr.last_branch}')


if __name__ == '__main__':
  main()
