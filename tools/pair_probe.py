"""cta_group::2 sparse-conv variant (DGR_TC_VARIANT=3) against the single-CTA kernel:
  1. correctness on small maps (run in a child process with a timeout: a protocol bug hangs),
  2. per-layer event timings of one 3DMatch-shape register() with each variant.
Usage: python tools/pair_probe.py [check|time|all]"""
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check():
  import numpy as np
  import torch
  from deepglobalregistration_b200 import _abi as abi
  from deepglobalregistration_b200.me.coords import CoordinateManager, CoordinateMapKey
  ok = True
  for D, cin, cout, npts in ((3, 32, 32, 500), (3, 64, 128, 4000), (3, 256, 256, 4000), (6, 64, 240, 3000),
                             (3, 128, 16, 2000)):
    g = np.random.default_rng(cin + cout)
    ext = 10 if D == 3 else 3
    coords = np.unique(g.integers(0, ext, size=(npts, D)), axis=0)
    coords = np.concatenate([np.zeros((len(coords), 1), np.int64), coords], 1).astype(np.int32)
    n = len(coords)
    tg = torch.Generator().manual_seed(5)
    feat = torch.randn(n, cin, generator=tg).cuda()
    W = (torch.randn(3 ** D, cin, cout, generator=tg) / np.sqrt(cin * 8)).cuda().contiguous()
    man = CoordinateManager(torch.from_numpy(coords).cuda())
    _, km = man.kernel_map(CoordinateMapKey(1), 1, 3)
    Wt = abi.pack_weight_tf32(W, 3 ** D, cin, cout)
    ref = torch.zeros(n, cout, device='cuda')
    abi.spconv_tc_fwd(feat, Wt, km, ref, passes=3, cluster=1)
    for rep in range(3):
      out = torch.zeros(n, cout, device='cuda')
      abi.spconv_tc_fwd(feat, Wt, km, out, passes=3, cluster=3)
      torch.cuda.synchronize()
      err = float((out - ref).abs().max() / (1 + ref.abs().max()))
      print(f'D={D} cin={cin} cout={cout} n={n} pairs={km.n_pairs} rep={rep}: max rel err {err:.2e}', flush=True)
      ok &= err < 1e-5
  print('CHECK', 'PASS' if ok else 'FAIL', flush=True)
  return ok


def time_layers():
  import torch
  from deepglobalregistration_b200 import _abi as abi
  from deepglobalregistration_b200 import synthetic as syn
  from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration
  state = syn.make_checkpoint(0)
  dgr = DeepGlobalRegistration(types.SimpleNamespace(weights=state, clip_weight_thresh=0.05, verbose=False))
  dgr.use_icp = False
  xyz0, xyz1, _ = syn.room_pair(0, n_raw=250_000)
  poses = {}
  for variant in (1, 3):
    abi.TC_VARIANT = variant
    for _ in range(3):
      T = dgr.register(xyz0, xyz1)
    abi.CONV_PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    T = dgr.register(xyz0, xyz1)
    e1.record()
    torch.cuda.synchronize()
    prof, abi.CONV_PROFILE = abi.CONV_PROFILE, None
    rows = [(name, a.elapsed_time(b), fl, nb) for name, a, b, fl, nb in prof]
    tc = [r for r in rows if r[0] == 'spconv_tc_kernel']
    tot = sum(r[1] for r in tc)
    big = sorted(tc, key=lambda r: -r[1])[:6]
    print(f'prefetch {os.environ.get("DGR_TC_PREFETCH", "default")} variant {variant}: step {e0.elapsed_time(e1):.2f} ms, tensor-core convs {tot:.2f} ms in {len(tc)} launches; '
          f'top: ' + ', '.join(f'{r[1]:.2f} ms ({r[2] / r[1] / 1e9:.0f} TF)' for r in big), flush=True)
    poses[variant] = T
  import numpy as np
  print('pose difference between variants:', float(np.abs(poses[1] - poses[3]).max()))


if __name__ == '__main__':
  mode = sys.argv[1] if len(sys.argv) > 1 else 'all'
  if mode == 'check':
    sys.exit(0 if check() else 1)
  if mode == 'time':
    time_layers()
    sys.exit(0)
  env = dict(os.environ, DGR_TC_EPILOGUE='1')
  r = subprocess.run(['timeout', '-s', 'KILL', '120', sys.executable, os.path.abspath(__file__), 'check'], env=env,
                     capture_output=True, text=True)
  print(f'coalesced epilogue: check exit code {r.returncode}:', r.stdout.strip().splitlines()[-3:], r.stderr[-300:],
        flush=True)
  for epi in ('0', '1'):
    env = dict(os.environ, DGR_TC_EPILOGUE=epi)
    print('DGR_TC_EPILOGUE =', epi, flush=True)
    r = subprocess.run(['timeout', '-s', 'KILL', '200', sys.executable, os.path.abspath(__file__), 'time'], env=env)
    print('time exit code', r.returncode, flush=True)
