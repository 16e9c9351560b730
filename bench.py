#!/usr/bin/env python
"""bench.py - scan-pairs/sec of DGR's pairwise-registration hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one full register() of one synthetic 3DMatch-shape scan pair (BASELINE.json
configs[1]/[3] shape: ~50k voxels per cloud at 0.05 m, FCGF feature dim 32, ResUNetBN2C for
both networks): voxelise x2 -> FCGF x2 -> feature kNN -> 6-D inlier network -> weights ->
weighted Procrustes + SE(3) refinement (SURVEY 8(d)'s unit).  The ICP fine-tune and the RANSAC
safeguard are built but outside the benchmarked unit on both arms.  Pairs are independent: each rank registers its own pairs (weak scaling) and
the poses are all-gathered over NCCL at the end of the timed region.

Prints ONE JSON line (rank 0).  `value` = pairs/s with the raw scans resident in HBM;
`e2e` = pairs/s through DeepGlobalRegistration.register(host ndarrays) including the H2D
copy of both scans and the D2H read of the pose.  `--impl reference` times the CPU oracle
port of the same path (MinkowskiEngine cannot be installed offline) on a bounded sample.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import numpy as np
import torch

from deepglobalregistration_b200 import synthetic as syn

WORKLOAD = '3dmatch_shape_pair_register'
N_RAW = 250_000               # raw points per scan -> ~50k voxels at 0.05 m
SAMPLE_N_RAW = 8_000          # CPU sample: same generator, ~4k voxels per cloud
SAMPLE_EXTENT = (1.5, 1.2, 1.0)
REF_TIME_BUDGET_S = 240       # --impl reference: stop starting full-size pairs once this is exceeded (>= 1 runs)
REF_MIN_N_RAW = 1500          # the reference arm's untimed warm-up sample
VOXEL = 0.05
POOL = 3                      # distinct pairs per rank, cycled over the steps


_emit = print


def log(*a):
  print(*a, file=sys.stderr, flush=True)


def base_config(n_gpus):
  """The SAME dict on both arms (the driver compares them): anything that differs between the arms
  (voxel counts seen, sample notes) goes into other keys of the line."""
  return {'workload': WORKLOAD, 'n_raw_points_per_scan': N_RAW, 'voxel_size': VOXEL, 'feat_dim': 32,
          'fcgf_model': 'ResUNetBN2C(D=3,conv1_k=7)', 'inlier_model': 'ResUNetBN2C(D=6,conv1_k=3)',
          'conv_arithmetic': 'tcgen05 3xTF32 (fp32-accurate) + fp32 FFMA for conv1',
          'parallelism': f'pair-sharded dp{n_gpus}', 'pairs_per_step_per_gpu': 1,
          'excluded_on_both_arms': 'ICP fine-tune and RANSAC safeguard (both built; the benchmarked unit is SURVEY 8(d)\'s: through the SE(3) refinement, and the benchmark pairs take the Procrustes branch)',
          'l2_policy': 'inputs larger than L2: every step streams the 944 MB inlier-net weights '
                       '(L2 = 126 MB) and cycles through %d distinct pairs' % POOL}


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
  """SM clock and throttle reasons during the timed region.  Uses NVML in-process (a query costs
  microseconds); spawning nvidia-smi in a loop was measured to stall the CUDA driver for ~200 ms
  per query on these hosts and is only the fallback."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index, period=0.1):
    self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
    self.period = period
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nvml = pynvml
      # torch's device index follows CUDA_VISIBLE_DEVICES; map through the UUID-free common case
      vis = os.environ.get('CUDA_VISIBLE_DEVICES')
      phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
      self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
      self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
      self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
      self.thread.start()
      return
    except Exception as e:   # noqa: BLE001
      log('NVML sampler unavailable (%s); falling back to nvidia-smi' % e)
      self.nvml = None
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                    '--format=csv,noheader,nounits', '-lms', '500'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception as e:   # noqa: BLE001
      log('clock sampler unavailable:', e)

  def _poll_nvml(self):
    n = self.nvml
    names = {'hw_slowdown': getattr(n, 'nvmlClocksEventReasonHwSlowdown', 0x8),
             'hw_thermal_slowdown': getattr(n, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
             'sw_thermal_slowdown': getattr(n, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
             'sw_power_cap': getattr(n, 'nvmlClocksEventReasonSwPowerCap', 0x4)}
    get_reasons = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons',
                          getattr(n, 'nvmlDeviceGetCurrentClocksThrottleReasons', None))
    while not self.stop_flag:
      try:
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        mask = int(get_reasons(self.handle)) if get_reasons else 0
        active = [k for k, bit in names.items() if mask & bit]
        self.rows.append((time.time(), sm, active))
      except Exception:   # noqa: BLE001
        pass
      time.sleep(self.period)

  def _read(self):
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for line in self.proc.stdout:
      f = [x.strip() for x in line.split(',')]
      try:
        self.max_sm = float(f[1])
        self.rows.append((time.time(), float(f[0]),
                          [nm for nm, v in zip(names, f[3:7]) if v.lower().startswith('active')]))
      except (ValueError, IndexError):
        continue

  def stop(self, t0, t1):
    self.stop_flag = True
    if self.proc is not None:
      self.proc.terminate()
    if self.proc is None and self.nvml is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
    sm, reasons = [], set()
    for ts, clk, active in self.rows:
      if t0 - 0.05 <= ts <= t1 + 0.15:
        sm.append(clk)
        reasons.update(active)
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': getattr(self, 'max_sm', None),
            'samples': len(sm), 'reasons': sorted(reasons),
            'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# ------------------------------------------------------------------------------------------
# CPU baseline (oracle port) - also the --impl reference arm
# ------------------------------------------------------------------------------------------
def cgroup_throttled_ms():
  """Cumulative time this container's CPU cgroup has been throttled (ms), or None."""
  for path in ('/sys/fs/cgroup/cpu.stat', '/sys/fs/cgroup/cpu/cpu.stat'):
    try:
      for line in open(path):
        k, v = line.split()
        if k == 'throttled_usec':
          return int(v) / 1e3
        if k == 'throttled_time':
          return int(v) / 1e6
    except Exception:   # noqa: BLE001
      continue
  return None


def effective_cpus():
  """Host threads this process may really use: affinity mask and cgroup CPU quota, not nproc."""
  n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
  try:
    quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
    if quota != 'max':
      n = min(n, max(1, int(np.ceil(int(quota) / int(period)))))
  except Exception:   # noqa: BLE001
    try:
      q = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())
      p = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
      if q > 0:
        n = min(n, max(1, int(np.ceil(q / p))))
    except Exception:   # noqa: BLE001
      pass
  return n


CPU_THREADS = None


def cpu_threads():
  """Threads for the CPU arm: all usable cores, capped at 32 (the per-offset mm / index_add
  of the oracle stop scaling well before that and oversubscription is catastrophic)."""
  global CPU_THREADS
  if CPU_THREADS is None:
    CPU_THREADS = max(1, min(effective_cpus(), 32))
    torch.set_num_threads(CPU_THREADS)
  return CPU_THREADS


def cpu_sample_time(state, seed, reps=1, n_raw=SAMPLE_N_RAW):
  """Seconds per pair of the CPU oracle on the bounded sample."""
  from oracle import pipeline as op
  cpu_threads()
  xyz0, xyz1, _ = syn.room_pair(seed, n_raw=n_raw, extent=SAMPLE_EXTENT)
  ts, info = [], {}
  for _ in range(reps):
    t = time.perf_counter()
    _, taps = op.register(state, xyz0, xyz1)
    ts.append(time.perf_counter() - t)
    info = {'n0': int(len(taps['coords0'])), 'n1': int(len(taps['coords1'])), 'branch': taps['branch']}
  return float(np.median(ts)), info


def sample_desc(info, n_raw=SAMPLE_N_RAW):
  return (f'1 pair of the same generator at {n_raw} raw points/scan -> N0={info["n0"]}, '
          f'N1={info["n1"]} voxels (the workload has ~51k/~40k); pairs/s of the SAMPLE, not extrapolated; '
          'CPU path = oracle port (torch-CPU index_select/mm/index_add per kernel offset, the algorithm '
          "of MinkowskiEngine's CPU backend) + restated kNN / Procrustes / Adam refinement")


def fixture_parity(T, key='T_refined'):
  """TE [m] / RE [rad] of a pose of the bench's first pair (rank 0, seed 0, 250k raw points) against the CPU
  oracle's pose stored in tests/golden/fullsize_config2.npz (tests/golden/make_golden_fullsize.py)."""
  try:
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'fullsize_config2.npz'))
    te, re = syn.rte_rre(np.asarray(T, np.float64).reshape(4, 4), g[key])
    return {'te_m': te, 're_rad': re, 'tolerance': '1e-3 m / 1e-3 rad (north_star)', 'within_tolerance': bool(te <= 1e-3 and re <= 1e-3),
            'against': 'CPU oracle pose before ICP of the same pair (BASELINE config 2: syn.room_pair(0, 250k raw points), '
                       f'N0={int(g["n0"])}, N1={int(g["n1"])} voxels), fixture tests/golden/fullsize_config2.npz',
            'config': WORKLOAD}
  except Exception as e:   # noqa: BLE001
    return {'te_m': None, 're_rad': None, 'error': repr(e)}


def run_reference(args):
  """The reference's CPU implementation of the path (the oracle port: MinkowskiEngine is not installable
  offline, nothing of the reference compiles into oracle/_ref) on the SAME configuration as the B200 arm:
  full-size pairs of the same generator and seeds.  One such pair is minutes of CPU, so the run executes as
  many of the K steps as fit REF_TIME_BUDGET_S after the first (at least one) and says how many
  (cpu_baseline.steps_executed); pairs/s is per executed full-size pair, nothing is extrapolated."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  from oracle import pipeline as op
  state = syn.make_checkpoint(0)
  cores = cpu_threads()
  log(f'[bench] reference arm: {cores} threads (nproc {os.cpu_count()}, usable {effective_cpus()})')
  # warm-up: thread pools / allocator on a tiny sample of the same generator (seconds), untimed
  t_warm, _ = cpu_sample_time(state, 100, n_raw=REF_MIN_N_RAW)
  log(f'[bench] reference arm: warm-up sample {t_warm:.1f} s; timing full-size pairs')
  times, info, parity = [], {}, None
  t_begin = time.perf_counter()
  for i in range(args.steps):
    xyz0, xyz1, _ = syn.room_pair(1000 * rank + (i % POOL), n_raw=N_RAW)
    t = time.perf_counter()
    T, taps = op.register(state, xyz0, xyz1)
    times.append(time.perf_counter() - t)
    info = {'n0': int(len(taps['coords0'])), 'n1': int(len(taps['coords1'])), 'branch': taps['branch']}
    if i == 0:
      parity = fixture_parity(T)
    log(f'[bench] reference arm: pair {i} N0={info["n0"]} N1={info["n1"]} {times[-1]:.1f} s')
    if time.perf_counter() - t_begin + times[-1] > REF_TIME_BUDGET_S:
      break
  n_exec = len(times)
  dt = float(sum(times))
  val = n_exec / dt
  line = {'impl': 'reference', 'metric': 'scan_pairs_per_sec', 'value': val, 'unit': 'pairs/s',
          'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
          'ms_per_step': 1e3 * dt / n_exec, 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': base_config(args.gpus),
          'cpu_baseline': {'value': val, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                           'sample': f'{n_exec} FULL-SIZE pair(s) of the workload (same generator and seeds as the B200 arm: '
                                     f'{N_RAW} raw points per scan -> N0={info["n0"]}, N1={info["n1"]} voxels), not a reduced sample; '
                                     'CPU path = oracle port (torch-CPU index_select/mm/index_add per kernel offset, the algorithm '
                                     "of MinkowskiEngine's CPU backend) + restated kNN / Procrustes / Adam refinement",
                           'steps_executed': n_exec, 'seconds_per_pair': times,
                           'warmup_executed': f'1 pair of {REF_MIN_N_RAW} raw points ({t_warm:.1f} s)'},
          'workload_detail': info, 'parity': parity,
          'e2e': {'value': val, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
  _emit(json.dumps(line))


# ------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
  import torch.distributed as dist
  from deepglobalregistration_b200 import _abi, sharding
  from deepglobalregistration_b200.core.deep_global_registration import DeepGlobalRegistration

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)

  state = syn.make_checkpoint(0)
  cfg = types.SimpleNamespace(weights=state, clip_weight_thresh=0.05, verbose=False)
  dgr = DeepGlobalRegistration(cfg, device=dev)
  dgr.use_icp = False     # the benchmarked unit is tap A (through the refinement), as on the CPU arm

  # this rank's pairs (seeds disjoint across ranks): host copies for e2e, device copies for `value`
  pairs_host = [syn.room_pair(1000 * rank + i, n_raw=N_RAW) for i in range(POOL)]
  pairs_dev = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b, _ in pairs_host]

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def gather_poses(rows):
    # the path's only collective: one all-gather of [pairs, 20] results (NCCL over NVLink)
    return sharding.gather_results(rows, world * len(rows), device=dev)

  def timed(n_steps, host_inputs, profile=False):
    """K steps bracketed by barrier + synchronize; device time by CUDA events."""
    barrier()
    if profile:
      _abi.CONV_PROFILE = []
    l0, d0 = _abi.lib().dgr_launch_count(), _abi.D2H_BYTES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    poses, marks = [], []
    for s in range(n_steps):
      a, b = (pairs_host[s % POOL][:2] if host_inputs else pairs_dev[s % POOL])
      T = dgr.register(a, b)
      poses.append(sharding.pack_result(T, dgr.last_info.get('wsum', 0.0), dgr.last_info.get('iterations', 0),
                                        dgr.last_branch))
      ev = torch.cuda.Event(enable_timing=True)
      ev.record()
      marks.append(ev)
    gathered = gather_poses(poses)
    e1.record()
    barrier()
    wall = time.perf_counter() - w0
    ms = e0.elapsed_time(e1)
    prof, _abi.CONV_PROFILE = _abi.CONV_PROFILE, None
    launches = _abi.lib().dgr_launch_count() - l0
    d2h = _abi.D2H_BYTES - d0
    if world > 1:
      tm = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
      dist.all_reduce(tm, op=dist.ReduceOp.MAX)
      ms, wall = float(tm[0]), float(tm[1]) / 1e3
    steps_ms = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    return dict(ms=ms, wall=wall, launches=launches, d2h=d2h, prof=prof, poses=gathered, steps_ms=steps_ms)

  # started before the warm-up: nvidia-smi's own start-up (NVML init) must not land in the timed region
  sampler = ClockSampler(local) if rank == 0 and not os.environ.get('DGR_BENCH_NO_SAMPLER') else None
  log(f'[bench] rank {rank}/{world}: model + {POOL} pairs ready, warming up')
  # warm-up: every pair of the pool at least max(W, 3) times on both input paths, so that the
  # caching allocator has seen every buffer size before anything is timed
  n_warm = max(args.warmup, 3) * POOL
  timed(n_warm, host_inputs=False)
  timed(POOL, host_inputs=True)
  # one untimed rehearsal of exactly the timed loops: flushes every lazy initialisation
  # (allocator size classes, NCCL channels, pinned staging) out of the measurement
  timed(args.steps, host_inputs=False)
  timed(args.steps, host_inputs=True)
  gc.collect()
  log('[bench] warm-up done, timing')

  t_start = time.time()
  thr0 = cgroup_throttled_ms()
  ms0 = torch.cuda.memory_stats(dev)
  res = timed(args.steps, host_inputs=False, profile=True)     # `value`: scans resident in HBM
  t_mid = time.time()
  res_e2e = timed(args.steps, host_inputs=True)                 # `e2e`: host buffers in, pose out
  t_end = time.time()
  thr1 = cgroup_throttled_ms()
  ms1 = torch.cuda.memory_stats(dev)
  alloc_delta = {k: int(ms1.get(k, 0) - ms0.get(k, 0)) for k in ('num_device_alloc', 'num_device_free',
                                                                  'num_alloc_retries', 'num_sync_all_streams')}
  clocks = sampler.stop(t_start, t_end) if sampler else None

  # supplementary: the literal reference call (use_icp = True, host arrays in, pose out).  Single-rank
  # only (a failure here must not strand peers in a collective) and never fatal to the contract line.
  e2e_icp = None
  if world == 1:
    try:
      dgr.use_icp = True
      timed(2 * POOL, host_inputs=True)
      r_icp = timed(args.steps, host_inputs=True)
      e2e_icp = {'value': args.steps / (r_icp['ms'] / 1e3), 'unit': 'pairs/s', 'ms_per_step': r_icp['ms'] / args.steps,
                 'step_ms_median': float(np.median(r_icp['steps_ms'])),
                 'icp_iterations_last_pair': dgr.last_info.get('icp_iterations'),
                 'note': 'register() exactly as the reference defaults it: Procrustes + refinement + '
                         'point-to-point ICP; outside the contract value, which is SURVEY 8(d)\'s unit'}
    except Exception as e:   # noqa: BLE001
      log(f'[bench] supplementary ICP timing failed: {e!r}')
    finally:
      dgr.use_icp = False

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  K = args.steps
  value = world * K / (res['ms'] / 1e3)
  e2e = world * K / (res_e2e['ms'] / 1e3)

  # ---- roofline of the dominant kernel (live CUDA events around every launch) ---------------
  peaks, peak_src = None, 'fallback (B200_PROFILING.md: 6650 GB/s, 1590 TFLOP/s bf16)'
  try:
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    peak_src = 'measured (MEASURED_PEAKS.json)'
  except Exception:   # noqa: BLE001
    pass
  hbm_peak = float(peaks['hbm_gbs']) if peaks else 6650.0
  bf16_peak = float(peaks.get('bf16_tflops_sustained', peaks['bf16_tflops'])) if peaks else 1400.0
  by = {}
  for name, a, b, flops, nbytes in res['prof']:
    d = by.setdefault(name, [0, 0.0, 0.0, 0.0])
    d[0] += 1
    d[1] += a.elapsed_time(b)
    d[2] += flops
    d[3] += nbytes
  dom = max(by, key=lambda k: by[k][1]) if by else None
  roofline, roofline_tensor, kernel_share = None, None, None
  if dom:
    n, ms, flops, nbytes = by[dom]
    gbs = nbytes / (ms * 1e-3) / 1e9
    tfs = flops / (ms * 1e-3) / 1e12
    traffic = None
    try:   # per-launch DRAM bytes from the committed ncu --set full capture, if present
      traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r01_spconv_tc_traffic.json')))['dram_bytes_per_launch']
    except Exception:   # noqa: BLE001
      pass
    roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': gbs / hbm_peak, 'traffic': traffic, 'peak_source': peak_src,
                'launches_per_step': n / K, 'avg_launch_ms': ms / n,
                'algorithmic_bytes_per_launch': nbytes / n,
                'bytes_model': 'SURVEY 8(d) gather-scatter model: P*(Cin+Cout)*4 + 8*P + K_nonempty*Cin*Cout*4'}
    roofline_tensor = {'kernel': dom, 'bound': 'tensor', 'achieved': tfs, 'peak': bf16_peak, 'unit': 'TFLOP/s',
                       'frac': tfs / bf16_peak, 'algorithmic_flops_per_launch': flops / n,
                       'note': 'algorithmic fp32 FLOPs 2*P*Cin*Cout; the kernel spends 3 TF32 MMAs per product '
                               '(3xTF32) and TF32 runs at half the bf16 rate, so its ceiling is peak/6'}
    kernel_share = {k: v[1] / (res['ms']) for k, v in by.items()}

  log(f'[bench] value {value:.2f} pairs/s, e2e {e2e:.2f} pairs/s; timing the CPU sample')
  # ---- CPU baseline on a bounded sample --------------------------------------------------------
  cpu_s, info = cpu_sample_time(state, 0, reps=1)
  cpu = {'value': 1.0 / cpu_s, 'unit': 'pairs/s', 'cores': cpu_threads(), 'kind': 'port',
         'sample': sample_desc(info), 'seconds_per_sample_pair': cpu_s}
  try:     # the same-config CPU time: one FULL-SIZE oracle run of this arm's first pair, recorded with the fixture
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'fullsize_config2.npz'))
    sec = json.loads(str(g['seconds']))
    cpu['full_size_pair'] = {'seconds_through_refine': sec['total_through_refine'], 'threads': sec['threads'],
                             'pairs_per_sec': 1.0 / sec['total_through_refine'],
                             'where': 'build container, recorded by tests/golden/make_golden_fullsize.py; '
                                      '`bench.py --impl reference` times the same full-size pair on this box'}
  except Exception:   # noqa: BLE001
    pass

  # pose of this arm's first pair (timed `value` loop, step 0) against the CPU oracle's pose of the same pair
  parity = fixture_parity(res['poses'][0][:16].numpy()) if rank == 0 else None
  detail = dict(n0=dgr.last_info.get('n0'), n1=dgr.last_info.get('n1'), branch=dgr.last_branch,
                refine_iterations=dgr.last_info.get('iterations'))
  cfg_out = base_config(world)
  h2d = int(sum(a.nbytes + b.nbytes for a, b, _ in pairs_host) / POOL)
  fixed_d2h = 2 * 8 + 8 + 9 * 4 + 8 + 64     # counts, spec flags, coarse-map sizes, wsum, pose
  line = {'metric': 'scan_pairs_per_sec', 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': K,
          'warmup': args.warmup, 'warmup_steps_run': max(args.warmup, 3) * POOL + POOL + 2 * K, 'ms_per_step': res['ms'] / K, 'higher_is_better': True,
          'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg_out,
          'e2e': {'value': e2e, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d,
                  'd2h_bytes_per_step': int(res_e2e['d2h'] / K) + fixed_d2h,
                  'ms_per_step': res_e2e['ms'] / K, 'wall_ms_per_step': 1e3 * res_e2e['wall'] / K},
          'gpu_launches': int(res['launches']), 'gpu_launches_per_step': res['launches'] / K,
          'clocks': clocks, 'roofline': roofline, 'roofline_tensor': roofline_tensor,
          'kernel_share_of_step': kernel_share, 'cpu_baseline': cpu, 'parity': parity, 'workload_detail': detail,
          'wall_ms_per_step': 1e3 * res['wall'] / K,
          'host_cgroup_throttled_ms_during_timing': (thr1 - thr0) if thr0 is not None and thr1 is not None else None,
          'cuda_allocator_events_during_timing': alloc_delta,
          'host_threads': {'torch_intraop': torch.get_num_threads(), 'usable_cpus': effective_cpus()},
          'step_ms': {'min': min(res['steps_ms']), 'median': float(np.median(res['steps_ms'])),
                      'max': max(res['steps_ms']), 'all': [round(x, 2) for x in res['steps_ms']]},
          'e2e_step_ms': {'min': min(res_e2e['steps_ms']), 'median': float(np.median(res_e2e['steps_ms'])),
                          'max': max(res_e2e['steps_ms']), 'all': [round(x, 2) for x in res_e2e['steps_ms']]},
          'e2e_with_icp': e2e_icp,
          'published_reference': '0.69 s/pair without safeguard+ICP (reference assets/results.npz, unknown GPU)'}
  _emit(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  # keep stdout clean for the ONE JSON line: libraries (NCCL's version banner, ...) write to fd 1
  real_stdout = os.dup(1)
  os.dup2(2, 1)
  global _emit
  _emit = lambda text: os.write(real_stdout, (text + '\n').encode())
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=None,
                  help='timed steps (default 100 for the B200 arm: ~1.5 s per region, so that one ~0.3 s host stall '
                       'of a shared box costs 10-20 %% instead of halving the number; 10 for --impl reference)')
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  args = ap.parse_args()
  if args.steps is None:
    args.steps = 10 if args.impl == 'reference' else 100
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
