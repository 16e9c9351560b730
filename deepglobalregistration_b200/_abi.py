"""ctypes binding of libdgr_b200.so (include/dgr_b200.h).

torch is used here for device memory and the current CUDA stream only; every
computation happens inside the library.  There is no CPU fallback: importing this
module without a built library, or calling into it without an sm_100 device,
raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdgr_b200.so')

MAX_COLS = 8
TILE_ROWS = 128


class KeySpec(C.Structure):
  _fields_ = [('ncols', C.c_int32), ('overflow', C.c_int32), ('lo', C.c_int32 * MAX_COLS),
              ('shift', C.c_int32 * MAX_COLS), ('bits', C.c_int32 * MAX_COLS)]


KEYSPEC_INTS = C.sizeof(KeySpec) // 4

_p, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> argtypes; every function returns int32 status unless listed in _RESTYPES
SIGNATURES = {
    'dgr_version': [],
    'dgr_last_error': [],
    'dgr_launch_count': [],
    'dgr_device_check': [_i32],
    'dgr_quantize_points': [_p, _i32, _i64, _f64, _i32, _p, _p, _p],
    'dgr_coords_minmax': [_p, _i64, _i32, _p, _p],
    'dgr_keyspec_build': [_p, _i32, _i32, _p, _p],
    'dgr_hash_clear': [_p, _p, _i64, _p],
    'dgr_unique_first': [_p, _i64, _i32, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p],
    'dgr_scan_ws_elems': [_i64],
    'dgr_hash_find': [_p, _i64, _i32, _p, _p, _p, _i64, _p, _p],
    'dgr_gather_rows_i32': [_p, _p, _i64, _i32, _p, _p],
    'dgr_stride_coords': [_p, _i64, _i32, _i32, _p, _p],
    'dgr_bloom_build': [_p, _i64, _p, _i64, _p],
    'dgr_kernel_map_table': [_p, _i64, _i32, _p, _p, _p, _i64, _p, _i64, _p, _i32, _p, _p, _p],
    'dgr_kmap_ws_elems': [_i32, _i64],
    'dgr_kernel_map_count': [_p, _i32, _i64, _p, _i32, _p, _p, _p],
    'dgr_kernel_map_fill': [_p, _i32, _i64, _p, _p, _p, _p],
    'dgr_kernel_map_tiles': [_p, _i32, _i32, _i32, _i32, _p, _p, _p],
    'dgr_kernel_map_tiles2': [_p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p],
    'dgr_spconv_fwd': [_p, _i32, _p, _i32, _p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p],
    'dgr_spconv_tc_supported': [_i32, _i32],
    'dgr_pack_weight_tf32': [_p, _i32, _i32, _i32, _p, _p],
    'dgr_spconv_tc_fwd': [_p, _i32, _p, _i32, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p],
    'dgr_spconv_table_fwd': [_p, _i32, _p, _i32, _p, _i32, _i64, _p, _p, _p, _p],
    'dgr_linear_fwd': [_p, _i32, _p, _i32, _i64, _p, _i32, _p, _i32, _i32, _p, _p],
    'dgr_affine_act': [_p, _i64, _i32, _p, _p, _p, _i32, _p, _p],
    'dgr_cat2': [_p, _i32, _p, _i32, _i64, _p, _p],
    'dgr_l2_normalize': [_p, _i64, _i32, _p, _p],
    'dgr_knn_top1': [_p, _i64, _p, _i64, _i32, _p, _p, _p, _p],
    'dgr_knn_tc_supported': [_i32],
    'dgr_knn_tc_ws_elems': [_i64, _i64],
    'dgr_knn_top1_tc': [_p, _i64, _p, _i64, _i32, _p, _p, _p, _p, _p],
    'dgr_inlier_coords': [_p, _p, _p, _i64, _p, _p],
    'dgr_sigmoid_clip_sum': [_p, _i64, _f32, _p, _p, _p],
    'dgr_icp_point_to_point': [_p, _i64, _p, _p, _p, _p, _i64, _i32, _f64, _f64, _p, _i32, _f64, _f64, _p, _p, _p],
    'dgr_ransac_ws_elems': [_i64, _i64, _p],
    'dgr_ransac_correspondence': [_p, _p, _p, _p, _i64, _f64, _i64, C.c_uint64, _p, _p, _p],
    'dgr_se3_register': [_p, _p, _p, _p, _i64, _f32, _i32, _i32, _f32, _f32, _f32, _p, _p, _p, _p],
    # ---- round 2: coordinate planning with device-side counts (csrc/coordplan.cu) ----
    'dgr_spconv_table_fwd_strided': [_p, _i32, _p, _i32, _p, _i32, _i64, _i64, _p, _p, _p, _p],
    'dgr_compact_voxel_pair': [_p, _p, _p, _i64, _i64, _p, _i32, _p, _i32, _p, _p, _p, _p],
    'dgr_table_build_unique': [_p, _i64, _p, _i32, _p, _p, _p, _i64, _p],
    'dgr_coarse_scan_elems': [_i64],
    'dgr_coarse_maps': [_p, _i64, _p, _i32, _p, _i32, _p, _p, _p, _i64, _p, _p, _p, _p, _p],
    'dgr_bloom2_build': [_p, _i64, _p, _i64, _p],
    'dgr_kmap_mask_words': [_i64],
    'dgr_kmap_cnt_elems': [_i32, _i64],
    'dgr_kmap_probe': [_p, _i64, _p, _i32, _p, _p, _p, _i64, _p, _i64, _p, _i32, _p, _p, _p, _p, _p],
    'dgr_kmap_fill': [_p, _p, _i32, _i64, _p, _i32, _p, _p, _p, _i64, _p, _p, _p, _p],
    'dgr_kmap_dense': [_p, _i64, _p, _i32, _p, _p, _p, _i64, _p, _i64, _p, _i32, _p, _i64, _p, _p],
    'dgr_spconv_ones_bits_fwd': [_p, _i32, _p, _i64, _i32, _i64, _p, _p, _p, _p],
    'dgr_spconv_os_supported': [_i32, _i32],
    'dgr_spconv_os_fwd': [_p, _i32, _p, _i32, _p, _i64, _i32, _i64, _p, _p, _p, _i32, _p, _p],
    'dgr_spconv_wgrad': [_p, _i32, _p, _i32, _p, _p, _p, _i32, _p, _p],
    'dgr_affine_act_amax': [_p, _i64, _i32, _p, _p, _p, _i32, _p, _p, _p],
    'dgr_absmax_f32': [_p, _i64, _p, _p],
    'dgr_spconv_tc_f16_supported': [_i32, _i32],
    'dgr_pack_weight_f16': [_p, _i32, _i32, _i32, _p, _p, _p],
    'dgr_spconv_tc_f16_fwd': [_p, _i32, _p, _i32, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p],
    # ---- round 2: native executor (csrc/exec.cu) ----
    'dgr_ctx_create': [_i32, _p, _p],
    'dgr_ctx_destroy': [_p],
    'dgr_ctx_stream': [_p],
    'dgr_ctx_stats': [_p, _p],
    'dgr_ctx_profile': [_p, _i32],
    'dgr_ctx_profile_read': [_p, _p, _i64],
    'dgr_ctx_stage_times': [_p, _p, _i32],
    'dgr_net_create': [_i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _p],
    'dgr_net_destroy': [_p],
    'dgr_net_forward': [_p, _p, _p, _i64, _p, _p],
    'dgr_pair_register': [_p, _p, _p, _p, _i64, _i32, _p, _i64, _i32, _i32, _f64, _f32, _i32, _p],
    'dgr_pair_safeguard': [_p, _f64, _i64, C.c_uint64, _i32, _p],
    'dgr_pair_tap': [_p, _i32, _p, _p, _p],
}
_RESTYPES = {'dgr_coarse_scan_elems': _i64, 'dgr_kmap_mask_words': _i64, 'dgr_kmap_cnt_elems': _i64, 'dgr_ctx_stream': C.c_void_p,
             'dgr_ctx_profile_read': _i64, 'dgr_last_error': C.c_char_p, 'dgr_knn_tc_ws_elems': _i64, 'dgr_launch_count': _i64, 'dgr_spconv_tc_supported': _i32, 'dgr_scan_ws_elems': _i64, 'dgr_kmap_ws_elems': _i64}

_lib = None


class DgrError(RuntimeError):
  pass


def lib():
  """Load the shared library (once).  Fails loudly when it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise DgrError(f'{LIB_PATH} is missing: run `python -m deepglobalregistration_b200.build` '
                     '(there is no CPU fallback)')
    l = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
      fn = getattr(l, name)
      fn.argtypes = args
      fn.restype = _RESTYPES.get(name, _i32)
    _lib = l
  return _lib


def _check(status, name):
  if status != 0:
    raise DgrError(f'{name} failed ({status}): {lib().dgr_last_error().decode()}')


_FN = {}

# When set to a list, every ABI call is bracketed by CUDA events on torch's current stream and appends
# (entry point, start event, end event): tools/abi_profile.py turns that into warm, in-context GPU time
# per entry point (ncu's per-launch times are cold-cache and serialised).  None = no overhead.
CALL_PROFILE = None


def call(name, *args):
  fn = _FN.get(name)
  if fn is None:
    fn = _FN[name] = getattr(lib(), name)
  if CALL_PROFILE is not None:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    status = fn(*args)
    e1.record()
    CALL_PROFILE.append((name, e0, e1))
  else:
    status = fn(*args)
  if status != 0:
    _check(status, name)


_checked_devices = set()


def require_device(device):
  device = torch.device(device)
  if device.type != 'cuda':
    raise DgrError(f'libdgr_b200 computes on CUDA (sm_100a) only, got device {device}; there is no '
                   'CPU fallback')
  idx = device.index if device.index is not None else torch.cuda.current_device()
  if idx not in _checked_devices:
    _check(lib().dgr_device_check(idx), 'dgr_device_check')
    _checked_devices.add(idx)
  return torch.device('cuda', idx)


def ptr(t):
  return 0 if t is None else t.data_ptr()


_STREAM = None


def stream():
  """Handle of the CUDA stream work is enqueued on.  torch.cuda.current_stream() costs ~12 us per
  call, so the handle is looked up once per public entry point (refresh_stream) and cached."""
  global _STREAM
  if _STREAM is None:
    _STREAM = torch.cuda.current_stream().cuda_stream
  return _STREAM


def refresh_stream():
  global _STREAM
  _STREAM = torch.cuda.current_stream().cuda_stream
  return _STREAM


def _chk(t, dtype, name):
  if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
    raise DgrError(f'{name}: expected contiguous CUDA {dtype}, got {t.dtype} on {t.device} '
                   f'(contiguous={t.is_contiguous()})')
  return t


def next_pow2(n):
  p = 1
  while p < n:
    p <<= 1
  return p


# --------------------------------------------------------------------------- #
# scratch arena: the big transient workspaces (dense neighbour tables, scan / pack buffers)
# are views of per-(device, stream) buffers that only ever grow, so steady-state calls
# never reach cudaMalloc/cudaFree whatever the cloud sizes are.  Work on one stream is
# ordered, which is what makes the reuse safe.
# --------------------------------------------------------------------------- #
_ARENA = {}


def scratch(name, numel, dtype, device):
  key = (name, dtype, device.index, stream())
  buf = _ARENA.get(key)
  if buf is None or buf.numel() < numel:
    buf = torch.empty(max(int(numel * 1.25), 1024), dtype=dtype, device=device)
    _ARENA[key] = buf
  return buf[:numel]


# --------------------------------------------------------------------------- #
# thin typed wrappers (allocate outputs / workspaces with torch, call the ABI)
# --------------------------------------------------------------------------- #
class HashTable:
  __slots__ = ('keys', 'vals', 'cap', '_bloom')

  def bloom(self):
    """(words, n_bits) miss filter, built on first use by a kernel map."""
    if getattr(self, '_bloom', None) is None:
      bits = 16 * self.cap
      words = torch.empty(bits // 32, dtype=torch.int32, device=self.keys.device)
      call('dgr_bloom_build', ptr(self.keys), self.cap, ptr(words), bits, stream())
      self._bloom = (words, bits)
    return self._bloom

  def __init__(self, n, device):
    self.cap = max(1024, next_pow2(2 * max(int(n), 1)))
    self.keys = torch.empty(self.cap, dtype=torch.int64, device=device)
    self.vals = torch.empty(self.cap, dtype=torch.int32, device=device)
    self._bloom = None
    call('dgr_hash_clear', ptr(self.keys), ptr(self.vals), self.cap, stream())


def quantize_points(xyz, voxel, batch=0):
  """xyz CUDA float64/float32 [n,3] -> (coords int32 [n,4], minmax int32 [8])."""
  assert xyz.is_cuda and xyz.is_contiguous() and xyz.shape[1] == 3
  is64 = {torch.float64: 1, torch.float32: 0}[xyz.dtype]
  n = xyz.shape[0]
  coords = torch.empty(n, 4, dtype=torch.int32, device=xyz.device)
  minmax = torch.empty(8, dtype=torch.int32, device=xyz.device)
  call('dgr_quantize_points', ptr(xyz), is64, n, float(voxel), int(batch), ptr(coords), ptr(minmax),
       stream())
  return coords, minmax


def coords_minmax(coords):
  _chk(coords, torch.int32, 'coords')
  minmax = torch.empty(2 * coords.shape[1], dtype=torch.int32, device=coords.device)
  call('dgr_coords_minmax', ptr(coords), coords.shape[0], coords.shape[1], ptr(minmax), stream())
  return minmax


def keyspec_build(minmax, ncols, margin=32):
  spec = torch.empty(KEYSPEC_INTS, dtype=torch.int32, device=minmax.device)
  call('dgr_keyspec_build', ptr(minmax), ncols, margin, ptr(spec), stream())
  return spec


def unique_first(coords, spec):
  """-> (table, sel int32 [n] (first m valid), inverse int32 [n], cnt device int32 [2] =
  (n_unique, key-overflow flag))."""
  _chk(coords, torch.int32, 'coords')
  n, ncols = coords.shape
  dev = coords.device
  table = HashTable(n, dev)
  sel = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
  inverse = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
  cnt = torch.zeros(2, dtype=torch.int32, device=dev)
  slot = scratch('uf_slot', max(n, 1), torch.int32, dev)
  rank = scratch('uf_rank', max(n, 1), torch.int32, dev)
  scan = scratch('uf_scan', lib().dgr_scan_ws_elems(n), torch.int32, dev)
  call('dgr_unique_first', ptr(coords), n, ncols, ptr(spec), ptr(table.keys), ptr(table.vals), table.cap,
       ptr(sel), ptr(inverse), ptr(cnt), ptr(slot), ptr(rank), ptr(scan), stream())
  return table, sel, inverse, cnt


def read_count(cnt):
  """Host read of unique_first's (count, overflow) pair; raises on key overflow."""
  global D2H_BYTES
  n, overflow = cnt.cpu().tolist()
  D2H_BYTES += 8
  if overflow:
    raise DgrError('coordinate extent does not fit a 63-bit packed key')
  return n


def hash_find(coords, spec, table):
  _chk(coords, torch.int32, 'coords')
  rows = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
  call('dgr_hash_find', ptr(coords), coords.shape[0], coords.shape[1], ptr(spec), ptr(table.keys),
       ptr(table.vals), table.cap, ptr(rows), stream())
  return rows


def gather_rows_i32(src, idx, n):
  out = torch.empty(n, src.shape[1], dtype=torch.int32, device=src.device)
  call('dgr_gather_rows_i32', ptr(src), ptr(idx), n, src.shape[1], ptr(out), stream())
  return out


def stride_coords(coords, out_stride):
  out = torch.empty_like(coords)
  call('dgr_stride_coords', ptr(coords), coords.shape[0], coords.shape[1], out_stride, ptr(out), stream())
  return out


class KernelMap:
  """Neighbour table + (kappa, j)-sorted pair lists + the gather-GEMM-scatter work list."""
  __slots__ = ('K', 'n_in', 'n_out', 'nbr', 'in_idx', 'out_idx', 'kofs', 'kofs_host', 'n_pairs',
               'tile_k', 'tile_start', 'n_tiles', '_paired')

  def paired_tiles(self):
    """(tile_k, tile_start, n_tiles) with an even tile count per offset (2-CTA cluster kernel)."""
    if getattr(self, '_paired', None) is None:
      counts = self.kofs_host[1:] - self.kofs_host[:-1]
      per_k = (counts + TILE_ROWS - 1) // TILE_ROWS
      n = int(((per_k + 1) // 2 * 2).sum())
      dev = self.kofs.device
      tk = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
      ts = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
      call('dgr_kernel_map_tiles', ptr(self.kofs), self.K, TILE_ROWS, n, 1, ptr(tk), ptr(ts), stream())
      self._paired = (tk, ts, n)
    return self._paired

  def transposed(self):
    t = KernelMap()
    for k in self.__slots__:
      setattr(t, k, getattr(self, k, None))
    t.in_idx, t.out_idx = self.out_idx, self.in_idx
    t.n_in, t.n_out = self.n_out, self.n_in
    t.nbr = None
    return t


def kernel_map_begin(out_coords, spec, in_table, n_in, offsets, keep_table=False, slot=0):
  """First half of a kernel map: neighbour table + bucket counts, all asynchronous.  `slot` selects
  the scratch buffers so that several maps can be in flight before ONE host read finishes them
  all (kernel_maps_finish)."""
  _chk(out_coords, torch.int32, 'out_coords')
  _chk(offsets, torch.int32, 'offsets')
  dev = out_coords.device
  n_out, ncols = out_coords.shape
  K = offsets.shape[0]
  if keep_table:
    nbr = torch.empty(K, max(n_out, 1), dtype=torch.int32, device=dev)
  else:
    nbr = scratch(('km_nbr', slot), K * max(n_out, 1), torch.int32, dev).view(K, max(n_out, 1))
  # the miss filter pays off when most probes miss: many offsets per row (6-D, 5^3, 7^3 kernels)
  bloom, bloom_bits = in_table.bloom() if K > 27 else (None, 0)
  ws = scratch(('km_ws', slot), lib().dgr_kmap_ws_elems(K, n_out), torch.int32, dev)
  call('dgr_kernel_map_table', ptr(out_coords), n_out, ncols, ptr(spec), ptr(in_table.keys),
       ptr(in_table.vals), in_table.cap, ptr(bloom), bloom_bits, ptr(offsets), K, ptr(nbr), ptr(ws), stream())
  kofs = torch.empty(K + 2, dtype=torch.int32, device=dev)
  call('dgr_kernel_map_count', ptr(nbr), K, n_out, ptr(ws), 1, ptr(kofs), ptr(spec), stream())
  return dict(K=K, n_in=n_in, n_out=n_out, nbr=nbr, ws=ws, kofs=kofs, keep=keep_table, dev=dev)


def kernel_map_finish(pend, kofs_all):
  """Second half: pair lists and work list, sized from the host copy of the bucket offsets."""
  global D2H_BYTES
  K, n_out, dev = pend['K'], pend['n_out'], pend['dev']
  D2H_BYTES += kofs_all.nbytes
  if kofs_all[K + 1] != 0:
    raise DgrError('coordinate extent does not fit a 63-bit packed key')
  kofs_host = kofs_all[:K + 1]
  P = int(kofs_host[K])
  counts = kofs_host[1:] - kofs_host[:-1]
  n_tiles = int(((counts + TILE_ROWS - 1) // TILE_ROWS).sum())
  km = KernelMap()
  km.K, km.n_in, km.n_out = K, pend['n_in'], n_out
  km.in_idx = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
  km.out_idx = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
  if n_out > 0:
    call('dgr_kernel_map_fill', ptr(pend['nbr']), K, n_out, ptr(pend['ws']), ptr(km.in_idx), ptr(km.out_idx),
         stream())
  km.tile_k = torch.empty(max(n_tiles, 1), dtype=torch.int32, device=dev)
  km.tile_start = torch.empty(max(n_tiles, 1), dtype=torch.int32, device=dev)
  call('dgr_kernel_map_tiles', ptr(pend['kofs']), K, TILE_ROWS, n_tiles, 0, ptr(km.tile_k), ptr(km.tile_start),
       stream())
  km.kofs, km.kofs_host, km.n_pairs, km.n_tiles = pend['kofs'], kofs_host, P, n_tiles
  km.nbr = pend['nbr'] if pend['keep'] else None
  km._paired = None
  return km


def kernel_maps_finish(pending):
  """Finish several begun kernel maps with a single device-to-host read."""
  if not pending:
    return []
  host = torch.cat([p['kofs'] for p in pending]).cpu().numpy() if len(pending) > 1 else \
      pending[0]['kofs'].cpu().numpy()
  out, ofs = [], 0
  for p in pending:
    n = p['K'] + 2
    out.append(kernel_map_finish(p, host[ofs:ofs + n]))
    ofs += n
  return out


def kernel_map(out_coords, spec, in_table, n_in, offsets, keep_table=False):
  """offsets: CUDA int32 [K, D] (scaled by the input tensor stride).  One host read."""
  return kernel_maps_finish([kernel_map_begin(out_coords, spec, in_table, n_in, offsets, keep_table)])[0]


def spconv_fwd(feat, weight, km, out, relu_in=False):
  """out[km.out_idx] += feat[km.in_idx] @ weight[kappa]; `out` holds the initial value."""
  _chk(feat, torch.float32, 'feat'); _chk(weight, torch.float32, 'weight'); _chk(out, torch.float32, 'out')
  cin, cout = feat.shape[1], out.shape[1]
  assert weight.numel() == km.K * cin * cout, (weight.shape, km.K, cin, cout)
  assert feat.shape[0] == km.n_in and out.shape[0] == km.n_out, (feat.shape, out.shape, km.n_in, km.n_out)
  _conv_profiled('spconv_fwd_kernel', km, cin, cout, lambda: call(
      'dgr_spconv_fwd', ptr(feat), cin, ptr(weight), cout, ptr(km.in_idx), ptr(km.out_idx), ptr(km.kofs),
      ptr(km.tile_k), ptr(km.tile_start), km.n_tiles, TILE_ROWS, int(relu_in), ptr(out), stream()))
  return out


def spconv_wgrad(feat, grad_out, km):
  """Weight gradient [K, cin, cout] of the convolution over `km` (training)."""
  _chk(feat, torch.float32, 'feat'); _chk(grad_out, torch.float32, 'grad_out')
  cin, cout = feat.shape[1], grad_out.shape[1]
  dw = torch.empty(km.K, cin, cout, dtype=torch.float32, device=feat.device)
  call('dgr_spconv_wgrad', ptr(feat), cin, ptr(grad_out), cout, ptr(km.in_idx), ptr(km.out_idx), ptr(km.kofs), km.K,
       ptr(dw), stream())
  return dw


def tc_supported(cin, cout):
  return bool(lib().dgr_spconv_tc_supported(int(cin), int(cout)))


def pack_weight_tf32(weight, K, cin, cout):
  """[K, cin, cout] -> packed TF32 hi/lo slabs [K, cin/32, 2, cout, 32] (shared-memory image
  order) for the tensor-core convolution."""
  _chk(weight, torch.float32, 'weight')
  packed = torch.empty(K, cin // 32, 2, cout, 32, dtype=torch.float32, device=weight.device)
  call('dgr_pack_weight_tf32', ptr(weight), K, cin, cout, ptr(packed), stream())
  return packed


# kernel variant of the tensor-core convolution: 1 = one CTA per tile, both operands in shared
# memory (default), 0 = A operand in tensor memory, 2 = CTA pairs with multicast weight tiles,
# 3 = cta_group::2 (one M = 256 MMA per tile pair, half of every weight tile per CTA; used for
# cout >= TC_PAIR_MIN_COUT, else 1).  All four are parity-tested.  With the line-coalesced epilogue
# variant 3 is ~6 % faster on the wide layers (profiles/r01_spconv_tc_experiments.txt); default since
# round 2 (the whole GPU suite runs with it).
TC_VARIANT = int(os.environ.get('DGR_TC_VARIANT', '3'))
TC_PAIR_MIN_COUT = int(os.environ.get('DGR_TC_PAIR_MIN_COUT', '128'))

# When set to a list, every sparse-convolution launch appends
# (kernel name, start event, end event, algorithmic flops, gather-scatter-model bytes):
# bench.py's live per-kernel roofline measurement (CUDA events on the launching stream).
CONV_PROFILE = None
D2H_BYTES = 0          # bytes this module has read back to the host (counts, bucket offsets, results)


def _conv_profiled(name, km, cin, cout, fn):
  if CONV_PROFILE is None:
    return fn()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  fn()
  e1.record()
  nonempty = int((km.kofs_host[1:] > km.kofs_host[:-1]).sum())
  flops = 2.0 * km.n_pairs * cin * cout
  # gather read + scatter write + (in, out) index pair + weights of non-empty offsets
  nbytes = km.n_pairs * (cin + cout) * 4.0 + 8.0 * km.n_pairs + nonempty * cin * cout * 4.0
  CONV_PROFILE.append((name, e0, e1, flops, nbytes))


def spconv_tc_fwd(feat, weight_t, km, out, passes=3, cluster=None):
  """Tensor-core gather-GEMM-scatter: out[km.out_idx] += feat[km.in_idx] @ W[kappa]."""
  _chk(feat, torch.float32, 'feat'); _chk(weight_t, torch.float32, 'weight_t'); _chk(out, torch.float32, 'out')
  cin, cout = feat.shape[1], out.shape[1]
  assert weight_t.numel() == 2 * km.K * cin * cout
  assert feat.shape[0] == km.n_in and out.shape[0] == km.n_out
  if cluster is None:
    cluster = TC_VARIANT
    if cluster == 3 and cout < TC_PAIR_MIN_COUT:
      cluster = 1               # narrow layers are bound by the gather, not by weight ingress: no pairing
  if cluster in (2, 3):
    tk, ts, nt = km.paired_tiles()
  else:
    tk, ts, nt = km.tile_k, km.tile_start, km.n_tiles
  _conv_profiled('spconv_tc_kernel', km, cin, cout, lambda: call(
      'dgr_spconv_tc_fwd', ptr(feat), cin, ptr(weight_t), cout, ptr(km.in_idx), ptr(km.out_idx),
      ptr(km.kofs), ptr(tk), ptr(ts), nt, TILE_ROWS, int(passes), cluster, ptr(out), stream()))
  return out


def spconv_tc_f16_fwd(feat, weight, km, out, amax=None):
  """3xFP16 cta_group::2 convolution (weight: the fp32 [K, cin, cout] kernel; packed per call - test helper)."""
  _chk(feat, torch.float32, 'feat'); _chk(weight, torch.float32, 'weight'); _chk(out, torch.float32, 'out')
  cin, cout = feat.shape[1], out.shape[1]
  packed = torch.empty(4 * km.K * cin * cout, dtype=torch.uint8, device=feat.device)
  wscale = torch.empty(2, dtype=torch.float32, device=feat.device)
  call('dgr_pack_weight_f16', ptr(weight), km.K, cin, cout, ptr(packed), ptr(wscale), stream())
  if amax is None:
    amax = torch.empty(1, dtype=torch.float32, device=feat.device)
    call('dgr_absmax_f32', ptr(feat), feat.numel(), ptr(amax), stream())
  tk, ts, nt = km.paired_tiles()
  call('dgr_spconv_tc_f16_fwd', ptr(feat), cin, ptr(packed), cout, ptr(km.in_idx), ptr(km.out_idx), ptr(km.kofs),
       ptr(tk), ptr(ts), nt, TILE_ROWS, ptr(amax), ptr(wscale), ptr(out), stream())
  return out


def spconv_os_fwd(feat, weight_t, nbr, cout, scale=None, shift=None, residual=None, relu=False):
  """Output-stationary tensor-core convolution with the fused epilogue; nbr: dense table [K, n_out] int32."""
  _chk(feat, torch.float32, 'feat'); _chk(weight_t, torch.float32, 'weight_t'); _chk(nbr, torch.int32, 'nbr')
  K, n_out = nbr.shape
  out = torch.empty(n_out, cout, dtype=torch.float32, device=feat.device)
  call('dgr_spconv_os_fwd', ptr(feat), feat.shape[1], ptr(weight_t), cout, ptr(nbr), nbr.stride(0), K, n_out,
       ptr(scale), ptr(shift), ptr(residual), int(relu), ptr(out), stream())
  return out


def spconv_table_fwd(feat, weight, km, cout, scale=None, shift=None):
  _chk(feat, torch.float32, 'feat'); _chk(weight, torch.float32, 'weight')
  assert km.nbr is not None
  out = torch.empty(km.n_out, cout, dtype=torch.float32, device=feat.device)
  call('dgr_spconv_table_fwd', ptr(feat), feat.shape[1], ptr(weight), cout, ptr(km.nbr), km.K, km.n_out,
       ptr(scale), ptr(shift), ptr(out), stream())
  return out


def linear_fwd(a, weight, bias=None, b=None, relu=False, normalize=False):
  _chk(a, torch.float32, 'a'); _chk(weight, torch.float32, 'weight')
  n, ca = a.shape
  cb = 0 if b is None else b.shape[1]
  cout = weight.numel() // (ca + cb)
  out = torch.empty(n, cout, dtype=torch.float32, device=a.device)
  call('dgr_linear_fwd', ptr(a), ca, ptr(b), cb, n, ptr(weight), cout, ptr(bias), int(relu), int(normalize),
       ptr(out), stream())
  return out


def affine_act(x, scale=None, shift=None, residual=None, relu=False, out=None):
  _chk(x, torch.float32, 'x')
  if out is None:
    out = torch.empty_like(x)
  call('dgr_affine_act', ptr(x), x.shape[0], x.shape[1], ptr(scale), ptr(shift), ptr(residual), int(relu),
       ptr(out), stream())
  return out


def cat2(a, b):
  _chk(a, torch.float32, 'a'); _chk(b, torch.float32, 'b')
  out = torch.empty(a.shape[0], a.shape[1] + b.shape[1], dtype=torch.float32, device=a.device)
  call('dgr_cat2', ptr(a), a.shape[1], ptr(b), b.shape[1], a.shape[0], ptr(out), stream())
  return out


def l2_normalize(x):
  _chk(x, torch.float32, 'x')
  out = torch.empty_like(x)
  call('dgr_l2_normalize', ptr(x), x.shape[0], x.shape[1], ptr(out), stream())
  return out


KNN_MODE = 'tc'      # 'tc': tensor-core pre-filter + exact candidates; 'simt': fp32 brute force


def knn_top1(f0, f1, return_distance=False, mode=None):
  """Top-1 neighbour of every f0 row in f1 (both modes return identical indices)."""
  _chk(f0, torch.float32, 'f0'); _chk(f1, torch.float32, 'f1')
  n0, n1, c = f0.shape[0], f1.shape[0], f0.shape[1]
  mode = mode or KNN_MODE
  ws = scratch('knn_ws', max(n0, 1), torch.int64, f0.device)
  idx = torch.empty(n0, dtype=torch.int32, device=f0.device)
  dist = torch.empty(n0, dtype=torch.float32, device=f0.device) if return_distance else None
  if mode == 'tc' and lib().dgr_knn_tc_supported(c) and n0 > 0:
    fws = scratch('knn_fws', lib().dgr_knn_tc_ws_elems(n0, n1), torch.float32, f0.device)
    call('dgr_knn_top1_tc', ptr(f0), n0, ptr(f1), n1, c, ptr(ws), ptr(fws), ptr(idx), ptr(dist), stream())
  else:
    call('dgr_knn_top1', ptr(f0), n0, ptr(f1), n1, c, ptr(ws), ptr(idx), ptr(dist), stream())
  return (idx, dist) if return_distance else idx


def inlier_coords(coords0, coords1, idx1):
  _chk(coords0, torch.int32, 'coords0'); _chk(coords1, torch.int32, 'coords1'); _chk(idx1, torch.int32, 'idx1')
  out = torch.empty(coords0.shape[0], 7, dtype=torch.int32, device=coords0.device)
  call('dgr_inlier_coords', ptr(coords0), ptr(coords1), ptr(idx1), coords0.shape[0], ptr(out), stream())
  return out


def sigmoid_clip_sum(logit, clip):
  _chk(logit, torch.float32, 'logit')
  w = torch.empty_like(logit)
  wsum = torch.empty(1, dtype=torch.float64, device=logit.device)
  call('dgr_sigmoid_clip_sum', ptr(logit), logit.numel(), float(clip), ptr(w), ptr(wsum), stream())
  return w, wsum


def se3_register(x, y, w, idx1=None, quantization_size=1.0, max_iter=1000, max_break_count=20,
                 break_threshold_ratio=1e-4, lr=0.1, gamma=0.999):
  """-> device float32 [16]: R (9), t (3), iterations, loss, break_count, n_active."""
  _chk(x, torch.float32, 'x'); _chk(y, torch.float32, 'y'); _chk(w, torch.float32, 'w')
  n = x.shape[0]
  pack = scratch('se3_pack', 7 * n, torch.float32, x.device)
  cnt = torch.empty(4, dtype=torch.int32, device=x.device)
  res = torch.empty(16, dtype=torch.float32, device=x.device)
  call('dgr_se3_register', ptr(x), ptr(y), ptr(idx1), ptr(w), n, float(quantization_size), int(max_iter),
       int(max_break_count), float(break_threshold_ratio), float(lr), float(gamma), ptr(pack), ptr(cnt),
       ptr(res), stream())
  return res


def icp_point_to_point(src, tgt, tgt_manager, voxel, max_dist, T_init, max_iter=30, rel_fitness=1e-6,
                       rel_rmse=1e-6, batch=0):
  """Point-to-point ICP (open3d defaults) of src onto tgt through tgt's voxel hash.
  src / tgt: CUDA float32 [n, 3]; tgt_manager: the CoordinateManager preprocess() built for tgt;
  T_init: 4x4 (numpy / tensor) or device double [12].  -> device double [20]."""
  _chk(src, torch.float32, 'src'); _chk(tgt, torch.float32, 'tgt')
  dev = src.device
  m = tgt_manager._maps[1]
  if not (isinstance(T_init, torch.Tensor) and T_init.is_cuda and T_init.numel() == 12):
    T_init = torch.as_tensor(T_init, dtype=torch.float64).reshape(4, 4)[:3].contiguous().to(dev)
  state = torch.empty(64, dtype=torch.float64, device=dev)
  res = torch.empty(20, dtype=torch.float64, device=dev)
  call('dgr_icp_point_to_point', ptr(src), src.shape[0], ptr(tgt), ptr(tgt_manager.spec), ptr(m.table.keys),
       ptr(m.table.vals), m.table.cap, int(batch), float(voxel), float(max_dist), ptr(T_init), int(max_iter),
       float(rel_fitness), float(rel_rmse), ptr(state), ptr(res), stream())
  return res


def ransac_correspondence(x, y, idx0, idx1, max_dist, num_hyp=4000000, seed=0):
  """Safeguard RANSAC over the correspondences (x[idx0[i]], y[idx1[i]]) (an index of None means
  arange).  x / y: CUDA float32 [n, 3]; idx: int32.  -> device double [20] (pose 16, fitness,
  inlier RMSE, winning hypothesis, its inlier count)."""
  _chk(x, torch.float32, 'x'); _chk(y, torch.float32, 'y')
  dev = x.device
  if idx0 is not None: _chk(idx0, torch.int32, 'idx0')
  if idx1 is not None: _chk(idx1, torch.int32, 'idx1')
  n = len(idx0) if idx0 is not None else (len(idx1) if idx1 is not None else x.shape[0])
  if idx0 is not None and idx1 is not None and len(idx0) != len(idx1):
    raise DgrError('idx0 and idx1 differ in length')
  words = C.c_int64(0)
  call('dgr_ransac_ws_elems', n, int(num_hyp), C.byref(words))
  ws = scratch('ransac', words.value, torch.int64, dev)
  res = torch.empty(20, dtype=torch.float64, device=dev)
  call('dgr_ransac_correspondence', ptr(x), ptr(y), ptr(idx0) if idx0 is not None else None,
       ptr(idx1) if idx1 is not None else None, n, float(max_dist), int(num_hyp), int(seed) & (2**64 - 1),
       ptr(ws), ptr(res), stream())
  return res
