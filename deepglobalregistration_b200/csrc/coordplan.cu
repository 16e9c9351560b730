// Coordinate planning with DEVICE-SIDE row counts: the integer work of one network forward
// (coarse coordinate maps, kernel maps) enqueued without a single host round trip.
//
// The round-1 builders (coords.cu) size every launch from host-side counts, so building the
// maps of one ResUNet cost one device-to-host read per dependent step (rows of each strided
// map, pairs of each kernel map).  Here every kernel takes (n_max, n_dev): a host-side upper
// bound that sizes buffers and grids, and a device pointer to the actual count that the
// kernels read.  The native executor (exec.cu) enqueues the whole coordinate phase of a
// network, then reads ONE small meta block (rows per level, pairs / tiles per map, key
// overflow) and launches the fill + convolution phase.
//
// Kernel maps no longer go through the dense neighbour table nbr[K][N_out] (150 MB per 6-D
// map, 99.7 % of it -1): the probe pass stores one BIT per (offset, output row) - a ballot
// word per (offset, warp of 32 rows), 32x smaller - together with per-(offset, block) hit
// counts; after an exclusive scan the fill pass walks the set bits, re-probes those (hits
// only) and writes the (kappa, j)-sorted pair lists, bit-identical to the round-1 builder.
// Misses, 96-99.7 % of all probes of a 6-D map, are answered by a blocked Bloom filter of the
// input table held in SHARED memory (one 32-bit word holds both bits of a key: one
// shared-memory load per probe) instead of an L2 round trip.
//
// Replaces the MinkowskiEngine coordinate manager / kernel-map builder behind
// model/residual_block.py:31-80 and model/resunet.py:598-649 (reference call sites); the pair
// list semantics are those frozen in oracle/sparse_ops.py.
#include <limits.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kScanElems = 2048;      // elements per block in the flag scans of the coarse maps (256 threads x 8)
constexpr int kCntWords = 256;        // mask words per counted block of a kernel map (one fill block, one word per thread)
constexpr int kProbeThreads = 512;    // 16 warps = 512 output rows per probe block
constexpr int kMaxLevels = 4;

__device__ __forceinline__ int dev_count(const int32_t* n_dev, int64_t n_max) {
  return n_dev != nullptr ? *n_dev : (int)n_max;
}

__device__ __forceinline__ int floor_to(int v, int stride) {
  int q = v / stride;
  if ((v % stride != 0) && (v < 0)) --q;
  return q * stride;
}

// key of a row floored to `stride` on the spatial columns (column 0 = batch is kept)
__device__ __forceinline__ uint64_t pack_key_strided(const int32_t* __restrict__ row, const dgr_keyspec_t& s,
                                                     int stride) {
  uint64_t k = 0;
#pragma unroll
  for (int i = 0; i < DGR_MAX_COLS; ++i)
    if (i < s.ncols) {
      const int v = (i == 0 || stride == 1) ? row[i] : floor_to(row[i], stride);
      k += (uint64_t)(uint32_t)(v - s.lo[i]) << s.shift[i];
    }
  return k;
}

// ---------------------------------------------------------------------------------------
// voxel compaction of a scan pair (device count of kept points)
// ---------------------------------------------------------------------------------------
template <typename T0, typename T1>
__global__ void compact_voxels_kernel(const int32_t* __restrict__ raw, const int32_t* __restrict__ sel,
                                      const int32_t* __restrict__ n_unique, int64_t n_raw0,
                                      const T0* __restrict__ xyz0, const T1* __restrict__ xyz1,
                                      int32_t* __restrict__ coords, float* __restrict__ xyz,
                                      int32_t* __restrict__ counts) {
  const int n = n_unique[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // sel is ascending: rows of cloud 0 come first; N0 = lower_bound(sel, n_raw0)
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sel[mid] < n_raw0) lo = mid + 1; else hi = mid;
    }
    counts[0] = n;
    counts[1] = lo;
    counts[2] = n - lo;
    counts[3] = n_unique[1];     // key-overflow flag
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = sel[i];
    reinterpret_cast<int4*>(coords)[i] = reinterpret_cast<const int4*>(raw)[r];
    float x, y, z;
    if (r < n_raw0) {
      x = (float)xyz0[3 * r]; y = (float)xyz0[3 * r + 1]; z = (float)xyz0[3 * r + 2];
    } else {
      const int64_t q = r - n_raw0;
      x = (float)xyz1[3 * q]; y = (float)xyz1[3 * q + 1]; z = (float)xyz1[3 * q + 2];
    }
    xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
  }
}

// ---------------------------------------------------------------------------------------
// hash table of rows known to be distinct (value = row index)
// ---------------------------------------------------------------------------------------
__global__ void table_clear_kernel(uint64_t* keys, int32_t* vals, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    keys[i] = DGR_EMPTY_KEY;
    vals[i] = INT_MAX;
  }
}

__global__ void insert_unique_kernel(const int32_t* __restrict__ coords, const int32_t* __restrict__ n_dev,
                                     int64_t n_max, int ncols, const dgr_keyspec_t* __restrict__ spec_p,
                                     uint64_t* keys, int32_t* vals, uint64_t mask) {
  const int n = dev_count(n_dev, n_max);
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const dgr_keyspec_t s = *spec_p;
  const uint32_t sl = dgr_hash_insert(keys, mask, dgr_pack_key(coords + r * ncols, s));
  atomicMin(vals + sl, (int32_t)r);      // distinct rows: one writer; min keeps duplicates deterministic
}

// ---------------------------------------------------------------------------------------
// coarse (strided) coordinate maps of up to kMaxLevels strides in one launch per phase
// ---------------------------------------------------------------------------------------
struct CoarseArgs {
  int n_levels;
  int stride[kMaxLevels];
  uint64_t* keys[kMaxLevels];
  int32_t* vals[kMaxLevels];
  int32_t* slot[kMaxLevels];
  int32_t* scan[kMaxLevels];
  int32_t* coords[kMaxLevels];
  int32_t* n_out[kMaxLevels];
};

__global__ void coarse_insert_kernel(const int32_t* __restrict__ fine, const int32_t* __restrict__ n_dev,
                                     int64_t n_max, int ncols, const dgr_keyspec_t* __restrict__ spec_p,
                                     uint64_t mask, CoarseArgs a) {
  const int n = dev_count(n_dev, n_max);
  const int l = blockIdx.y;
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const dgr_keyspec_t s = *spec_p;
  const uint32_t sl = dgr_hash_insert(a.keys[l], mask, pack_key_strided(fine + r * ncols, s, a.stride[l]));
  atomicMin(a.vals[l] + sl, (int32_t)r);
  a.slot[l][r] = (int32_t)sl;
}

// winners (first fine row of every coarse cell) are marked by keeping slot >= 0, losers get ~slot;
// per-2048-row block winner counts go to scan[l][block]
__global__ void coarse_flag_kernel(const int32_t* __restrict__ n_dev, int64_t n_max, CoarseArgs a) {
  const int n = dev_count(n_dev, n_max);
  const int l = blockIdx.y;
  const int64_t start = (int64_t)blockIdx.x * kScanElems;
  int c = 0;
#pragma unroll
  for (int e = 0; e < kScanElems / kThreads; ++e) {
    const int64_t r = start + e * kThreads + threadIdx.x;
    if (r < n) {
      const int sl = a.slot[l][r];
      const bool win = a.vals[l][sl] == (int32_t)r;
      if (!win) a.slot[l][r] = ~sl;
      c += win;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  __shared__ int ws[kThreads / 32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += ws[w];
    a.scan[l][blockIdx.x] = t;
  }
}

// exclusive scan of `nb` ints in place with one block; returns the total (valid in every thread)
__device__ int block_scan_inplace(int32_t* cnt, int64_t nb) {
  __shared__ int carry_s;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < nb; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const int v = (i < nb) ? cnt[i] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const int sm = wsum[w];
      if (w < warp) wbase += sm;
      tot += sm;
    }
    const int carry = carry_s;
    if (i < nb) cnt[i] = carry + wbase + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  return carry_s;
}

__global__ void coarse_scan_kernel(const int32_t* __restrict__ n_dev, int64_t n_max, CoarseArgs a) {
  const int n = dev_count(n_dev, n_max);
  const int l = blockIdx.x;
  const int64_t nb = (n + kScanElems - 1) / kScanElems;
  const int total = block_scan_inplace(a.scan[l], nb);
  if (threadIdx.x == 0) a.n_out[l][0] = total;
}

// rank the winners in row order: coarse row `pos` = floor(fine row r); table value <- pos
__global__ void coarse_scatter_kernel(const int32_t* __restrict__ fine, const int32_t* __restrict__ n_dev,
                                      int64_t n_max, int ncols, CoarseArgs a) {
  const int n = dev_count(n_dev, n_max);
  const int l = blockIdx.y;
  const int64_t start = (int64_t)blockIdx.x * kScanElems + (int64_t)threadIdx.x * 8;
  if ((int64_t)blockIdx.x * kScanElems >= n) return;      // uniform per block
  int sl[8], c = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int64_t r = start + e;
    sl[e] = (r < n) ? a.slot[l][r] : -1;
    c += (sl[e] >= 0);
  }
  int pos = a.scan[l][blockIdx.x] + dgr_block_exclusive_scan_256(c, nullptr);
  const int stride = a.stride[l];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (sl[e] >= 0) {
      const int64_t r = start + e;
      int32_t* dst = a.coords[l] + (int64_t)pos * ncols;
      dst[0] = fine[r * ncols];
      for (int col = 1; col < ncols; ++col) dst[col] = floor_to(fine[r * ncols + col], stride);
      a.vals[l][sl[e]] = pos;
      ++pos;
    }
  }
}

// ---------------------------------------------------------------------------------------
// blocked Bloom filter of a table: both bits of a key live in ONE 32-bit word
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bloom_mix(uint64_t key) {
  uint32_t h = (uint32_t)key ^ ((uint32_t)(key >> 32) * 0x9E3779B1u);
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t bloom_bits(uint32_t h) { return (1u << ((h >> 20) & 31)) | (1u << ((h >> 26) & 31)); }

__global__ void bloom2_build_kernel(const uint64_t* __restrict__ keys, int64_t cap, uint32_t* words,
                                    uint32_t word_mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  const uint64_t k = keys[i];
  if (k == DGR_EMPTY_KEY) return;
  const uint32_t h = bloom_mix(k);
  atomicOr(words + (h & word_mask), bloom_bits(h));
}

// ---------------------------------------------------------------------------------------
// kernel maps: probe -> bit masks + block counts, scan, fill
// ---------------------------------------------------------------------------------------
// Probe kernel.  grid (row blocks of 512 rows, kappa chunks); thread = output row.
//
// Two phases per warp, decoupled by a shared-memory queue:
//   generate  every lane tests its row's neighbour key against the Bloom filter (shared memory, no global
//             traffic); lanes whose key MAY exist push (kappa, lane) into the warp's queue;
//   drain     whenever 32 candidates are queued, every lane takes ONE and does the hash-table lookup (L2):
//             all 32 lanes busy, 32 independent L2 chains in flight.
// Probing in place instead would make the whole warp wait for an L2 round trip whenever ANY of its 32 lanes
// passes the filter (a 3 % pass rate per lane is 62 % per warp) with one or two lanes doing useful work.
// Output: kDense ? nbr[kappa * nbr_stride + j] = input row or -1
//                : bits[kappa * W + w] = ballot of warp w's rows (+ per-(kappa, 256-word block) counts)
template <bool kBloom, bool kDense>
__global__ void __launch_bounds__(kProbeThreads, 2)
kmap_probe_kernel(const int32_t* __restrict__ out_coords, const int32_t* __restrict__ n_out_dev, int64_t n_out_max,
                  int ncols, const dgr_keyspec_t* __restrict__ spec_p, const uint64_t* __restrict__ keys,
                  const int32_t* __restrict__ vals, uint64_t mask, const uint32_t* __restrict__ bloom,
                  uint32_t n_bloom_words, const int32_t* __restrict__ offsets, int K, int k_per_block,
                  uint32_t* __restrict__ bits, int W, int32_t* block_cnt, int bpk, int32_t* __restrict__ nbr,
                  int64_t nbr_stride, int32_t* hit_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kWarps = kProbeThreads / 32;
  long long* delta = reinterpret_cast<long long*>(smem_raw);                               // [k_per_block]
  uint32_t* mtile = reinterpret_cast<uint32_t*>(smem_raw + (size_t)k_per_block * 8);       // [kWarps][k_per_block]
  uint16_t* queue = reinterpret_cast<uint16_t*>(mtile + (size_t)kWarps * k_per_block);     // [kWarps][64]
  uint32_t* bloom_s = reinterpret_cast<uint32_t*>(queue + kWarps * 64);                    // [n_bloom_words]
  const int n_out = dev_count(n_out_dev, n_out_max);
  const int k0 = blockIdx.y * k_per_block;
  const int kn = min(k_per_block, K - k0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int64_t)blockIdx.x * kProbeThreads >= n_out) {
    // rows beyond the actual count: their mask words must still read as zero (a dense table is never read there)
    if (!kDense) {
      const int w = blockIdx.x * kWarps + warp;
      if (lane == 0 && w < W)
        for (int kk = 0; kk < kn; ++kk) bits[(int64_t)(k0 + kk) * W + w] = 0u;
    }
    return;
  }
  const dgr_keyspec_t s = *spec_p;
  for (int kk = threadIdx.x; kk < kn; kk += blockDim.x) {
    long long d = 0;
    const int32_t* o = offsets + (int64_t)(k0 + kk) * (ncols - 1);
    for (int a = 0; a < ncols - 1; ++a) d += (long long)o[a] * (1ll << s.shift[a + 1]);
    delta[kk] = d;
  }
  if (!kDense)
    for (int e = threadIdx.x; e < kWarps * k_per_block; e += blockDim.x) mtile[e] = 0u;
  if (kBloom)
    for (uint32_t i = threadIdx.x; i < n_bloom_words; i += blockDim.x) bloom_s[i] = bloom[i];
  __syncthreads();
  const int64_t j = (int64_t)blockIdx.x * kProbeThreads + threadIdx.x;
  const bool live = j < n_out;
  const uint64_t key = live ? dgr_pack_key(out_coords + j * ncols, s) : 0;
  const uint32_t key_lo = (uint32_t)key, key_hi = (uint32_t)(key >> 32);
  const int64_t row0 = j - lane;                       // first row of this warp
  const uint32_t wm = n_bloom_words - 1;
  uint16_t* q = queue + warp * 64;
  uint32_t* mt = mtile + warp * k_per_block;
  int qn = 0;                                          // queued candidates (warp-uniform)

  auto drain = [&](int count) {                        // lanes < count take one candidate each
    __syncwarp();
    const uint32_t e = lane < count ? q[lane] : 0u;
    const int src = e & 31, kq = e >> 5;
    const uint32_t lo = __shfl_sync(0xffffffffu, key_lo, src), hi = __shfl_sync(0xffffffffu, key_hi, src);
    if (lane < count) {
      const uint64_t qq = (((uint64_t)hi << 32) | lo) + (uint64_t)delta[kq];
      const int32_t i = dgr_hash_lookup(keys, vals, mask, qq);
      if (kDense) {
        nbr[(int64_t)(k0 + kq) * nbr_stride + row0 + src] = i;
        if (hit_count != nullptr) {
          const uint32_t act = __activemask();
          const uint32_t hits = __ballot_sync(act, i >= 0);
          if (lane == (__ffs(act) - 1) && hits) atomicAdd(hit_count, __popc(hits));
        }
      } else if (i >= 0) {
        atomicOr(mt + kq, 1u << src);
      }
    }
    __syncwarp();
    // move the (at most 31) entries behind the drained ones to the front
    const uint32_t rest = (lane + 32 < qn) ? q[lane + 32] : 0u;
    __syncwarp();
    if (lane + 32 < qn) q[lane] = (uint16_t)rest;
    qn = qn > 32 ? qn - 32 : 0;
    __syncwarp();
  };

  if (!kBloom && !kDense) {
    // no filter (3^3 kernels: ~60 % of the probes hit): probe in place, four independent lookups in flight
    const int w = (int)(j >> 5);
    const int cnt_col = w / kCntWords;
    for (int kk = 0; kk < kn; kk += 4) {
      bool found[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        found[u] = live && kk + u < kn && dgr_hash_lookup(keys, vals, mask, key + (uint64_t)delta[min(kk + u, kn - 1)]) >= 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kk + u >= kn) break;                                     // uniform
        const uint32_t m = __ballot_sync(0xffffffffu, found[u]);
        if (lane == 0 && w < W) {
          bits[(int64_t)(k0 + kk + u) * W + w] = m;
          if (m) atomicAdd(block_cnt + (int64_t)(k0 + kk + u) * bpk + cnt_col, __popc(m));
        }
      }
    }
    return;
  }
  for (int kk = 0; kk < kn; ++kk) {
    bool maybe = live;
    if (kBloom && live) {
      const uint32_t h = bloom_mix(key + (uint64_t)delta[kk]);
      const uint32_t b = bloom_bits(h);
      maybe = (bloom_s[h & wm] & b) == b;
    }
    if (kDense && live && !maybe) nbr[(int64_t)(k0 + kk) * nbr_stride + j] = -1;
    const uint32_t cand = __ballot_sync(0xffffffffu, maybe);
    if (cand) {
      if (maybe) q[qn + __popc(cand & ((1u << lane) - 1u))] = (uint16_t)((kk << 5) | lane);
      qn += __popc(cand);
      if (qn >= 32) drain(32);
    }
  }
  if (qn > 0) drain(qn);
  if (!kDense) {
    __syncwarp();
    const int w = (int)(j >> 5);
    if (w < W) {
      const int cnt_col = w / kCntWords;
      for (int kk = lane; kk < kn; kk += 32) {
        const uint32_t m = mt[kk];
        bits[(int64_t)(k0 + kk) * W + w] = m;
        if (m) atomicAdd(block_cnt + (int64_t)(k0 + kk) * bpk + cnt_col, __popc(m));
      }
    }
  }
}

// exclusive scan of the K x bpk block counts (one block), bucket offsets, work-list sizes
// meta[0..4] = (pairs P, tiles, tiles with every offset rounded up to an even count,
//              non-empty offsets, key-overflow flag of spec)
__global__ void kmap_scan_kernel(int32_t* cnt, int K, int bpk, int tile_rows, int32_t* kofs,
                                 int32_t* meta, const dgr_keyspec_t* spec) {
  const int64_t nb = (int64_t)K * bpk;
  const int total = block_scan_inplace(cnt, nb);
  if (threadIdx.x == 0) cnt[nb] = total;
  __syncthreads();
  __shared__ int red[3][32];
  int tiles = 0, ptiles = 0, nonempty = 0;
  for (int k = threadIdx.x; k <= K; k += blockDim.x) {
    const int lo = cnt[(int64_t)k * bpk];        // k == K reads cnt[nb] = total
    kofs[k] = lo;
    if (k < K) {
      const int hi = (k + 1 < K) ? cnt[(int64_t)(k + 1) * bpk] : total;
      const int t = (hi - lo + tile_rows - 1) / tile_rows;
      tiles += t;
      ptiles += (t + 1) & ~1;
      nonempty += (hi > lo);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    tiles += __shfl_xor_sync(0xffffffffu, tiles, d);
    ptiles += __shfl_xor_sync(0xffffffffu, ptiles, d);
    nonempty += __shfl_xor_sync(0xffffffffu, nonempty, d);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = tiles;
    red[1][threadIdx.x >> 5] = ptiles;
    red[2][threadIdx.x >> 5] = nonempty;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, b = 0, c = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
    const int ovf = spec != nullptr ? spec->overflow : 0;
    kofs[K + 1] = ovf;
    meta[0] = total; meta[1] = a; meta[2] = b; meta[3] = c; meta[4] = ovf;
  }
}

// Fill: grid (bpk, K); block (b, kappa) owns the 256 mask words [b * 256, +256) of bucket kappa, one per thread;
// its output offset is the scanned count of exactly that block, so blocks without a hit leave at once (a 6-D
// map at stride 1 has a hit in ~10 % of its blocks).  A word's 32 hits are resolved by the 32 LANES of a warp in
// parallel (each lane: one hash lookup, consecutive output positions): the longest serial chain is the 32 words
// of a warp, not the hits of a thread.
__global__ void __launch_bounds__(kThreads)
kmap_fill_kernel(const uint32_t* __restrict__ bits, int W, int bpk, const int32_t* __restrict__ block_ofs,
                 const int32_t* __restrict__ out_coords, int ncols, const dgr_keyspec_t* __restrict__ spec_p,
                 const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals, uint64_t mask,
                 const int32_t* __restrict__ offsets, int32_t* __restrict__ in_idx, int32_t* __restrict__ out_idx) {
  __shared__ uint32_t s_mask[kThreads];
  __shared__ int s_base[kThreads];
  const int kappa = blockIdx.y;
  const int64_t e = (int64_t)kappa * bpk + blockIdx.x;
  const int base = block_ofs[e];
  if (block_ofs[e + 1] == base) return;                        // no pair in this block (uniform)
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int w_own = blockIdx.x * kCntWords + t;
  const uint32_t m_own = w_own < W ? bits[(int64_t)kappa * W + w_own] : 0u;
  const int excl = dgr_block_exclusive_scan_256(__popc(m_own), nullptr);
  s_mask[t] = m_own;
  s_base[t] = base + excl;
  __syncthreads();
  const dgr_keyspec_t s = *spec_p;
  long long d = 0;
  const int32_t* o = offsets + (int64_t)kappa * (ncols - 1);
  for (int a = 0; a < ncols - 1; ++a) d += (long long)o[a] * (1ll << s.shift[a + 1]);
  const int w_first = blockIdx.x * kCntWords + warp * 32;      // this warp resolves its own 32 words
  for (int u = 0; u < 32; ++u) {
    const uint32_t m = s_mask[warp * 32 + u];
    if (m == 0u) continue;                                     // uniform
    if ((m >> lane) & 1u) {
      const int64_t j = (int64_t)(w_first + u) * 32 + lane;
      const int pos = s_base[warp * 32 + u] + __popc(m & ((1u << lane) - 1u));
      const uint64_t qq = dgr_pack_key(out_coords + j * ncols, s) + (uint64_t)d;
      in_idx[pos] = dgr_hash_lookup(keys, vals, mask, qq);
      out_idx[pos] = (int32_t)j;
    }
  }
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

int32_t dgr_compact_voxel_pair(const int32_t* raw_coords, const int32_t* sel, const int32_t* n_unique,
                               int64_t n_raw0, int64_t n_raw1, const void* xyz0, int32_t is_f64_0,
                               const void* xyz1, int32_t is_f64_1, int32_t* coords, float* xyz,
                               int32_t* counts, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_max = n_raw0 + n_raw1;
  unsigned blocks = dgr_blocks(n_max > 0 ? n_max : 1, kThreads);
  if (blocks > 1184) blocks = 1184;
#define DGR_CV(T0, T1)                                                                                         \
  compact_voxels_kernel<T0, T1><<<blocks, kThreads, 0, st>>>(raw_coords, sel, n_unique, n_raw0, (const T0*)xyz0, \
                                                             (const T1*)xyz1, coords, xyz, counts)
  if (is_f64_0 && is_f64_1) DGR_CV(double, double);
  else if (is_f64_0) DGR_CV(double, float);
  else if (is_f64_1) DGR_CV(float, double);
  else DGR_CV(float, float);
#undef DGR_CV
  dgr_note_launches(1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

int32_t dgr_table_build_unique(const int32_t* coords, int64_t n_max, const int32_t* n_dev, int32_t ncols,
                               const dgr_keyspec_t* spec, uint64_t* keys, int32_t* vals, int64_t cap,
                               void* stream) {
  DGR_ARG_CHECK(cap > 0 && (cap & (cap - 1)) == 0, "capacity must be a power of two");
  DGR_ARG_CHECK(cap >= 2 * n_max, "capacity must be at least 2 n_max");
  cudaStream_t st = (cudaStream_t)stream;
  table_clear_kernel<<<dgr_blocks(cap, kThreads), kThreads, 0, st>>>(keys, vals, cap);
  if (n_max > 0)
    insert_unique_kernel<<<dgr_blocks(n_max, kThreads), kThreads, 0, st>>>(coords, n_dev, n_max, ncols, spec, keys,
                                                                         vals, (uint64_t)cap - 1);
  dgr_note_launches(n_max > 0 ? 2 : 1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

int64_t dgr_coarse_scan_elems(int64_t n_max) { return (n_max + kScanElems - 1) / kScanElems + 2; }

int32_t dgr_coarse_maps(const int32_t* fine, int64_t n_max, const int32_t* n_dev, int32_t ncols,
                        const dgr_keyspec_t* spec, int32_t n_levels, const int32_t* strides, uint64_t* keys,
                        int32_t* vals, int64_t cap, int32_t* coords_out, int32_t* n_out, int32_t* slot_ws,
                        int32_t* scan_ws, void* stream) {
  DGR_ARG_CHECK(n_levels >= 1 && n_levels <= kMaxLevels, "1..4 levels per call");
  DGR_ARG_CHECK(cap > 0 && (cap & (cap - 1)) == 0 && cap >= 2 * n_max, "capacity: power of two >= 2 n_max");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nmx = n_max > 0 ? n_max : 1;
  const int64_t scan_elems = dgr_coarse_scan_elems(nmx);
  CoarseArgs a;
  a.n_levels = n_levels;
  for (int l = 0; l < kMaxLevels; ++l) {
    const int k = l < n_levels ? l : 0;
    DGR_ARG_CHECK(strides[k] >= 1, "stride must be positive");
    a.stride[l] = strides[k];
    a.keys[l] = keys + (int64_t)k * cap;
    a.vals[l] = vals + (int64_t)k * cap;
    a.slot[l] = slot_ws + (int64_t)k * nmx;
    a.scan[l] = scan_ws + (int64_t)k * scan_elems;
    a.coords[l] = coords_out + (int64_t)k * nmx * ncols;
    a.n_out[l] = n_out + k;
  }
  table_clear_kernel<<<dgr_blocks(cap * n_levels, kThreads), kThreads, 0, st>>>(keys, vals, cap * n_levels);
  const unsigned nb = dgr_blocks(nmx, kScanElems);
  coarse_insert_kernel<<<dim3(dgr_blocks(nmx, kThreads), n_levels), kThreads, 0, st>>>(fine, n_dev, n_max, ncols, spec,
                                                                                      (uint64_t)cap - 1, a);
  coarse_flag_kernel<<<dim3(nb, n_levels), kThreads, 0, st>>>(n_dev, n_max, a);
  coarse_scan_kernel<<<n_levels, 1024, 0, st>>>(n_dev, n_max, a);
  coarse_scatter_kernel<<<dim3(nb, n_levels), kThreads, 0, st>>>(fine, n_dev, n_max, ncols, a);
  dgr_note_launches(5);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

int32_t dgr_bloom2_build(const uint64_t* keys, int64_t cap, uint32_t* words, int64_t n_words, void* stream) {
  DGR_ARG_CHECK(cap > 0 && (cap & (cap - 1)) == 0, "capacity must be a power of two");
  DGR_ARG_CHECK(n_words >= 32 && (n_words & (n_words - 1)) == 0, "n_words must be a power of two");
  cudaStream_t st = (cudaStream_t)stream;
  DGR_CUDA_CHECK(cudaMemsetAsync(words, 0, (size_t)n_words * 4, st));
  bloom2_build_kernel<<<dgr_blocks(cap, kThreads), kThreads, 0, st>>>(keys, cap, words, (uint32_t)n_words - 1);
  dgr_note_launches(1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

/* words per offset of the bit-mask representation (one word per 32 output rows) */
int64_t dgr_kmap_mask_words(int64_t n_out_max) {
  const int64_t w = (n_out_max + 31) / 32;
  return w < 1 ? 1 : w;
}
/* ints of the block-count workspace: K * ceil(W / 2048) + 2 */
int64_t dgr_kmap_cnt_elems(int32_t K, int64_t n_out_max) {
  const int64_t W = dgr_kmap_mask_words(n_out_max);
  return (int64_t)K * ((W + kCntWords - 1) / kCntWords) + 2;
}

static size_t probe_smem_bytes(int k_per_block, int64_t n_bloom_words) {
  return (size_t)k_per_block * 8 + (size_t)(kProbeThreads / 32) * k_per_block * 4 + (size_t)(kProbeThreads / 32) * 64 * 2 +
         (size_t)n_bloom_words * 4;
}
static int probe_k_per_block(int K, unsigned row_blocks) {
  // kappa chunks: enough blocks for ~2 waves of the 148 SMs at 2 blocks per SM, at most 96 offsets per block
  int k_per_block = K;
  while (k_per_block > 16 && (int64_t)row_blocks * ((K + k_per_block - 1) / k_per_block) < 592) k_per_block = (k_per_block + 1) / 2;
  if (k_per_block > 96) k_per_block = 96;
  return (k_per_block + 3) & ~3;
}

int32_t dgr_kmap_probe(const int32_t* out_coords, int64_t n_out_max, const int32_t* n_out_dev, int32_t ncols,
                       const dgr_keyspec_t* spec, const uint64_t* in_keys, const int32_t* in_vals, int64_t in_cap,
                       const uint32_t* bloom_words, int64_t n_bloom_words, const int32_t* offsets, int32_t K,
                       uint32_t* bits, int32_t* block_cnt, int32_t* kofs, int32_t* meta, void* stream) {
  DGR_ARG_CHECK(in_cap > 0 && (in_cap & (in_cap - 1)) == 0, "capacity must be a power of two");
  DGR_ARG_CHECK(K >= 1 && K <= 65535, "K out of range");
  DGR_ARG_CHECK(bloom_words == nullptr || (n_bloom_words >= 32 && (n_bloom_words & (n_bloom_words - 1)) == 0 &&
                                           n_bloom_words <= 32768),
                "bloom: power of two, 32..32768 words");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nmx = n_out_max > 0 ? n_out_max : 1;
  const int W = (int)dgr_kmap_mask_words(nmx);
  const int bpk = (W + kCntWords - 1) / kCntWords;
  DGR_CUDA_CHECK(cudaMemsetAsync(block_cnt, 0, (size_t)dgr_kmap_cnt_elems(K, nmx) * sizeof(int32_t), st));
  const unsigned row_blocks = dgr_blocks(nmx, kProbeThreads);
  const int k_per_block = probe_k_per_block(K, row_blocks);
  const dim3 grid(row_blocks, (K + k_per_block - 1) / k_per_block);
  if (bloom_words != nullptr) {
    const size_t smem = probe_smem_bytes(k_per_block, n_bloom_words);
    DGR_ENSURE_SMEM((kmap_probe_kernel<true, false>), smem);
    kmap_probe_kernel<true, false><<<grid, kProbeThreads, smem, st>>>(
        out_coords, n_out_dev, n_out_max, ncols, spec, in_keys, in_vals, (uint64_t)in_cap - 1, bloom_words,
        (uint32_t)n_bloom_words, offsets, K, k_per_block, bits, W, block_cnt, bpk, nullptr, 0, nullptr);
  } else {
    const size_t smem = probe_smem_bytes(k_per_block, 0);
    kmap_probe_kernel<false, false><<<grid, kProbeThreads, smem, st>>>(
        out_coords, n_out_dev, n_out_max, ncols, spec, in_keys, in_vals, (uint64_t)in_cap - 1, nullptr, 1u, offsets, K,
        k_per_block, bits, W, block_cnt, bpk, nullptr, 0, nullptr);
  }
  kmap_scan_kernel<<<1, 1024, 0, st>>>(block_cnt, K, bpk, 128, kofs, meta, spec);
  dgr_note_launches(2);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

int32_t dgr_kmap_fill(const uint32_t* bits, const int32_t* block_cnt, int32_t K, int64_t n_out_max,
                      const int32_t* out_coords, int32_t ncols, const dgr_keyspec_t* spec,
                      const uint64_t* in_keys, const int32_t* in_vals, int64_t in_cap, const int32_t* offsets,
                      int32_t* in_idx, int32_t* out_idx, void* stream) {
  const int64_t nmx = n_out_max > 0 ? n_out_max : 1;
  const int W = (int)dgr_kmap_mask_words(nmx);
  const int bpk = (W + kCntWords - 1) / kCntWords;
  kmap_fill_kernel<<<dim3(bpk, K), kThreads, 0, (cudaStream_t)stream>>>(bits, W, bpk, block_cnt, out_coords, ncols,
                                                                            spec, in_keys, in_vals, (uint64_t)in_cap - 1,
                                                                            offsets, in_idx, out_idx);
  dgr_note_launches(1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

int32_t dgr_kmap_dense(const int32_t* out_coords, int64_t n_out_max, const int32_t* n_out_dev, int32_t ncols,
                       const dgr_keyspec_t* spec, const uint64_t* in_keys, const int32_t* in_vals, int64_t in_cap,
                       const uint32_t* bloom_words, int64_t n_bloom_words, const int32_t* offsets, int32_t K,
                       int32_t* nbr, int64_t nbr_stride, int32_t* hit_count, void* stream) {
  DGR_ARG_CHECK(in_cap > 0 && (in_cap & (in_cap - 1)) == 0, "capacity must be a power of two");
  DGR_ARG_CHECK(nbr_stride >= n_out_max, "row stride below the row bound");
  if (hit_count != nullptr) DGR_CUDA_CHECK(cudaMemsetAsync(hit_count, 0, sizeof(int32_t), (cudaStream_t)stream));
  DGR_ARG_CHECK(bloom_words == nullptr || (n_bloom_words >= 32 && (n_bloom_words & (n_bloom_words - 1)) == 0 &&
                                           n_bloom_words <= 32768),
                "bloom: power of two, 32..32768 words");
  if (n_out_max == 0) return DGR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned row_blocks = dgr_blocks(n_out_max, kProbeThreads);
  const int k_per_block = probe_k_per_block(K, row_blocks);
  const dim3 grid(row_blocks, (K + k_per_block - 1) / k_per_block);
  if (bloom_words != nullptr) {
    const size_t smem = probe_smem_bytes(k_per_block, n_bloom_words);
    DGR_ENSURE_SMEM((kmap_probe_kernel<true, true>), smem);
    kmap_probe_kernel<true, true><<<grid, kProbeThreads, smem, st>>>(
        out_coords, n_out_dev, n_out_max, ncols, spec, in_keys, in_vals, (uint64_t)in_cap - 1, bloom_words,
        (uint32_t)n_bloom_words, offsets, K, k_per_block, nullptr, 0, nullptr, 0, nbr, nbr_stride, hit_count);
  } else {
    const size_t smem = probe_smem_bytes(k_per_block, 0);
    kmap_probe_kernel<false, true><<<grid, kProbeThreads, smem, st>>>(
        out_coords, n_out_dev, n_out_max, ncols, spec, in_keys, in_vals, (uint64_t)in_cap - 1, nullptr, 1u, offsets, K,
        k_per_block, nullptr, 0, nullptr, 0, nbr, nbr_stride, hit_count);
  }
  dgr_note_launches(1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

}  // extern "C"
