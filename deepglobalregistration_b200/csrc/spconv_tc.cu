// Sparse convolution forward on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Same contract as dgr_spconv_fwd (gather -> per-offset sub-GEMM -> scatter-add over the
// (kappa, j)-sorted pair lists) with the sub-GEMM issued as tcgen05.mma kind::tf32 and the
// accumulator in tensor memory:
//
//   * one work item = one 128-pair tile of one kernel offset, M = 128 rows (pairs),
//     N = cout (16..256), K = cin in chunks of 32 floats (one 128-byte swizzle row);
//   * 8 loader warps gather the 128 input rows (coalesced 16-byte pieces, 8 lanes per row),
//     and stream the [cout x 32] slab of the offset's weight matrix, split every fp32 value
//     into a TF32 "hi" part and a TF32 "lo" residual in registers, and store both into
//     shared memory in the canonical K-major SWIZZLE_128B layout;
//   * one elected thread of the MMA warp issues, per 8-wide k-step, the three products
//     hi*hi + lo*hi + hi*lo (3xTF32: fp32-accurate to ~2^-21 relative) into TMEM;
//     tcgen05.commit on an mbarrier frees the shared-memory stage / publishes the tile;
//   * 4 epilogue warps read the accumulator with tcgen05.ld (warp w owns TMEM lanes
//     32(w%4).. = pairs 32(w%4)..) and scatter-add rows with red.global.add.v4.f32; two
//     accumulators in TMEM let the epilogue of tile t overlap the MMAs of tile t+1.
//
// CTAs are persistent (grid = resident CTAs); stages are mbarrier-pipelined so the gather
// of chunk c+1 overlaps the MMAs of chunk c.  Weights are expected TRANSPOSED per offset,
// [K, cout, cin] (K-major B operand); the host caches that layout per layer.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kLoaderWarps = 8;
constexpr int kLoaderThreads = kLoaderWarps * 32;   // warps 0..7: gather + TF32 split
constexpr int kMmaWarp = kLoaderWarps;               // warp 8: tcgen05.mma issuer
constexpr int kEpiWarp0 = kLoaderWarps + 1;          // warps 9..12: TMEM -> red.global
constexpr int kThreadsTC = (kLoaderWarps + 5) * 32;  // 416
constexpr int kTileM = 128;
constexpr int kChunk = 32;               // floats of K per stage (128 bytes)
constexpr int kATileBytes = kTileM * 128;

// split one float4 into tf32 hi / lo and store both at the swizzled 16-byte slot
__device__ __forceinline__ void split_store(float4 v, unsigned char* hi_tile, unsigned char* lo_tile, int row,
                                            int piece) {
  float4 h, l;
  h.x = tf32_round(v.x); h.y = tf32_round(v.y); h.z = tf32_round(v.z); h.w = tf32_round(v.w);
  l.x = tf32_round(v.x - h.x); l.y = tf32_round(v.y - h.y);
  l.z = tf32_round(v.z - h.z); l.w = tf32_round(v.w - h.w);
  const int off = row * 128 + ((piece ^ (row & 7)) << 4);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

struct TcShared {
  unsigned long long full[4];
  unsigned long long empty[4];
  unsigned long long acc_full[2];
  unsigned long long acc_empty[2];
  uint32_t tmem_base;
  int s_in[kTileM];
};

// Warp-specialised persistent kernel.  Roles iterate the same tile sequence
// (tile = blockIdx.x + i * gridDim.x) and meet only through mbarriers:
//   loaders  --full[s]-->  MMA issuer  --empty[s]-->  loaders        (shared-memory stages)
//   MMA issuer  --acc_full[b]-->  epilogue  --acc_empty[b]-->  MMA   (two TMEM accumulators)
__global__ void __launch_bounds__(kThreadsTC, 1)
spconv_tc_kernel(const float* __restrict__ in_feat, int cin, const float* __restrict__ wt, int cout,
                 const int32_t* __restrict__ in_idx, const int32_t* __restrict__ out_idx,
                 const int32_t* __restrict__ kofs, const int32_t* __restrict__ tile_k,
                 const int32_t* __restrict__ tile_start, int n_tiles, int n_stages, int tmem_cols,
                 int passes, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  TcShared& sh = *reinterpret_cast<TcShared*>(smem_dyn);
  // stage buffers start at the next 1024-byte boundary (SWIZZLE_128B atom alignment)
  unsigned char* stage0 = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_dyn) + sizeof(TcShared) + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = cout * 128;
  const int stage_bytes = 2 * kATileBytes + 2 * b_tile_bytes;
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int n_chunks = cin / kChunk;
  const uint32_t acc_stride = (uint32_t)tmem_cols >> 1;   // columns between the two accumulators

  if (t == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(smem_u32(&sh.full[s]), kLoaderThreads);
      mbar_init(smem_u32(&sh.empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&sh.acc_full[b]), 1);
      mbar_init(smem_u32(&sh.acc_empty[b]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&sh.tmem_base)),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;

  if (warp < kLoaderWarps) {
    // ================================ loaders ============================================
    const int piece = t & 7, rgrp = t >> 3;   // 8 lanes cover one 128-byte row; 32 row groups
    const int b_rows_per_thread = (cout + 31) / 32;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int kappa = tile_k[tile];
      const int p0 = tile_start[tile];
      const int rows = min(kTileM, kofs[kappa + 1] - p0);
      asm volatile("bar.sync 1, 256;" ::: "memory");   // everyone is done with the previous s_in
      if (t < kTileM) sh.s_in[t] = (t < rows) ? in_idx[p0 + t] : -1;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int src_row[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) src_row[i] = sh.s_in[i * 32 + rgrp];
      const float* wk = wt + (size_t)kappa * cout * cin;
      for (int c = 0; c < n_chunks; ++c, ++it) {
        const int s = it % n_stages;
        const uint32_t ph = (it / n_stages) & 1;
        const int c0 = c * kChunk + piece * 4;
        // issue every global load of this chunk before touching any of them
        float4 av[4], bv[8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          av[i] = src_row[i] >= 0
                      ? __ldg(reinterpret_cast<const float4*>(in_feat + (size_t)src_row[i] * cin + c0))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 32 + rgrp;
          if (i < b_rows_per_thread && r < cout)
            bv[i] = __ldg(reinterpret_cast<const float4*>(wk + (size_t)r * cin + c0));
        }
        mbar_wait(smem_u32(&sh.empty[s]), ph ^ 1);
        unsigned char* a_hi = stage0 + (size_t)s * stage_bytes;
        unsigned char* a_lo = a_hi + kATileBytes;
        unsigned char* b_hi = a_lo + kATileBytes;
        unsigned char* b_lo = b_hi + b_tile_bytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) split_store(av[i], a_hi, a_lo, i * 32 + rgrp, piece);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 32 + rgrp;
          if (i < b_rows_per_thread && r < cout) split_store(bv[i], b_hi, b_lo, r, piece);
        }
        fence_proxy_async();
        mbar_arrive(smem_u32(&sh.full[s]));
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer =========================================
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = cout, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(cout >> 3) << 17) |
                           ((uint32_t)(kTileM >> 4) << 24);
    uint32_t it = 0, tile_iter = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
      const uint32_t buf = tile_iter & 1;
      const uint32_t tmem_d = tmem_base + buf * acc_stride;
      mbar_wait(smem_u32(&sh.acc_empty[buf]), ((tile_iter >> 1) & 1) ^ 1);   // epilogue drained it
      tc_fence_after();
      for (int c = 0; c < n_chunks; ++c, ++it) {
        const int s = it % n_stages;
        const uint32_t ph = (it / n_stages) & 1;
        mbar_wait(smem_u32(&sh.full[s]), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = smem_u32(stage0 + (size_t)s * stage_bytes);
          const uint32_t a_lo = a_hi + kATileBytes;
          const uint32_t b_hi = a_lo + kATileBytes;
          const uint32_t b_lo = b_hi + b_tile_bytes;
#pragma unroll
          for (int ks = 0; ks < kChunk / 8; ++ks) {
            const uint32_t ko = ks * 32;   // 8 tf32 = 32 bytes along K inside the swizzle row
            const uint64_t dah = umma_desc(a_hi + ko), dal = umma_desc(a_lo + ko);
            const uint64_t dbh = umma_desc(b_hi + ko), dbl = umma_desc(b_lo + ko);
            tc_mma_tf32(tmem_d, dah, dbh, idesc, (c | ks) != 0);
            if (passes == 3) {
              tc_mma_tf32(tmem_d, dal, dbh, idesc, 1);
              tc_mma_tf32(tmem_d, dah, dbl, idesc, 1);
            }
          }
          tc_commit(smem_u32(&sh.empty[s]));
          if (c == n_chunks - 1) tc_commit(smem_u32(&sh.acc_full[buf]));
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue ===========================================
    const int lane_grp = warp & 3;            // TMEM lanes 32 * (warp % 4) .. + 31
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
      const uint32_t buf = tile_iter & 1;
      const int kappa = tile_k[tile];
      const int p0 = tile_start[tile];
      const int rows = min(kTileM, kofs[kappa + 1] - p0);
      const int r = lane_grp * 32 + lane;
      const int j = r < rows ? out_idx[p0 + r] : -1;
      mbar_wait(smem_u32(&sh.acc_full[buf]), (tile_iter >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * acc_stride + ((uint32_t)(lane_grp * 32) << 16);
      int col = 0;
      for (; col + 32 <= cout; col += 32) {
        uint32_t v[32];
        tc_ld32(taddr + col, v);
        if (j >= 0) {
          float* dst = out + (size_t)j * cout + col;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            red_add_v4(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
      }
      if (col < cout) {   // cout % 32 == 16
        uint32_t v[16];
        tc_ld16(taddr + col, v);
        if (j >= 0) {
          float* dst = out + (size_t)j * cout + col;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            red_add_v4(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&sh.acc_empty[buf]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)tmem_cols)
                 : "memory");
  }
}

// [K, cin, cout] -> [K, cout, cin]
__global__ void transpose_weight_kernel(const float* __restrict__ w, int cin, int cout, float* __restrict__ wt) {
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * cin * cout;
  int ci = blockIdx.y * 32 + threadIdx.y, co = blockIdx.x * 32 + threadIdx.x;
  if (ci < cin && co < cout) tile[threadIdx.y][threadIdx.x] = w[base + (size_t)ci * cout + co];
  __syncthreads();
  co = blockIdx.x * 32 + threadIdx.y;
  ci = blockIdx.y * 32 + threadIdx.x;
  if (ci < cin && co < cout) wt[base + (size_t)co * cin + ci] = tile[threadIdx.x][threadIdx.y];
}

}  // namespace

extern "C" {

// Layout transform the tensor-core path needs once per layer: W[K, cin, cout] -> Wt[K, cout, cin].
int32_t dgr_transpose_weight(const float* w, int32_t K, int32_t cin, int32_t cout, float* wt, void* stream) {
  DGR_ARG_CHECK(K >= 1 && cin >= 1 && cout >= 1, "bad weight shape");
  DGR_ARG_CHECK(K <= 65535, "K too large");
  dim3 grid((cout + 31) / 32, (cin + 31) / 32, K);
  transpose_weight_kernel<<<grid, dim3(32, 32), 0, (cudaStream_t)stream>>>(w, cin, cout, wt);
  dgr_note_launches(1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

// 1 if dgr_spconv_tc_fwd supports the shape (cin % 32 == 0, cout % 16 == 0, 16 <= cout <= 256).
int32_t dgr_spconv_tc_supported(int32_t cin, int32_t cout) {
  return (cin >= 32 && cin % 32 == 0 && cout >= 16 && cout <= 256 && cout % 16 == 0) ? 1 : 0;
}

// Tensor-core variant of dgr_spconv_fwd.  weight_t is [K, cout, cin] (dgr_transpose_weight).
// passes = 3: 3xTF32 (fp32-accurate, default); passes = 1: single TF32 product (~1e-3 rel.).
int32_t dgr_spconv_tc_fwd(const float* in_feat, int32_t cin, const float* weight_t, int32_t cout,
                          const int32_t* in_idx, const int32_t* out_idx, const int32_t* kofs,
                          const int32_t* tile_k, const int32_t* tile_start, int32_t n_tiles,
                          int32_t tile_rows, int32_t passes, float* out, void* stream) {
  DGR_ARG_CHECK(tile_rows == kTileM, "tile_rows must be 128");
  DGR_ARG_CHECK(dgr_spconv_tc_supported(cin, cout), "shape not supported by the tensor-core path");
  DGR_ARG_CHECK(passes == 1 || passes == 3, "passes must be 1 or 3");
  if (n_tiles == 0) return DGR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int stage_bytes = 2 * kATileBytes + 2 * cout * 128;
  const int n_chunks = cin / kChunk;
  int n_stages = (200 * 1024) / stage_bytes;
  if (n_stages > 4) n_stages = 4;
  if (n_stages < 2) n_stages = 2;
  const size_t smem = sizeof(TcShared) + 1024 + (size_t)n_stages * stage_bytes;
  int acc_cols = 32;                       // one accumulator: power of two >= cout
  while (acc_cols < cout) acc_cols <<= 1;
  const int tmem_cols = 2 * acc_cols;      // two accumulators: epilogue overlaps the next tile
  DGR_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  int dev = 0, sms = 148;
  DGR_CUDA_CHECK(cudaGetDevice(&dev));
  DGR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  (void)n_chunks;
  int grid = sms;                          // persistent: one warp-specialised CTA per SM
  if (grid > n_tiles) grid = n_tiles;
  spconv_tc_kernel<<<grid, kThreadsTC, smem, st>>>(in_feat, cin, weight_t, cout, in_idx, out_idx, kofs,
                                                   tile_k, tile_start, n_tiles, n_stages, tmem_cols, passes,
                                                   out);
  dgr_note_launches(1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

}  // extern "C"
