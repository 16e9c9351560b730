// Safeguard registration: RANSAC over a fixed correspondence list, the job open3d's
// registration_ransac_based_on_correspondence does for the reference
// (core/deep_global_registration.py:50-64, called from :302-315 when the weight sum is below
// the gate).  ransac_n = 4, TransformationEstimationPointToPoint(with_scaling=False), no
// checkers; a hypothesis is better when it has more inliers (|T p - q| < max_dist over ALL
// correspondences) or as many with a lower inlier RMSE.  The reference passes
// RANSACConvergenceCriteria(4000000, num_iterations = 80000): the 80000 lands in the
// confidence slot, open3d clamps it to 1, the early-exit estimate log(1 - confidence) / ... is
// never reached and all max_iteration hypotheses are evaluated - which is what happens here.
//
// One thread owns kHyp hypotheses (R, t in registers); the correspondences stream through
// shared memory in tiles and every thread reads the same element (broadcast), so the kernel
// is bound by the fp32 pipe: ~19 instructions per (hypothesis, correspondence).
// Sampling is a counter-based hash of (seed, hypothesis, slot): reproducible, and restated
// in oracle/ransac.py so the CPU checker draws the very same hypotheses.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kabsch.cuh"

namespace {

constexpr int kRansacThreads = 256;
constexpr int kHyp = 4;        // hypotheses per thread
constexpr int kTile = 512;     // correspondences per shared-memory tile (16 KB)

__global__ void ransac_pack_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                   const int32_t* __restrict__ idx0, const int32_t* __restrict__ idx1,
                                   int64_t n, float4* __restrict__ src, float4* __restrict__ tgt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t a = idx0 != nullptr ? (int64_t)idx0[i] : i;
  int64_t b = idx1 != nullptr ? (int64_t)idx1[i] : i;
  src[i] = make_float4(x[3 * a], x[3 * a + 1], x[3 * a + 2], 0.f);
  tgt[i] = make_float4(y[3 * b], y[3 * b + 1], y[3 * b + 2], 0.f);
}

// slot-th sample of hypothesis h: uniform in [0, n) (with replacement, like open3d's
// per-iteration uniform_int_distribution draws)
__device__ __forceinline__ uint32_t ransac_pick(uint64_t seed, uint64_t h, int slot, uint32_t n) {
  uint64_t z = dgr_mix64(seed + (h * 4 + (uint64_t)slot + 1) * 0x9E3779B97F4A7C15ull);
  return (uint32_t)(((z >> 32) * (uint64_t)n) >> 32);
}

// Umeyama without scaling on the 4 sampled correspondences, fp64
__device__ void ransac_hypothesis(const float4* __restrict__ src, const float4* __restrict__ tgt, uint32_t n,
                                  uint64_t seed, uint64_t h, double R[3][3], double t[3]) {
  double p[4][3], q[4][3], mp[3] = {0, 0, 0}, mq[3] = {0, 0, 0};
  for (int j = 0; j < 4; ++j) {
    uint32_t i = ransac_pick(seed, h, j, n);
    float4 a = __ldg(src + i), b = __ldg(tgt + i);
    p[j][0] = a.x; p[j][1] = a.y; p[j][2] = a.z;
    q[j][0] = b.x; q[j][1] = b.y; q[j][2] = b.z;
    for (int c = 0; c < 3; ++c) { mp[c] += p[j][c]; mq[c] += q[j][c]; }
  }
  for (int c = 0; c < 3; ++c) { mp[c] *= 0.25; mq[c] *= 0.25; }
  double S[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double acc = 0;
      for (int j = 0; j < 4; ++j) acc += (q[j][r] - mq[r]) * (p[j][c] - mp[c]);
      S[r][c] = acc * 0.25;
    }
  kabsch_rotation(S, R);
  for (int r = 0; r < 3; ++r) t[r] = mq[r] - (R[r][0] * mp[0] + R[r][1] * mp[1] + R[r][2] * mp[2]);
}

// (inliers, sum d^2) -> a key whose unsigned order is open3d's IsBetterRANSACThan: more
// inliers first; at equal count the smaller squared-error sum (= smaller RMSE)
__device__ __forceinline__ unsigned long long ransac_key(uint32_t cnt, float err2) {
  return ((unsigned long long)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - __float_as_uint(err2));
}

__global__ void __launch_bounds__(kRansacThreads)
ransac_eval_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt, uint32_t n, uint64_t seed,
                   uint64_t num_hyp, float max_d2, unsigned long long* __restrict__ blk_key,
                   unsigned long long* __restrict__ blk_hyp) {
  __shared__ float4 s_src[kTile], s_tgt[kTile];
  __shared__ unsigned long long s_key[kRansacThreads / 32], s_hyp[kRansacThreads / 32];
  const int tid = threadIdx.x;
  const uint64_t h0 = ((uint64_t)blockIdx.x * kRansacThreads + tid) * kHyp;

  float R[kHyp][9], t[kHyp][3];
#pragma unroll
  for (int h = 0; h < kHyp; ++h) {
    if (h0 + h < num_hyp) {
      double Rd[3][3], td[3];
      ransac_hypothesis(src, tgt, n, seed, h0 + h, Rd, td);
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[h][3 * r + c] = (float)Rd[r][c];
        t[h][r] = (float)td[r];
      }
    } else {      // past the end: a pose nothing can be an inlier of
      for (int k = 0; k < 9; ++k) R[h][k] = 0.f;
      t[h][0] = t[h][1] = t[h][2] = 3.0e18f;
    }
  }

  uint32_t cnt[kHyp];
  float err[kHyp];
#pragma unroll
  for (int h = 0; h < kHyp; ++h) { cnt[h] = 0; err[h] = 0.f; }

  for (uint32_t base = 0; base < n; base += kTile) {
    const int m = (int)min((uint32_t)kTile, n - base);
    __syncthreads();
    for (int i = tid; i < m; i += kRansacThreads) {
      s_src[i] = __ldg(src + base + i);
      s_tgt[i] = __ldg(tgt + base + i);
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < m; ++i) {
      const float4 a = s_src[i], b = s_tgt[i];
#pragma unroll
      for (int h = 0; h < kHyp; ++h) {
        float dx = fmaf(R[h][0], a.x, fmaf(R[h][1], a.y, fmaf(R[h][2], a.z, t[h][0] - b.x)));
        float dy = fmaf(R[h][3], a.x, fmaf(R[h][4], a.y, fmaf(R[h][5], a.z, t[h][1] - b.y)));
        float dz = fmaf(R[h][6], a.x, fmaf(R[h][7], a.y, fmaf(R[h][8], a.z, t[h][2] - b.z)));
        float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        bool in = d2 < max_d2;
        cnt[h] += in ? 1u : 0u;
        err[h] += in ? d2 : 0.f;
      }
    }
  }

  // best of the thread, the warp, the block; ties go to the lowest hypothesis number
  unsigned long long key = 0, hyp = ~0ull;
#pragma unroll
  for (int h = 0; h < kHyp; ++h) {
    unsigned long long k = cnt[h] ? ransac_key(cnt[h], err[h]) : 0ull;
    if (k > key) { key = k; hyp = h0 + h; }
  }
  for (int d = 16; d > 0; d >>= 1) {
    unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, d);
    unsigned long long oh = __shfl_xor_sync(0xffffffffu, hyp, d);
    if (ok > key || (ok == key && oh < hyp)) { key = ok; hyp = oh; }
  }
  if ((tid & 31) == 0) { s_key[tid >> 5] = key; s_hyp[tid >> 5] = hyp; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kRansacThreads / 32; ++w)
      if (s_key[w] > key || (s_key[w] == key && s_hyp[w] < hyp)) { key = s_key[w]; hyp = s_hyp[w]; }
    blk_key[blockIdx.x] = key;
    blk_hyp[blockIdx.x] = hyp;
  }
}

// winner over the blocks, its pose regenerated in fp64 and re-scored in fp64 for the report
__global__ void __launch_bounds__(1024)
ransac_final_kernel(const float4* __restrict__ src, const float4* __restrict__ tgt, uint32_t n, uint64_t seed,
                    const unsigned long long* __restrict__ blk_key, const unsigned long long* __restrict__ blk_hyp,
                    uint32_t n_blk, double max_dist, double* __restrict__ result) {
  __shared__ unsigned long long s_key[32], s_hyp[32];
  __shared__ double s_T[12], s_red[32][2];
  const int tid = threadIdx.x;
  unsigned long long key = 0, hyp = ~0ull;
  for (uint32_t b = tid; b < n_blk; b += blockDim.x) {
    unsigned long long k = blk_key[b], h = blk_hyp[b];
    if (k > key || (k == key && h < hyp)) { key = k; hyp = h; }
  }
  for (int d = 16; d > 0; d >>= 1) {
    unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, d);
    unsigned long long oh = __shfl_xor_sync(0xffffffffu, hyp, d);
    if (ok > key || (ok == key && oh < hyp)) { key = ok; hyp = oh; }
  }
  if ((tid & 31) == 0) { s_key[tid >> 5] = key; s_hyp[tid >> 5] = hyp; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (s_key[w] > key || (s_key[w] == key && s_hyp[w] < hyp)) { key = s_key[w]; hyp = s_hyp[w]; }
    s_key[0] = key;
    s_hyp[0] = hyp;
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, t[3] = {0, 0, 0};
    if (key != 0) ransac_hypothesis(src, tgt, n, seed, hyp, R, t);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) s_T[4 * r + c] = R[r][c];
      s_T[4 * r + 3] = t[r];
    }
  }
  __syncthreads();
  key = s_key[0];
  hyp = s_hyp[0];
  double cnt = 0, err = 0;
  if (key != 0) {
    for (uint32_t i = tid; i < n; i += blockDim.x) {
      float4 a = src[i], b = tgt[i];
      double dx = s_T[0] * a.x + s_T[1] * a.y + s_T[2] * a.z + s_T[3] - b.x;
      double dy = s_T[4] * a.x + s_T[5] * a.y + s_T[6] * a.z + s_T[7] - b.y;
      double dz = s_T[8] * a.x + s_T[9] * a.y + s_T[10] * a.z + s_T[11] - b.z;
      double d2 = dx * dx + dy * dy + dz * dz;
      if (sqrt(d2) < max_dist) { cnt += 1.0; err += d2; }
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    err += __shfl_xor_sync(0xffffffffu, err, d);
  }
  if ((tid & 31) == 0) { s_red[tid >> 5][0] = cnt; s_red[tid >> 5][1] = err; }
  __syncthreads();
  if (tid == 0) {
    cnt = 0; err = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { cnt += s_red[w][0]; err += s_red[w][1]; }
    for (int k = 0; k < 12; ++k) result[k] = s_T[k];
    result[12] = 0; result[13] = 0; result[14] = 0; result[15] = 1;
    result[16] = n ? cnt / (double)n : 0.0;                 // fitness
    result[17] = cnt > 0 ? sqrt(err / cnt) : 0.0;           // inlier RMSE
    result[18] = key != 0 ? (double)hyp : -1.0;             // winning hypothesis
    result[19] = (double)(uint32_t)(key >> 32);             // its fp32 inlier count
  }
}

inline uint32_t ransac_blocks(int64_t num_hyp) {
  const int64_t per_block = (int64_t)kRansacThreads * kHyp;
  return (uint32_t)((num_hyp + per_block - 1) / per_block);
}

}  // namespace

extern "C" {

int32_t dgr_ransac_ws_elems(int64_t n_corr, int64_t num_hyp, int64_t* n_elems) {
  DGR_ARG_CHECK(n_elems != nullptr && n_corr >= 0 && num_hyp >= 0, "bad arguments");
  *n_elems = 4 * n_corr + 2 * (int64_t)ransac_blocks(num_hyp) + 2;
  return DGR_OK;
}

int32_t dgr_ransac_correspondence(const float* x, const float* y, const int32_t* idx0, const int32_t* idx1,
                                  int64_t n_corr, double max_dist, int64_t num_hyp, uint64_t seed, uint64_t* ws,
                                  double* result, void* stream) {
  DGR_ARG_CHECK(x != nullptr && y != nullptr && ws != nullptr && result != nullptr, "null pointer");
  DGR_ARG_CHECK(n_corr >= 1 && n_corr < (1ll << 31), "correspondence count out of range");
  DGR_ARG_CHECK(num_hyp >= 1 && num_hyp <= (1ll << 40), "hypothesis count out of range");
  DGR_ARG_CHECK(max_dist > 0, "max_dist must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  float4* src = reinterpret_cast<float4*>(ws);
  float4* tgt = src + n_corr;
  unsigned long long* blk_key = reinterpret_cast<unsigned long long*>(tgt + n_corr);
  const uint32_t n_blk = ransac_blocks(num_hyp);
  unsigned long long* blk_hyp = blk_key + n_blk;
  ransac_pack_kernel<<<dgr_blocks(n_corr, 256), 256, 0, st>>>(x, y, idx0, idx1, n_corr, src, tgt);
  ransac_eval_kernel<<<n_blk, kRansacThreads, 0, st>>>(src, tgt, (uint32_t)n_corr, seed, (uint64_t)num_hyp,
                                                       (float)(max_dist * max_dist), blk_key, blk_hyp);
  ransac_final_kernel<<<1, 1024, 0, st>>>(src, tgt, (uint32_t)n_corr, seed, blk_key, blk_hyp, n_blk, max_dist,
                                          result);
  dgr_note_launches(3);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

}  // extern "C"
