// Correspondence post-processing and pose estimation:
//   dgr_inlier_coords      6-D coordinate assembly   (core/deep_global_registration.py:261-262)
//   dgr_sigmoid_clip_sum   inlier weights + gate sum (core/deep_global_registration.py:269-272)
//   dgr_se3_register       weighted Procrustes (core/registration.py:91-113) followed by the
//                          robust SE(3) refinement (core/registration.py:135-194) with the loss
//                          of core/loss.py:42-61.
//
// The reference runs the refinement as <=1000 PyTorch iterations of ~15 tiny kernels and
// three .item() host syncs each.  Here the whole optimisation is ONE launch: a thread-block
// cluster of 8 CTAs keeps the active (non-zero weight) correspondences resident in its
// 8 x ~200 KB of shared memory, every iteration reduces the 13 loss/gradient moments with
// warp shuffles, exchanges the per-CTA partials through distributed shared memory, and
// each CTA redundantly applies the (deterministic) Gram-Schmidt backward pass, the Adam
// update and the reference's stopping rule - no host round trip, no global memory traffic
// after the prologue.  The 3x3 SVD is a one-sided Jacobi iteration in fp64 on one thread.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kabsch.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kClusterSize = 8;
constexpr int kRegThreads = 512;
constexpr int kMaxVals = 16;
constexpr int kSmemPoints = 7168;   // per CTA: 7 floats * 7168 = 196 KB

__global__ void inlier_coords_kernel(const int32_t* __restrict__ c0, const int32_t* __restrict__ c1,
                                     const int32_t* __restrict__ idx1, int64_t n0,
                                     int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0) return;
  int4 a = reinterpret_cast<const int4*>(c0)[i];
  int4 b = reinterpret_cast<const int4*>(c1)[idx1[i]];
  int32_t* o = out + i * 7;
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.y; o[5] = b.z; o[6] = b.w;
}

__global__ void sigmoid_clip_sum_kernel(const float* __restrict__ logit, int64_t n, float clip,
                                        float* __restrict__ w, double* wsum) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s = 1.f / (1.f + expf(-logit[i]));
    if (clip > 0.f && s < clip) s = 0.f;
    w[i] = s;
    acc += (double)s;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += ws[k];
    atomicAdd(wsum, t);
  }
}

// gather the correspondences into structure-of-arrays form: 7 arrays of length n
__global__ void pack_corr_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                 const int32_t* __restrict__ idx1, const float* __restrict__ w,
                                 int64_t n, float* __restrict__ pack) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t j = idx1 != nullptr ? (int64_t)idx1[i] : i;
  pack[0 * n + i] = x[3 * i + 0];
  pack[1 * n + i] = x[3 * i + 1];
  pack[2 * n + i] = x[3 * i + 2];
  pack[3 * n + i] = y[3 * j + 0];
  pack[4 * n + i] = y[3 * j + 1];
  pack[5 * n + i] = y[3 * j + 2];
  pack[6 * n + i] = w[i];
}

// ---------------------------------------------------------------------------------------
// rot6d -> R (ortho2rotation, core/registration.py:16-64) and its backward pass
// ---------------------------------------------------------------------------------------
struct Rot6dCache {
  float x[3], y[3], yr[3], nx, nu, ip, n2, f;
  bool clamp_x, clamp_n2, clamp_u;
};

__device__ void rot6d_forward(const float p[6], float R[9], Rot6dCache& c) {
  float n = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  c.clamp_x = n < 1e-8f;
  c.nx = fmaxf(n, 1e-8f);
  for (int k = 0; k < 3; ++k) { c.x[k] = p[k] / c.nx; c.yr[k] = p[3 + k]; }
  c.ip = c.x[0] * c.yr[0] + c.x[1] * c.yr[1] + c.x[2] * c.yr[2];
  float n2 = c.x[0] * c.x[0] + c.x[1] * c.x[1] + c.x[2] * c.x[2];
  c.clamp_n2 = n2 < 1e-8f;
  c.n2 = fmaxf(n2, 1e-8f);
  c.f = c.ip / c.n2;
  float u[3];
  for (int k = 0; k < 3; ++k) u[k] = c.yr[k] - c.f * c.x[k];
  float nu = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  c.clamp_u = nu < 1e-8f;
  c.nu = fmaxf(nu, 1e-8f);
  for (int k = 0; k < 3; ++k) c.y[k] = u[k] / c.nu;
  float z[3] = {c.x[1] * c.y[2] - c.x[2] * c.y[1], c.x[2] * c.y[0] - c.x[0] * c.y[2],
                c.x[0] * c.y[1] - c.x[1] * c.y[0]};
  for (int k = 0; k < 3; ++k) { R[3 * k + 0] = c.x[k]; R[3 * k + 1] = c.y[k]; R[3 * k + 2] = z[k]; }
}

// G = dL/dR row-major; out g[6] = dL/d rot6d
__device__ void rot6d_backward(const Rot6dCache& c, const float G[9], float g[6]) {
  float gx[3], gy[3], gz[3];
  for (int k = 0; k < 3; ++k) { gx[k] = G[3 * k + 0]; gy[k] = G[3 * k + 1]; gz[k] = G[3 * k + 2]; }
  // z = x cross y
  float gxt[3] = {gx[0] + (c.y[1] * gz[2] - c.y[2] * gz[1]), gx[1] + (c.y[2] * gz[0] - c.y[0] * gz[2]),
                  gx[2] + (c.y[0] * gz[1] - c.y[1] * gz[0])};
  float gyt[3] = {gy[0] + (gz[1] * c.x[2] - gz[2] * c.x[1]), gy[1] + (gz[2] * c.x[0] - gz[0] * c.x[2]),
                  gy[2] + (gz[0] * c.x[1] - gz[1] * c.x[0])};
  // y = u / max(|u|, 1e-8)
  float gu[3];
  if (c.clamp_u) {
    for (int k = 0; k < 3; ++k) gu[k] = gyt[k] / c.nu;
  } else {
    float d = c.y[0] * gyt[0] + c.y[1] * gyt[1] + c.y[2] * gyt[2];
    for (int k = 0; k < 3; ++k) gu[k] = (gyt[k] - c.y[k] * d) / c.nu;
  }
  // u = y_raw - f x
  float gyr[3] = {gu[0], gu[1], gu[2]};
  float gf = -(gu[0] * c.x[0] + gu[1] * c.x[1] + gu[2] * c.x[2]);
  for (int k = 0; k < 3; ++k) gxt[k] -= c.f * gu[k];
  // f = ip / max(n2, 1e-8)
  float gip = gf / c.n2;
  float gn2 = c.clamp_n2 ? 0.f : -gf * c.ip / (c.n2 * c.n2);
  for (int k = 0; k < 3; ++k) {
    gxt[k] += gip * c.yr[k] + 2.f * gn2 * c.x[k];
    gyr[k] += gip * c.x[k];
  }
  // x = x_raw / max(|x_raw|, 1e-8)
  if (c.clamp_x) {
    for (int k = 0; k < 3; ++k) g[k] = gxt[k] / c.nx;
  } else {
    float d = c.x[0] * gxt[0] + c.x[1] * gxt[1] + c.x[2] * gxt[2];
    for (int k = 0; k < 3; ++k) g[k] = (gxt[k] - c.x[k] * d) / c.nx;
  }
  for (int k = 0; k < 3; ++k) g[3 + k] = gyr[k];
}

// ---------------------------------------------------------------------------------------
// the cluster kernel
// ---------------------------------------------------------------------------------------
struct RegShared {
  double slots[2][kClusterSize][kMaxVals];
  double warp_part[kRegThreads / 32][kMaxVals];
  double tot[kMaxVals];
  float params[16];   // R (9) + t (3)
  int counts[kClusterSize];
  int flags[4];
};

// Sum `nv` per-thread doubles over the whole cluster.  Result in sh.tot[0..nv) of every CTA.
template <int NV>
__device__ __forceinline__ void cluster_allreduce(cg::cluster_group& cluster, RegShared& sh,
                                                  double (&v)[NV], int& parity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if (lane == 0) sh.warp_part[warp][k] = x;
  }
  __syncthreads();
  const unsigned rank = cluster.block_rank();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < kRegThreads / 32; ++w) s += sh.warp_part[w][threadIdx.x];
    for (unsigned r = 0; r < kClusterSize; ++r) {
      RegShared* remote = cluster.map_shared_rank(&sh, r);
      remote->slots[parity][rank][threadIdx.x] = s;
    }
  }
  cluster.sync();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int r = 0; r < kClusterSize; ++r) s += sh.slots[parity][r][threadIdx.x];
    sh.tot[threadIdx.x] = s;
  }
  __syncthreads();
  parity ^= 1;
}

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kRegThreads, 1)
se3_register_kernel(const float* __restrict__ pack, int64_t n, float q, int max_iter,
                    int max_break_count, float break_ratio, float lr0, float gamma, float eps,
                    float* __restrict__ result) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RegShared& sh = *reinterpret_cast<RegShared*>(smem_raw);
  float* pts = reinterpret_cast<float*>(smem_raw + ((sizeof(RegShared) + 15) / 16) * 16);
  const int tid = threadIdx.x;
  const unsigned rank = cluster.block_rank();
  int parity = 0;

  // ---- prologue: this CTA's slice, active correspondences compacted into shared memory ----
  const int64_t per = (n + kClusterSize - 1) / kClusterSize;
  const int64_t lo = min(n, (int64_t)rank * per), hi = min(n, lo + per);
  __shared__ int s_count;
  if (tid == 0) s_count = 0;
  __syncthreads();
  // ordered compaction, 512 candidates per round
  for (int64_t base = lo; base < hi; base += kRegThreads) {
    int64_t i = base + tid;
    float wv = (i < hi) ? pack[6 * n + i] : 0.f;
    int act = (wv != 0.f) ? 1 : 0;
    // block exclusive scan of `act`
    int inc = act;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    __shared__ int wsum[kRegThreads / 32];
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
    for (int k = 0; k < kRegThreads / 32; ++k) {
      int s = wsum[k];
      if (k < warp) wbase += s;
      tot += s;
    }
    int pos = s_count + wbase + inc - act;
    if (act && pos < kSmemPoints) {
#pragma unroll
      for (int a = 0; a < 7; ++a) pts[a * kSmemPoints + pos] = pack[a * n + i];
    }
    __syncthreads();
    if (tid == 0) s_count += tot;
    __syncthreads();
  }
  const int m_local = s_count;
  // agree on the mode: resident (all slices fit) or streaming from global memory
  if (tid == 0) {
    for (unsigned r = 0; r < kClusterSize; ++r) cluster.map_shared_rank(&sh, r)->counts[rank] = m_local;
  }
  cluster.sync();
  bool resident = true;
  int m_total = 0;
  for (int r = 0; r < kClusterSize; ++r) {
    resident = resident && (sh.counts[r] <= kSmemPoints);
    m_total += sh.counts[r];
  }
  const int64_t cnt = resident ? (int64_t)m_local : (hi - lo);
  auto load_pt = [&](int64_t k, float (&x)[3], float (&y)[3], float& w) {
    if (resident) {
      x[0] = pts[0 * kSmemPoints + k]; x[1] = pts[1 * kSmemPoints + k]; x[2] = pts[2 * kSmemPoints + k];
      y[0] = pts[3 * kSmemPoints + k]; y[1] = pts[4 * kSmemPoints + k]; y[2] = pts[5 * kSmemPoints + k];
      w = pts[6 * kSmemPoints + k];
    } else {
      const int64_t i = lo + k;
      x[0] = pack[0 * n + i]; x[1] = pack[1 * n + i]; x[2] = pack[2 * n + i];
      y[0] = pack[3 * n + i]; y[1] = pack[4 * n + i]; y[2] = pack[5 * n + i];
      w = pack[6 * n + i];
    }
  };

  // ---- weighted Procrustes: first moments -------------------------------------------------
  {
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    float a[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t k = tid; k < cnt; k += kRegThreads) {
      float x[3], y[3], w;
      load_pt(k, x, y, w);
      a[0] += fabsf(w);
      a[1] += w * x[0]; a[2] += w * x[1]; a[3] += w * x[2];
      a[4] += w * y[0]; a[5] += w * y[1]; a[6] += w * y[2];
    }
    for (int k = 0; k < 7; ++k) v[k] = (double)a[k];
    cluster_allreduce<7>(cluster, sh, v, parity);
  }
  const double W1 = sh.tot[0];
  const float wden = (float)W1 + eps;
  float mux[3], muy[3];
  for (int k = 0; k < 3; ++k) {
    mux[k] = (float)(sh.tot[1 + k] / (double)wden);
    muy[k] = (float)(sh.tot[4 + k] / (double)wden);
  }
  __syncthreads();
  // ---- second moments Sxy = sum wn (y - muy)(x - mux)^T -------------------------------------
  {
    float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t k = tid; k < cnt; k += kRegThreads) {
      float x[3], y[3], w;
      load_pt(k, x, y, w);
      const float wn = w / wden;
      const float dx[3] = {wn * (x[0] - mux[0]), wn * (x[1] - mux[1]), wn * (x[2] - mux[2])};
      const float dy[3] = {y[0] - muy[0], y[1] - muy[1], y[2] - muy[2]};
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) a[3 * r + c] += dy[r] * dx[c];
    }
    double v[9];
    for (int k = 0; k < 9; ++k) v[k] = (double)a[k];
    cluster_allreduce<9>(cluster, sh, v, parity);
  }
  if (tid == 0) {
    double S[3][3], R[3][3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) S[r][c] = (double)(float)sh.tot[3 * r + c];   // fp32 Sxy, as the reference
    if (m_total > 0) {
      kabsch_rotation(S, R);
    } else {
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r][c] = (r == c);
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) sh.params[3 * r + c] = (float)R[r][c];
    for (int r = 0; r < 3; ++r) {
      float rm = sh.params[3 * r + 0] * mux[0] + sh.params[3 * r + 1] * mux[1] + sh.params[3 * r + 2] * mux[2];
      sh.params[9 + r] = muy[r] - rm;
    }
  }
  __syncthreads();

  // ---- refinement state (replicated in thread 0 of every CTA) --------------------------------
  float p6[6], tr[3], m1[9], m2[9];
  double lr = lr0, b1t = 1.0, b2t = 1.0;
  float loss_prev = 0.f, loss = 0.f;
  int breaks = 0, iters = 0;
  Rot6dCache cache;
  if (tid == 0) {
    for (int k = 0; k < 3; ++k) { p6[k] = sh.params[3 * k + 0]; p6[3 + k] = sh.params[3 * k + 1]; tr[k] = sh.params[9 + k]; }
    for (int k = 0; k < 9; ++k) m1[k] = m2[k] = 0.f;
  }
  const float invq = 1.f / q;
  for (int it = 0; it < max_iter; ++it) {
    iters = it;
    if (tid == 0) {
      float R[9];
      rot6d_forward(p6, R, cache);
      for (int k = 0; k < 9; ++k) sh.params[k] = R[k];
      for (int k = 0; k < 3; ++k) sh.params[9 + k] = tr[k];
    }
    __syncthreads();
    float R[9], t[3];
    for (int k = 0; k < 9; ++k) R[k] = sh.params[k];
    for (int k = 0; k < 3; ++k) t[k] = sh.params[9 + k];
    float a[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) a[k] = 0.f;
    for (int64_t k = tid; k < cnt; k += kRegThreads) {
      float x[3], y[3], w;
      load_pt(k, x, y, w);
      float r[3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        r[c] = ((R[3 * c + 0] * x[0] + R[3 * c + 1] * x[1] + R[3 * c + 2] * x[2] + t[c]) - y[c]) / q;
      const float s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      float rho, drho;   // rho(s), d rho / d s
      if (s < 1.f) {
        rho = 0.5f * s;
        drho = 0.5f;
      } else {
        const float rt = sqrtf(s + eps);
        rho = 0.5f * (rt - 0.5f);
        drho = 0.25f / rt;
      }
      a[0] += w * rho;
      const float gs = w * drho * 2.f * invq;
      const float g[3] = {gs * r[0], gs * r[1], gs * r[2]};
      a[1] += g[0]; a[2] += g[1]; a[3] += g[2];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        a[4 + 3 * c + 0] += g[c] * x[0];
        a[4 + 3 * c + 1] += g[c] * x[1];
        a[4 + 3 * c + 2] += g[c] * x[2];
      }
    }
    double v[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) v[k] = (double)a[k];
    cluster_allreduce<13>(cluster, sh, v, parity);
    // every thread evaluates the stop rule on identical data so the cluster stays in lock step
    const double w1 = W1;   // weights are >= 0: sum w == sum |w|
    loss = (float)(sh.tot[0] / w1);
    if (it == 0) loss_prev = loss;   // the reference evaluates loss_prev on the initial pose
    if (loss < 1e-7f) break;
    if (tid == 0) {
      float G[9], grad[9];
      for (int k = 0; k < 9; ++k) G[k] = (float)(sh.tot[4 + k] / w1);
      rot6d_backward(cache, G, grad);
      for (int k = 0; k < 3; ++k) grad[6 + k] = (float)(sh.tot[1 + k] / w1);
      b1t *= 0.9;
      b2t *= 0.999;
      const float step_size = (float)(lr / (1.0 - b1t));
      const float bc2_sqrt = (float)sqrt(1.0 - b2t);
      for (int k = 0; k < 9; ++k) {
        m1[k] = m1[k] + (grad[k] - m1[k]) * 0.1f;
        m2[k] = m2[k] * 0.999f + grad[k] * grad[k] * 0.001f;
        const float denom = sqrtf(m2[k]) / bc2_sqrt + 1e-8f;
        const float upd = step_size * (m1[k] / denom);
        if (k < 6) p6[k] -= upd; else tr[k - 6] -= upd;
      }
      lr *= (double)gamma;
    }
    bool stop = false;
    if (fabsf(loss_prev - loss) < loss_prev * break_ratio) {
      ++breaks;
      if (breaks >= max_break_count) stop = true;
    }
    loss_prev = loss;
    if (stop) break;
  }
  if (rank == 0 && tid == 0) {
    float R[9];
    if (max_iter > 0) {
      rot6d_forward(p6, R, cache);
    } else {
      for (int k = 0; k < 9; ++k) R[k] = sh.params[k];
      for (int k = 0; k < 3; ++k) tr[k] = sh.params[9 + k];
    }
    for (int k = 0; k < 9; ++k) result[k] = R[k];
    for (int k = 0; k < 3; ++k) result[9 + k] = tr[k];
    result[12] = (float)iters;
    result[13] = loss;
    result[14] = (float)breaks;
    result[15] = (float)m_total;
  }
  cluster.sync();   // nobody exits while peers may still write into its shared memory
}


// ---------------------------------------------------------------------------------------
// point-to-point ICP (open3d registration_icp as called at core/deep_global_registration.py:
// 317-322): nearest target point within a radius through the target's voxel hash, Kabsch
// update in fp64, open3d's convergence rule.  No host round trip: all max_iter + 1
// (match, update) launch pairs are enqueued up front and turn into no-ops once the
// device-side `done` flag is set.
// ---------------------------------------------------------------------------------------
struct IcpState {
  double T[12];          // current pose, row-major [R | t]
  double sums[17];       // n, sum d2, sum p (3), sum q (3), sum q p^T (9)
  double prev_fitness, prev_rmse, fitness, rmse;
  int iteration, done;
};

__global__ void icp_init_kernel(const double* __restrict__ T_init, IcpState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int k = 0; k < 12; ++k) st->T[k] = T_init[k];
    for (int k = 0; k < 17; ++k) st->sums[k] = 0.0;
    st->prev_fitness = st->prev_rmse = st->fitness = st->rmse = 0.0;
    st->iteration = 0;
    st->done = 0;
  }
}

__global__ void __launch_bounds__(256)
icp_match_kernel(const float* __restrict__ src, int64_t n_src, const float* __restrict__ tgt,
                 const dgr_keyspec_t* __restrict__ spec_p, const uint64_t* __restrict__ keys,
                 const int32_t* __restrict__ vals, uint64_t mask, int32_t batch, double voxel, double max_dist,
                 IcpState* st) {
  if (st->done) return;
  const dgr_keyspec_t s = *spec_p;
  double T[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) T[k] = st->T[k];
  double acc[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) acc[k] = 0.0;
  const int reach = (int)ceil(max_dist / voxel);     // cells to search on every side (2 for DGR)
  const int side = 2 * reach + 1, n_cells = side * side * side;
  // 8 lanes share one source point: the (2 reach + 1)^3 = 125 candidate cells are probed 8 at a time (a thread
  // per point walked them serially - 125 dependent L2 round trips - and left the SMs 92 % idle), then the lanes'
  // best (d2, then lower row) is reduced with shuffles; lane 0 of the group accumulates the moments
  const int sub = threadIdx.x & 7;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) >> 3;
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i0 < ((n_src + groups - 1) / groups) * groups;
       i0 += groups) {
    const bool have = i0 < n_src;                      // whole warps stay in the loop for the shuffles
    const int64_t i = have ? i0 : 0;
    const double x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
    const double p[3] = {T[0] * x + T[1] * y + T[2] * z + T[3], T[4] * x + T[5] * y + T[6] * z + T[7],
                         T[8] * x + T[9] * y + T[10] * z + T[11]};
    int cell[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) cell[a] = (int)floor(p[a] / voxel);
    double best = max_dist * max_dist;
    int best_j = -1;
    for (int c = sub; c < n_cells && have; c += 8) {
      const int dx = c % side - reach, dy = (c / side) % side - reach, dz = c / (side * side) - reach;
      const int32_t row[4] = {batch, cell[0] + dx, cell[1] + dy, cell[2] + dz};
      bool inside = true;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const long long d = (long long)row[q] - s.lo[q];
        inside = inside && d >= 0 && d < (1ll << s.bits[q]);
      }
      if (!inside) continue;
      const int32_t j = dgr_hash_lookup(keys, vals, mask, dgr_pack_key(row, s));
      if (j < 0) continue;
      const double ex = p[0] - tgt[3 * (int64_t)j], ey = p[1] - tgt[3 * (int64_t)j + 1],
                   ez = p[2] - tgt[3 * (int64_t)j + 2];
      const double d2 = ex * ex + ey * ey + ez * ez;
      if (d2 < best || (d2 == best && (best_j < 0 || j < best_j))) {
        best = d2;
        best_j = j;
      }
    }
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {                  // nearest, lower row index on ties: order-independent
      const double ob = __shfl_xor_sync(0xffffffffu, best, d);
      const int oj = __shfl_xor_sync(0xffffffffu, best_j, d);
      if (oj >= 0 && (best_j < 0 || ob < best || (ob == best && oj < best_j))) {
        best = ob;
        best_j = oj;
      }
    }
    if (have && sub == 0 && best_j >= 0) {
      const double q[3] = {tgt[3 * (int64_t)best_j], tgt[3 * (int64_t)best_j + 1], tgt[3 * (int64_t)best_j + 2]};
      acc[0] += 1.0;
      acc[1] += best;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        acc[2 + a] += p[a];
        acc[5 + a] += q[a];
#pragma unroll
        for (int b = 0; b < 3; ++b) acc[8 + 3 * a + b] += q[a] * p[b];
      }
    }
  }
  __shared__ double red[8][17];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 17; ++k) {
    double v = acc[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 17) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    if (v != 0.0) atomicAdd(&st->sums[threadIdx.x], v);
  }
}

__global__ void icp_update_kernel(IcpState* st, int64_t n_src, int max_iter, double rel_fitness, double rel_rmse,
                                  double* __restrict__ result) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (!st->done) {
    const double n = st->sums[0];
    const double fitness = n_src > 0 ? n / (double)n_src : 0.0;
    const double rmse = n > 0 ? sqrt(st->sums[1] / n) : 0.0;
    const int k = st->iteration;
    bool stop = false;
    if (k > 0 && fabs(st->prev_fitness - fitness) < rel_fitness && fabs(st->prev_rmse - rmse) < rel_rmse) stop = true;
    if (k >= max_iter) stop = true;
    st->fitness = fitness;
    st->rmse = rmse;
    if (stop) {
      st->done = 1;
    } else {
      double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, t[3] = {0, 0, 0};
      if (n > 0) {
        double mp[3], mq[3], S[3][3];
        for (int a = 0; a < 3; ++a) { mp[a] = st->sums[2 + a] / n; mq[a] = st->sums[5 + a] / n; }
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) S[a][b] = st->sums[8 + 3 * a + b] / n - mq[a] * mp[b];
        kabsch_rotation(S, R);
        for (int a = 0; a < 3; ++a) t[a] = mq[a] - (R[a][0] * mp[0] + R[a][1] * mp[1] + R[a][2] * mp[2]);
      }
      double Tn[12];
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 4; ++b) {
          double v = R[a][0] * st->T[b] + R[a][1] * st->T[4 + b] + R[a][2] * st->T[8 + b];
          if (b == 3) v += t[a];
          Tn[4 * a + b] = v;
        }
      }
      for (int q = 0; q < 12; ++q) st->T[q] = Tn[q];
      st->prev_fitness = fitness;
      st->prev_rmse = rmse;
      st->iteration = k + 1;
      for (int q = 0; q < 17; ++q) st->sums[q] = 0.0;
    }
  }
  for (int q = 0; q < 12; ++q) result[q] = st->T[q];
  result[12] = 0.0; result[13] = 0.0; result[14] = 0.0; result[15] = 1.0;
  result[16] = st->fitness;
  result[17] = st->rmse;
  result[18] = (double)st->iteration;
  result[19] = st->sums[0];
}

}  // namespace

extern "C" {

int32_t dgr_inlier_coords(const int32_t* coords0, const int32_t* coords1, const int32_t* idx1,
                          int64_t n0, int32_t* out, void* stream) {
  if (n0 == 0) return DGR_OK;
  inlier_coords_kernel<<<dgr_blocks(n0, 256), 256, 0, (cudaStream_t)stream>>>(coords0, coords1, idx1, n0,
                                                                           out);
  dgr_note_launches(1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

int32_t dgr_sigmoid_clip_sum(const float* logit, int64_t n, float clip, float* w, double* wsum,
                             void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DGR_CUDA_CHECK(cudaMemsetAsync(wsum, 0, sizeof(double), st));
  if (n == 0) return DGR_OK;
  unsigned blocks = dgr_blocks(n, 256 * 4);
  if (blocks > 592) blocks = 592;
  sigmoid_clip_sum_kernel<<<blocks, 256, 0, st>>>(logit, n, clip, w, wsum);
  dgr_note_launches(1);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

int32_t dgr_se3_register(const float* x, const float* y, const int32_t* idx1, const float* w,
                         int64_t n, float quantization_size, int32_t max_iter,
                         int32_t max_break_count, float break_threshold_ratio, float lr, float gamma,
                         float* pack_ws, int32_t* cnt_ws, float* result, void* stream) {
  (void)cnt_ws;
  DGR_ARG_CHECK(n >= 1, "need at least one correspondence");
  DGR_ARG_CHECK(quantization_size > 0, "quantization_size must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  pack_corr_kernel<<<dgr_blocks(n, 256), 256, 0, st>>>(x, y, idx1, w, n, pack_ws);
  const size_t smem = ((sizeof(RegShared) + 15) / 16) * 16 + (size_t)7 * kSmemPoints * sizeof(float);
  DGR_ENSURE_SMEM(se3_register_kernel, smem);
  const float eps = 1.1920928955078125e-07f;   // np.finfo(np.float32).eps, core/loss.py:44
  se3_register_kernel<<<kClusterSize, kRegThreads, smem, st>>>(pack_ws, n, quantization_size, max_iter,
                                                              max_break_count, break_threshold_ratio,
                                                              lr, gamma, eps, result);
  dgr_note_launches(2);
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

// Point-to-point ICP with open3d's defaults (what registration_icp(source, target,
// max_correspondence_distance, init) does at core/deep_global_registration.py:317-322).
// Nearest neighbours come from the TARGET's voxel hash (keys / vals / spec of the table
// dgr_unique_first built for the target cloud at `voxel`, rows = rows of tgt).
// T_init: device double[12] row-major [R | t]; state_ws: 64 doubles; result: device double[20] =
// 4x4 pose (16), fitness, inlier rmse, iterations, correspondences of the last evaluation.
int32_t dgr_icp_point_to_point(const float* src, int64_t n_src, const float* tgt, const dgr_keyspec_t* spec,
                               const uint64_t* keys, const int32_t* vals, int64_t cap, int32_t batch,
                               double voxel, double max_dist, const double* T_init, int32_t max_iter,
                               double rel_fitness, double rel_rmse, double* state_ws, double* result,
                               void* stream) {
  DGR_ARG_CHECK(cap > 0 && (cap & (cap - 1)) == 0, "capacity must be a power of two");
  DGR_ARG_CHECK(voxel > 0 && max_dist > 0 && max_iter >= 0, "bad ICP parameters");
  DGR_ARG_CHECK(max_dist / voxel <= 4.0, "search radius above 4 voxels is not supported");
  static_assert(sizeof(IcpState) <= 64 * sizeof(double), "state workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  IcpState* state = reinterpret_cast<IcpState*>(state_ws);
  icp_init_kernel<<<1, 32, 0, st>>>(T_init, state);
  unsigned blocks = dgr_blocks(n_src * 8, 256);      // 8 lanes per source point
  if (blocks > 2368) blocks = 2368;
  for (int k = 0; k <= max_iter; ++k) {
    icp_match_kernel<<<blocks, 256, 0, st>>>(src, n_src, tgt, spec, keys, vals, (uint64_t)cap - 1, batch, voxel,
                                             max_dist, state);
    icp_update_kernel<<<1, 32, 0, st>>>(state, n_src, max_iter, rel_fitness, rel_rmse, result);
  }
  dgr_note_launches(1 + 2 * (max_iter + 1));
  DGR_LAUNCH_CHECK();
  return DGR_OK;
}

}  // extern "C"
