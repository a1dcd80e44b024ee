// Stand-alone trajectory generation (K4 of SURVEY.md §2.1): pos / vel [B, T, dof] to HBM.
//
// Replaces traj_gen.get_traj_pos() / get_traj_vel() (black_box_wrapper.py:117-118).  The kernel is
// bound by HBM *store* bandwidth (8 bytes per (t,dof) element), so the layout work is all about
// the stores: one warp owns one env, lanes own consecutive time points, each lane's dof values
// are transposed through a per-warp shared-memory tile and leave as fully coalesced 128-byte
// warp stores.
#pragma once
#include "fg_device.cuh"

namespace fg {

constexpr int kTrajThreads = 256;
constexpr int kTrajWarps = kTrajThreads / 32;

// closed-form MPs: ProMP (KW = K) and ProDMP (KW = K+3); KW == 0 selects the run-time-K fallback
template <int MP, int N, int KW>
__global__ void __launch_bounds__(kTrajThreads)
k_trajgen_closed(const __grid_constant__ DevCfg c, const float* __restrict__ params, const float* __restrict__ bc_pos,
                 const float* __restrict__ bc_vel, float* __restrict__ pos_out, float* __restrict__ vel_out,
                 const long long B) {
  extern __shared__ float smem[];
  const int T = c.T, K = c.K;
  const int kw = (KW > 0) ? KW : c.cols_a;
  float* tabA = smem;                          // [T, kw]
  float* tabB = tabA + T * kw;                 // ProMP: [T-1] time increments; ProDMP: [T, kw]
  float* tile = tabB + c.rows_b * c.cols_b;    // [warps][2][32*N]
  float* wgen = tile + kTrajWarps * 2 * 32 * N;   // fallback only: [warps][N*kw]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < T * kw; i += kTrajThreads) tabA[i] = c.tab_a[i];
  for (int i = tid; i < c.rows_b * c.cols_b; i += kTrajThreads) tabB[i] = c.tab_b[i];
  __syncthreads();

  float* tp = tile + warp * 2 * 32 * N;
  float* tv = tp + 32 * N;
  const int KP = (MP == FG_MP_PROMP) ? K : K + 1;
  const long long warps_total = (long long)gridDim.x * kTrajWarps;
  for (long long b = (long long)blockIdx.x * kTrajWarps + warp; b < B; b += warps_total) {
    // per-env weight vector, identical in every lane (broadcast loads)
    float w[N][(KW > 0) ? KW : 1];
    float* wg = wgen + warp * N * kw;
    if constexpr (MP == FG_MP_PROMP) {
#pragma unroll
      for (int d = 0; d < N; ++d)
        for (int k = 0; k < kw; ++k) {
          const float x = params[b * N * KP + d * KP + k];
          if constexpr (KW > 0) w[d][k] = x; else if (lane == 0) wg[d * kw + k] = x;
        }
    } else {
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const float yb = bc_pos[b * N + d];
        const float vb = __fmul_rn(bc_vel[b * N + d], c.tau);
        for (int k = 0; k < kw; ++k) {
          float x;
          if (k == 0) x = yb;
          else if (k == 1) x = vb;
          else {
            x = params[b * N * KP + d * KP + (k - 2)];
            if (c.rel_goal && k == kw - 1) x = __fadd_rn(x, yb);
          }
          if constexpr (KW > 0) w[d][k] = x; else if (lane == 0) wg[d * kw + k] = x;
        }
      }
    }
    if constexpr (KW == 0) __syncwarp();

    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      if (t < T) {
        float p0[N], p1[N], vv[N];
        const float* r0 = tabA + t * kw;
        if constexpr (MP == FG_MP_PROMP) {
          const int tn = (t < T - 1) ? t + 1 : t;          // last row: velocity copied from T-2 below
          const int tb = (t < T - 1) ? t : T - 2;
          const float* ra = tabA + tb * kw;
          const float* rb = tabA + (tb + 1) * kw;
          const float dtt = tabB[tb];
          (void)tn;
#pragma unroll
          for (int d = 0; d < N; ++d) {
            float a0 = 0.f, aa = 0.f, ab = 0.f;
#pragma unroll
            for (int k = 0; k < kw; ++k) {
              const float wk = (KW > 0) ? w[d][(KW > 0) ? k : 0] : wg[d * kw + k];
              a0 = fmaf(r0[k], wk, a0);
              aa = fmaf(ra[k], wk, aa);
              ab = fmaf(rb[k], wk, ab);
            }
            p0[d] = a0;
            p1[d] = ab;
            vv[d] = __fdiv_rn(__fsub_rn(ab, aa), dtt);
          }
          (void)p1;
        } else {
          const float* rv = tabB + t * kw;
#pragma unroll
          for (int d = 0; d < N; ++d) {
            float ap = 0.f, av = 0.f;
#pragma unroll
            for (int k = 0; k < kw; ++k) {
              const float wk = (KW > 0) ? w[d][(KW > 0) ? k : 0] : wg[d * kw + k];
              ap = fmaf(r0[k], wk, ap);
              av = fmaf(rv[k], wk, av);
            }
            p0[d] = ap;
            vv[d] = __fdiv_rn(av, c.tau);
          }
        }
#pragma unroll
        for (int d = 0; d < N; ++d) {
          tp[lane * N + d] = p0[d];
          tv[lane * N + d] = vv[d];
        }
      }
      __syncwarp();
      const int rows = min(32, T - t0);
      const long long base = (b * T + t0) * N;
      for (int i = lane; i < rows * N; i += 32) {
        pos_out[base + i] = tp[i];
        vel_out[base + i] = tv[i];
      }
      __syncwarp();
    }
  }
}

// DMP: serial semi-implicit Euler per (env, dof); one thread per (env, dof) pair.
__global__ void __launch_bounds__(kTrajThreads)
k_trajgen_dmp(const __grid_constant__ DevCfg c, const float* __restrict__ params, const float* __restrict__ bc_pos,
              const float* __restrict__ bc_vel, float* __restrict__ pos_out, float* __restrict__ vel_out,
              const long long B) {
  extern __shared__ float smem[];
  const int T = c.T, K = c.K, N = c.n_dof;
  float* tabA = smem;            // [T, K]
  float* tabB = tabA + T * K;    // [T-1]
  for (int i = threadIdx.x; i < T * K; i += kTrajThreads) tabA[i] = c.tab_a[i];
  for (int i = threadIdx.x; i < T - 1; i += kTrajThreads) tabB[i] = c.tab_b[i];
  __syncthreads();
  const long long e = (long long)blockIdx.x * kTrajThreads + threadIdx.x;   // (b, d) flat
  if (e >= B * N) return;
  const long long b = e / N;
  const int d = (int)(e % N);
  const float* pr = params + b * N * (K + 1) + d * (K + 1);
  float w[16];
  const int Kc = min(K, 16);
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = (k < Kc) ? __fmul_rn(pr[k], c.wscale) : 0.f;
  const float g = __fmul_rn(pr[K], c.gscale);
  float y = bc_pos[e], yd = __fmul_rn(bc_vel[e], c.tau);
  for (int t = 0; t < T; ++t) {
    pos_out[(b * T + t) * N + d] = y;
    vel_out[(b * T + t) * N + d] = __fdiv_rn(yd, c.tau);
    if (t < T - 1) {
      const float* row = tabA + t * K;
      float f = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < Kc) f = fmaf(row[k], w[k], f);
      for (int k = Kc; k < K; ++k) f = fmaf(row[k], __fmul_rn(pr[k], c.wscale), f);
      float a = __fmul_rn(c.beta, __fsub_rn(g, y));
      a = __fmul_rn(c.alpha, __fsub_rn(a, yd));
      a = __fadd_rn(a, f);
      const float h = tabB[t];
      yd = __fadd_rn(yd, __fmul_rn(h, a));
      y = __fadd_rn(y, __fmul_rn(h, yd));
    }
  }
}

}  // namespace fg
