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

// closed-form MPs: ProMP (KW = K weighted columns) and ProDMP (KW = K+3: [y_b, tau*dy_b, w.., g]).
// KW > 0: compile-time column count, weights in registers, float4 row loads; KW == 0: run-time fallback.
// One warp owns one env and walks its T rows 32 (ProDMP) or 31 (ProMP) at a time: for ProMP lane 31 only
// contributes pos[t+1] to lane 30's finite difference (warp shuffle), so no row is evaluated twice.
template <int MP, int N, int KW>
__global__ void __launch_bounds__(kTrajThreads)
k_trajgen_closed(const __grid_constant__ DevCfg c, const float* __restrict__ params, const float* __restrict__ bc_pos,
                 const float* __restrict__ bc_vel, float* __restrict__ pos_out, float* __restrict__ vel_out,
                 const long long B) {
  extern __shared__ __align__(16) float smem[];
  const int T = c.T, K = c.K;
  const int kw = (KW > 0) ? KW : c.cols_a;
  const int RA = (kw + 3) & ~3, RB = (c.cols_b + 3) & ~3;
  float* tabA = smem;                          // [T, RA]
  float* tabB = tabA + T * RA;                 // ProMP: [T-1, 4] time increments; ProDMP: [T, RB]
  float* tabR = tabB + c.rows_b * RB;          // ProMP: reciprocals of the increments
  float* tile = tabR + ((c.rows_b + 3) & ~3);  // [warps][2][32*N]
  float* wgen = tile + kTrajWarps * 2 * 32 * N;   // fallback only: [warps][N*kw]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < T * RA; i += kTrajThreads) {
    const int r = i / RA, col = i - r * RA;
    tabA[i] = (col < kw) ? c.tab_a[r * kw + col] : 0.f;
  }
  for (int i = tid; i < c.rows_b * RB; i += kTrajThreads) {
    const int r = i / RB, col = i - r * RB;
    tabB[i] = (col < c.cols_b) ? c.tab_b[r * c.cols_b + col] : 0.f;
  }
  if constexpr (MP == FG_MP_PROMP)
    for (int i = tid; i < c.rows_b; i += kTrajThreads) tabR[i] = __frcp_rn(c.tab_b[i]);
  __syncthreads();

  float* tp = tile + warp * 2 * 32 * N;
  float* tv = tp + 32 * N;
  const int KP = (MP == FG_MP_PROMP) ? K : K + 1;
  const float r_tau = __frcp_rn(c.tau);
  constexpr int STEP = (MP == FG_MP_PROMP) ? 31 : 32;
  constexpr int KWC = (KW > 0) ? KW : 1;
  constexpr int RAC = (KW > 0) ? ((KW + 3) & ~3) : 4;
  const long long warps_total = (long long)gridDim.x * kTrajWarps;
  for (long long b = (long long)blockIdx.x * kTrajWarps + warp; b < B; b += warps_total) {
    // per-env weight vector, identical in every lane (broadcast loads)
    float w[N][KWC];
    float* wg = wgen + warp * N * kw;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      float yb = 0.f, vb = 0.f;
      if constexpr (MP == FG_MP_PRODMP) {
        yb = bc_pos[b * N + d];
        vb = __fmul_rn(bc_vel[b * N + d], c.tau);
      }
#pragma unroll
      for (int k = 0; k < ((KW > 0) ? KW : 64); ++k) {
        if (k >= kw) break;
        float x;
        if constexpr (MP == FG_MP_PROMP) {
          x = params[b * N * KP + d * KP + k];
        } else {
          if (k == 0) x = yb;
          else if (k == 1) x = vb;
          else {
            x = params[b * N * KP + d * KP + (k - 2)];
            if (c.rel_goal && k == kw - 1) x = __fadd_rn(x, yb);
          }
        }
        if constexpr (KW > 0) w[d][k] = x; else if (lane == 0) wg[d * kw + k] = x;
      }
    }
    if constexpr (KW == 0) __syncwarp();

    auto dot_row = [&](const float* row, int d) -> float {
      float acc = 0.f;
      if constexpr (KW > 0) {
        float r[RAC];
#pragma unroll
        for (int j = 0; j < RAC / 4; ++j) {
          const float4 x = reinterpret_cast<const float4*>(row)[j];
          r[4 * j] = x.x; r[4 * j + 1] = x.y; r[4 * j + 2] = x.z; r[4 * j + 3] = x.w;
        }
#pragma unroll
        for (int k = 0; k < KWC; ++k) acc = fmaf(r[k], w[d][k], acc);
      } else {
        for (int k = 0; k < kw; ++k) acc = fmaf(row[k], wg[d * kw + k], acc);
      }
      return acc;
    };

    float carry[N];     // ProMP: velocity of the last finished row (for vel[T-1] = vel[T-2])
#pragma unroll
    for (int d = 0; d < N; ++d) carry[d] = 0.f;
    for (int t0 = 0; t0 < T; t0 += STEP) {
      const int t = t0 + lane;
      float p0[N], vv[N];
      const bool has_row = t < T;
      const int tc = has_row ? t : T - 1;
      if constexpr (MP == FG_MP_PROMP) {
        const float* r0 = tabA + tc * RA;
        const bool has_next = t < T - 1;
        const float dtt = has_next ? tabB[tc * RB] : 1.f, rdt = has_next ? tabR[tc] : 1.f;
#pragma unroll
        for (int d = 0; d < N; ++d) {
          p0[d] = dot_row(r0, d);
          const float pn = __shfl_down_sync(0xffffffffu, p0[d], 1);
          float vel = div_by(__fsub_rn(pn, p0[d]), dtt, rdt);
          const float prev = __shfl_up_sync(0xffffffffu, vel, 1);
          if (t == T - 1) vel = (lane == 0) ? carry[d] : prev;      // last row duplicates vel[T-2]
          vv[d] = vel;
          carry[d] = __shfl_sync(0xffffffffu, vel, STEP - 1);       // row t0+30 is the predecessor of the next tile
        }
      } else {
        const float* r0 = tabA + tc * RA;
        const float* rv = tabB + tc * RB;
#pragma unroll
        for (int d = 0; d < N; ++d) {
          p0[d] = dot_row(r0, d);
          vv[d] = div_by(dot_row(rv, d), c.tau, r_tau);
        }
      }
      const int rows = min(STEP, T - t0);
      if (lane < rows) {
#pragma unroll
        for (int d = 0; d < N; ++d) {
          tp[lane * N + d] = p0[d];
          tv[lane * N + d] = vv[d];
        }
      }
      __syncwarp();
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
