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

// ---- TMA bulk store (shared -> global), SASS: UBLKCP --------------------------------------------
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int kTrajStageFloats = 1024;   // per warp and per output array: a whole [T, dof] trajectory when T*dof <= 1024

__host__ __device__ constexpr int traj_row_stride(int kw) {   // float4-padded, and an ODD number of float4 per row so
  int r = (kw + 3) & ~3;                                      // that 8 lanes reading 8 different rows hit 8 bank groups
  return ((r / 4) & 1) ? r : r + 4;
}

// closed-form MPs: ProMP (KW = K weighted columns) and ProDMP (KW = K+3: [y_b, tau*dy_b, w.., g]).
// KW > 0: compile-time column count, weights in registers, float4 row loads; KW == 0: run-time fallback.
// One warp owns one env: lanes own consecutive time points (31 per pass for ProMP, whose lane 31 only supplies
// pos[t+1] for lane 30's finite difference through a warp shuffle; 32 for ProDMP).  The env's whole pos / vel
// block is staged in shared memory and leaves the SM as two TMA bulk stores (cp.async.bulk shared->global) issued
// by one lane, i.e. full 128-byte lines and no per-element store instructions.
template <int MP, int N, int KW>
__global__ void __launch_bounds__(kTrajThreads)
k_trajgen_closed(const __grid_constant__ DevCfg c, const float* __restrict__ params, const float* __restrict__ bc_pos,
                 const float* __restrict__ bc_vel, float* __restrict__ pos_out, float* __restrict__ vel_out,
                 const long long B) {
  extern __shared__ __align__(128) float smem[];
  const int T = c.T, K = c.K;
  const int kw = (KW > 0) ? KW : c.cols_a;
  const int RA = traj_row_stride(kw);
  const int RB = (MP == FG_MP_PROMP) ? 1 : RA;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* stage = smem;                                          // [warps][2][kTrajStageFloats]  (128-byte aligned)
  float* tabA = stage + kTrajWarps * 2 * kTrajStageFloats;      // [T, RA]
  float* tabB = tabA + T * RA;                                  // ProMP: [T-1] time increments; ProDMP: [T, RB]
  float* tabR = tabB + ((c.rows_b * RB + 3) & ~3);              // ProMP: reciprocals of the increments
  float* wgen = tabR + ((c.rows_b + 3) & ~3);                   // fallback only: [warps][N*kw]
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

  float* tp = stage + warp * 2 * kTrajStageFloats;
  float* tv = tp + kTrajStageFloats;
  const int KP = (MP == FG_MP_PROMP) ? K : K + 1;
  const float r_tau = __frcp_rn(c.tau);
  constexpr int STEP = (MP == FG_MP_PROMP) ? 31 : 32;
  constexpr int KWC = (KW > 0) ? KW : 1;
  constexpr int RAC = (KW > 0) ? ((KW + 3) & ~3) : 4;
  const int TN = T * N;
  const bool whole = TN <= kTrajStageFloats;          // stage the whole env, else one pass of rows at a time
  const bool bulk = whole && (TN % 4 == 0);           // TMA needs 16-byte sizes / addresses
  const long long warps_total = (long long)gridDim.x * kTrajWarps;
  bool pending = false;                               // lane 0: a bulk store of this warp's stage is in flight
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
    if (pending) {                      // the previous env's stage must have been read out before it is overwritten
      bulk_wait_read0();
      pending = false;
    }
    __syncwarp();

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

    float* gp = pos_out + b * TN;
    float* gv = vel_out + b * TN;
    for (int t0 = 0; t0 < T; t0 += STEP) {
      const int t = t0 + lane;
      float p0[N], vv[N];
      const int tc = (t < T) ? t : T - 1;
      const float* r0 = tabA + tc * RA;
      if constexpr (MP == FG_MP_PROMP) {
        const int tb = (t < T - 1) ? t : T - 2;
        const float dtt = tabB[tb], rdt = tabR[tb];
#pragma unroll
        for (int d = 0; d < N; ++d) {
          p0[d] = dot_row(r0, d);
          const float pn = __shfl_down_sync(0xffffffffu, p0[d], 1);
          vv[d] = div_by(__fsub_rn(pn, p0[d]), dtt, rdt);     // garbage on row T-1 (and lane 31): fixed up below
        }
      } else {
        const float* rv = tabB + tc * RB;
#pragma unroll
        for (int d = 0; d < N; ++d) {
          p0[d] = dot_row(r0, d);
          vv[d] = div_by(dot_row(rv, d), c.tau, r_tau);
        }
      }
      const int rows = min(STEP, T - t0);
      const int off = whole ? t0 * N : 0;
      if (lane < rows) {
#pragma unroll
        for (int d = 0; d < N; ++d) {
          tp[off + lane * N + d] = p0[d];
          tv[off + lane * N + d] = vv[d];
        }
      }
      if (!whole) {       // long trajectories: flush this pass with plain coalesced stores
        __syncwarp();
        if (MP == FG_MP_PROMP && t0 + rows == T && rows >= 2 && lane < N)
          tv[(rows - 1) * N + lane] = tv[(rows - 2) * N + lane];
        __syncwarp();
        for (int i = lane; i < rows * N; i += 32) {
          gp[t0 * N + i] = tp[i];
          gv[t0 * N + i] = tv[i];
        }
        __syncwarp();
      }
    }
    if (whole) {
      __syncwarp();
      if (MP == FG_MP_PROMP && lane < N) tv[(T - 1) * N + lane] = tv[(T - 2) * N + lane];   // vel[T-1] = vel[T-2]
      if (bulk) {
        fence_proxy_async_smem();       // generic-proxy smem writes -> visible to the async (TMA) proxy
        __syncwarp();
        if (lane == 0) {
          bulk_store_s2g(gp, tp, (unsigned)(TN * sizeof(float)));
          bulk_store_s2g(gv, tv, (unsigned)(TN * sizeof(float)));
          bulk_commit();
          pending = true;
        }
      } else {
        __syncwarp();
        for (int i = lane; i < TN; i += 32) {
          gp[i] = tp[i];
          gv[i] = tv[i];
        }
        __syncwarp();
      }
    }
  }
  if (pending) bulk_wait_read0();       // the stage must stay valid until the TMA engine has read it
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
