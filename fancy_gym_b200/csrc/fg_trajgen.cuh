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

// closed-form MPs: ProMP (KW = K weighted columns) and ProDMP (KW = K+3: [y_b, tau*dy_b, w.., g]).
// KW > 0: compile-time column count, weights in registers; KW == 0: run-time fallback with the weights in shared memory.
//
// One warp owns one env; a lane owns quads of 4 consecutive time points, i.e. 4*N consecutive output floats, which it
// writes to the warp's staging buffer as N 16-byte vectors per array.  The env's whole pos / vel block ([T, N] floats
// each) then leaves the SM as two TMA bulk stores (cp.async.bulk shared->global) issued by one lane: full lines, no
// per-element global stores.  The next env's weights are prefetched before the current env is evaluated, and its
// evaluation overlaps the drain of the bulk stores.
#ifndef FG_TRAJ_MINB
#define FG_TRAJ_MINB 2
#endif
template <int MP, int N, int KW>
__global__ void __launch_bounds__(kTrajThreads, FG_TRAJ_MINB)
k_trajgen_closed(const __grid_constant__ DevCfg c, const float* __restrict__ params, const float* __restrict__ bc_pos,
                 const float* __restrict__ bc_vel, float* __restrict__ pos_out, float* __restrict__ vel_out,
                 const long long B, const int envs_per_block) {
  extern __shared__ __align__(128) float smem[];
  const int T = c.T;
  const int kw = (KW > 0) ? KW : c.cols_a;
  const int R4 = traj_r4(kw), REC4 = traj_rec4(MP, kw);
  const int nq = (T + 3) >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* stage = smem;                                                      // [warps][2][kTrajStageFloats]
  float4* rec = reinterpret_cast<float4*>(stage + kTrajWarps * 2 * kTrajStageFloats);   // [nq][REC4]
  float* wgen = reinterpret_cast<float*>(rec + nq * REC4);                  // KW == 0 only: [warps][N * kw]

  // ---- stage the quad records (packed once in fg_create; L2 resident) ----
  for (int i = tid; i < nq * REC4; i += kTrajThreads) rec[i] = c.quad_rec[i];
  __syncthreads();

  float* tp = stage + warp * 2 * kTrajStageFloats;
  float* tv = tp + kTrajStageFloats;
  float* wg = wgen + warp * N * kw;
  // params per dof: compile-time when the column count is (immediate load offsets)
  const int KP = (KW > 0) ? ((MP == FG_MP_PROMP) ? KW : KW - 2) : ((MP == FG_MP_PROMP) ? c.K : c.K + 1);
  const float r_tau = __frcp_rn(c.tau);
  constexpr int KWC = (KW > 0) ? KW : 1;
  constexpr int R4C = (KW > 0) ? traj_r4(KW) : 1;
  const int TN = T * N;
  const bool bulk = (TN % 4 == 0);                    // TMA needs 16-byte sizes / addresses
  const int chunk_rows = min(nq * 4, (kTrajStageFloats / (4 * N)) * 4);   // rows staged at a time (a multiple of 4)
  // Work distribution: a block owns `envs_per_block` consecutive envs (its warps take them round-robin) and the grid
  // is NOT persistent: the hardware block scheduler hands out blocks as SMs free up, which balances the two dies'
  // different distance to the L2 slices (measured: statically partitioned persistent grids top out ~15 % lower in
  // write bandwidth, profiles/README.md "store patterns").
  const long long b_end = min(B, ((long long)blockIdx.x + 1) * envs_per_block);
  constexpr long long warps_total = kTrajWarps;
  bool pending = false;                               // lane 0: a bulk store of this warp's stage is in flight

  // per-env weight vector, identical in every lane (broadcast loads)
  auto load_weights = [&](long long b, float (&w)[N][KWC]) {
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
  };

  long long b = (long long)blockIdx.x * envs_per_block + warp;
#ifdef FG_TRAJ_NOPREFETCH
  float w[N][KWC];
  for (; b < b_end; b += warps_total) {
    if constexpr (KW > 0) {
      load_weights(b, w);
    } else {
#else
  float w[N][KWC], wn[N][KWC];
  if constexpr (KW > 0) {
    if (b < b_end) load_weights(b, wn);
  }
  for (; b < b_end; b += warps_total) {
    if constexpr (KW > 0) {
#pragma unroll
      for (int d = 0; d < N; ++d)
#pragma unroll
        for (int k = 0; k < KWC; ++k) w[d][k] = wn[d][k];
      if (b + warps_total < b_end) load_weights(b + warps_total, wn);     // prefetch: in flight while this env is evaluated
    } else {
#endif
      __syncwarp();
      load_weights(b, w);
      __syncwarp();
    }

    // dot(table row, weights of dof d): FMA chain in index order, accumulator starts at 0 (oracle 'mirror' mode)
    auto dot_row = [&](const float4* row, int d) -> float {
      float acc = 0.f;
      if constexpr (KW > 0) {
        float r[R4C * 4];
#pragma unroll
        for (int j = 0; j < R4C; ++j) {
          const float4 x = row[j];
          r[4 * j] = x.x; r[4 * j + 1] = x.y; r[4 * j + 2] = x.z; r[4 * j + 3] = x.w;
        }
#pragma unroll
        for (int k = 0; k < KWC; ++k) acc = fmaf(r[k], w[d][k], acc);
      } else {
        const float* rw = reinterpret_cast<const float*>(row);
        for (int k = 0; k < kw; ++k) acc = fmaf(rw[k], wg[d * kw + k], acc);
      }
      return acc;
    };

    float* gp = pos_out + b * TN;
    float* gv = vel_out + b * TN;
    for (int c0 = 0; c0 < T; c0 += chunk_rows) {            // one chunk when the env fits the stage (T*N <= 1024)
      const int rows = min(chunk_rows, T - c0);
      if (pending) {                    // the previous bulk store must have read the stage before it is overwritten
        bulk_wait_read0();
        pending = false;
      }
      __syncwarp();
      for (int q = (c0 >> 2) + lane; 4 * q < c0 + rows; q += 32) {
        const float4* rq = rec + q * REC4;
        float pv[4 * N], vv[4 * N];     // this quad's 4*N consecutive output floats
        if constexpr (MP == FG_MP_PROMP) {
          const float4 dt4 = rq[5 * R4], rd4 = rq[5 * R4 + 1];
          const float dts[4] = {dt4.x, dt4.y, dt4.z, dt4.w}, rds[4] = {rd4.x, rd4.y, rd4.z, rd4.w};
          float prev[N];
#pragma unroll
          for (int d = 0; d < N; ++d) prev[d] = dot_row(rq, d);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int d = 0; d < N; ++d) {
              const float nxt = dot_row(rq + (r + 1) * R4, d);
              pv[r * N + d] = prev[d];
              vv[r * N + d] = div_by(__fsub_rn(nxt, prev[d]), dts[r], rds[r]);   // row T-1 is fixed up below
              prev[d] = nxt;
            }
          }
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int d = 0; d < N; ++d) {
              pv[r * N + d] = dot_row(rq + r * R4, d);
              vv[r * N + d] = div_by(dot_row(rq + (4 + r) * R4, d), c.tau, r_tau);
            }
          }
        }
        const int off = (4 * q - c0) * N;
        if (4 * q + 4 <= T) {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            reinterpret_cast<float4*>(tp + off)[j] = make_float4(pv[4 * j], pv[4 * j + 1], pv[4 * j + 2], pv[4 * j + 3]);
            reinterpret_cast<float4*>(tv + off)[j] = make_float4(vv[4 * j], vv[4 * j + 1], vv[4 * j + 2], vv[4 * j + 3]);
          }
        } else {                        // ragged tail (T not a multiple of 4)
#pragma unroll
          for (int j = 0; j < 4 * N; ++j)
            if (4 * q + j / N < T) {
              tp[off + j] = pv[j];
              tv[off + j] = vv[j];
            }
        }
      }
      __syncwarp();
      if (MP == FG_MP_PROMP && c0 + rows == T && lane < N) {          // vel[T-1] = vel[T-2]
        float v2 = 0.f;
        if (T - 2 >= c0) {
          v2 = tv[(T - 2 - c0) * N + lane];
        } else if (T >= 2) {            // T-2 sits in the previous chunk: re-evaluate it from the records
          const int q2 = (T - 2) >> 2, r2 = (T - 2) & 3;
          const float4* rq = rec + q2 * REC4;
          const float dtt = reinterpret_cast<const float*>(rq + 5 * R4)[r2], rdt = reinterpret_cast<const float*>(rq + 5 * R4 + 1)[r2];
          float a0 = 0.f, a1 = 0.f;
          const float* x0 = reinterpret_cast<const float*>(rq + r2 * R4);
          const float* x1 = reinterpret_cast<const float*>(rq + (r2 + 1) * R4);
          for (int k = 0; k < kw; ++k) {
            const float wk = params[b * N * KP + lane * KP + k];
            a0 = fmaf(x0[k], wk, a0);
            a1 = fmaf(x1[k], wk, a1);
          }
          v2 = div_by(__fsub_rn(a1, a0), dtt, rdt);
        }
        tv[(T - 1 - c0) * N + lane] = v2;
      }
      if (bulk) {
        fence_proxy_async_smem();       // generic-proxy smem writes -> visible to the async (TMA) proxy
        __syncwarp();
        if (lane == 0) {
          bulk_store_s2g(gp + c0 * N, tp, (unsigned)(rows * N * sizeof(float)));
          bulk_store_s2g(gv + c0 * N, tv, (unsigned)(rows * N * sizeof(float)));
          bulk_commit();
          pending = true;
        }
      } else {
        __syncwarp();
        for (int i = lane; i < rows * N; i += 32) {
          gp[c0 * N + i] = tp[i];
          gv[c0 * N + i] = tv[i];
        }
      }
    }
  }
  if (pending) bulk_wait_read0();       // the stage must stay valid until the TMA engine has read it
}

// DMP: the Euler recurrence is serial in time, so a lane owns one (env, dof) pair and a warp G = 32 / N envs.  The lanes
// write their position / scaled velocity into a per-warp staging buffer, CH time points at a time, and every env's
// [CH, N] blocks then leave the SM as TMA bulk stores (instead of 20-byte pieces scattered over G different lines per
// store instruction); the recurrence of the next chunk overlaps the drain.  Chunk size, measured at 262 144 envs x [200, 5]:
// 8 points 0.771 ms, 12: 0.690, 16: 0.687, 20: 0.623, 40: 0.797, 100: 1.38, 200: 2.44 — the recurrence is a latency chain, so
// resident warps (i.e. little staging memory per warp) matter more than large bulk copies.
#ifndef FG_DMP_THREADS
#define FG_DMP_THREADS 128
#endif
#ifndef FG_DMP_CHUNK
#define FG_DMP_CHUNK 20
#endif
constexpr int kDmpThreads = FG_DMP_THREADS;
constexpr int kDmpWarps = kDmpThreads / 32;
constexpr int kDmpChunk = FG_DMP_CHUNK;  // time points staged at a time (a multiple of 4: 16-byte sized bulk copies)

template <int N>
__global__ void __launch_bounds__(kDmpThreads)
k_trajgen_dmp(const __grid_constant__ DevCfg c, const float* __restrict__ params, const float* __restrict__ bc_pos,
              const float* __restrict__ bc_vel, float* __restrict__ pos_out, float* __restrict__ vel_out,
              const long long B, const int envs_per_block) {
  extern __shared__ __align__(128) float smem[];
  constexpr int G = 32 / N;                                  // envs per warp
  constexpr int SE = 2 * kDmpChunk * N;                      // staging floats per env (pos | vel)
  const int T = c.T, K = c.K;
  const int RA = (K + 3) & ~3;
  float* stage = smem;                                       // [warps][G][2][kDmpChunk * N]
  float* tabA = stage + kDmpWarps * G * SE;                  // [T, RA]
  float* tabB = tabA + T * RA;                               // [T]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < T * RA; i += kDmpThreads) {
    const int r = i / RA, col = i - r * RA;
    tabA[i] = (col < K) ? c.tab_a[r * K + col] : 0.f;
  }
  for (int i = tid; i < T - 1; i += kDmpThreads) tabB[i] = c.tab_b[i];
  __syncthreads();
  const int e = lane / N, d = lane - e * N;
  const bool lane_used = lane < G * N;
  const float r_tau = __frcp_rn(c.tau);
  const bool bulk = ((T * N) % 4 == 0);
  const long long b_end = min(B, ((long long)blockIdx.x + 1) * envs_per_block);
  float* st = stage + (warp * G + e) * SE;
  bool pending = false;
  for (long long b0 = (long long)blockIdx.x * envs_per_block + warp * G; b0 < b_end; b0 += kDmpWarps * G) {
    const long long b = b0 + e;
    const bool valid = lane_used && b < b_end;
    float w[16];
    float g = 0.f, y = 0.f, yd = 0.f;
    if (valid) {
      const float* pr = params + (b * N + d) * (K + 1);
#pragma unroll
      for (int k = 0; k < 16; ++k) w[k] = (k < K) ? __fmul_rn(pr[k], c.wscale) : 0.f;
      g = __fmul_rn(pr[K], c.gscale);
      y = bc_pos[b * N + d];
      yd = __fmul_rn(bc_vel[b * N + d], c.tau);
    }
    for (int c0 = 0; c0 < T; c0 += kDmpChunk) {
      const int rows = min(kDmpChunk, T - c0);
      if (pending) {
        bulk_wait_read0();
        pending = false;
      }
      __syncwarp();
      if (valid) {
        for (int r = 0; r < rows; ++r) {
          const int t = c0 + r;
          st[r * N + d] = y;
          st[kDmpChunk * N + r * N + d] = div_by(yd, c.tau, r_tau);
          if (t < T - 1) {
            const float* row = tabA + t * RA;
            float f = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              if (4 * k4 < K) {
                const float4 x = reinterpret_cast<const float4*>(row)[k4];
                f = fmaf(x.x, w[4 * k4], f);
                if (4 * k4 + 1 < K) f = fmaf(x.y, w[4 * k4 + 1], f);
                if (4 * k4 + 2 < K) f = fmaf(x.z, w[4 * k4 + 2], f);
                if (4 * k4 + 3 < K) f = fmaf(x.w, w[4 * k4 + 3], f);
              }
            float a = __fmul_rn(c.beta, __fsub_rn(g, y));
            a = __fmul_rn(c.alpha, __fsub_rn(a, yd));
            a = __fadd_rn(a, f);
            const float h = tabB[t];
            yd = __fadd_rn(yd, __fmul_rn(h, a));
            y = __fadd_rn(y, __fmul_rn(h, yd));
          }
        }
      }
      float* gp = pos_out + (b * T + c0) * N;
      float* gv = vel_out + (b * T + c0) * N;
      if (bulk) {
        fence_proxy_async_smem();
        __syncwarp();
        if (valid && d == 0) {            // one lane per env issues that env's two bulk stores
          bulk_store_s2g(gp, st, (unsigned)(rows * N * sizeof(float)));
          bulk_store_s2g(gv, st + kDmpChunk * N, (unsigned)(rows * N * sizeof(float)));
          bulk_commit();
          pending = true;
        }
      } else {
        __syncwarp();
        for (int ee = 0; ee < G; ++ee) {
          const long long be = b0 + ee;
          if (be >= b_end) break;
          const float* se = stage + (warp * G + ee) * SE;
          for (int i = lane; i < rows * N; i += 32) {
            pos_out[(be * T + c0) * N + i] = se[i];
            vel_out[(be * T + c0) * N + i] = se[kDmpChunk * N + i];
          }
        }
        __syncwarp();
      }
    }
  }
  if (pending) bulk_wait_read0();
}

}  // namespace fg
