// Trajectory generation with a PER-ENV phase (learned tau / delay): the basis can no longer come from
// shared tables, so it is evaluated in the kernel — float32 linear phase with the library's elementwise ops, float64
// transcendental part rounded once to float32, i.e. exactly what the host does when it builds the shared tables
// (fancy_gym_b200/mp/basis_gn.py) and what oracle/mp.py 'mirror' mode specifies.  One block owns one env, a thread owns a
// time point; positions are exchanged through shared memory for ProMP's finite-difference velocity.  DMP takes two
// kernels: the block evaluates the forcing term for all time points in parallel (written to the velocity buffer), then
// k_dmp_integrate_phase runs the serial Euler recurrence with one LANE per (env, dof) pair — the recurrence is a latency
// chain, so it wants as many independent chains in flight as the machine holds, not one thread per dof of one env.
// The fused rollout consumes the result through FG_MP_TRAJ (8 KB per env from HBM: far below its compute time).
#include <cuda_runtime.h>

#include <cstdlib>

#include "fg_device.cuh"
#include "fg_dispatch.h"
#include "fg_trajgen.cuh"

namespace fg {

constexpr int kMaxRbf = 16;
#ifndef FG_PHASE_MINB
#define FG_PHASE_MINB 4
#endif

// normalised RBFs at linear phase z (float32), evaluated like basis_gn.basis64: float64; returns the phase x as well.
// NT = number of RBFs incl. zero padding, a template parameter: phi stays in registers (a run-time trip count puts it on
// the local-memory stack) and no exp() is issued for functions that do not exist (a run-time guard inside an unrolled
// loop gets if-converted: all 16 exp() would issue).  phi / sum: one reciprocal + a Markstein step per function
// (q = phi r; q += fma(-q, sum, phi) r) — the correctly rounded float64 quotient in 3 instructions instead of ~20.
template <int NT>
__device__ __forceinline__ double eval_basis(const PhaseArgs& a, float z, double (&phi)[NT]) {
  const double ph = a.phase_kind ? exp(-a.alpha_phase * (double)z) : (double)z;
  if constexpr (NT >= 3) {       // ProMP, linear phase: the recurrence, unless a value is too close to a float32 rounding boundary
    if (a.rec.on && !rbf_recurrence_eval<NT>(a.rec, a.cen, a.bw, ph, a.first, phi)) return ph;
  }
  double sum = 0.0;
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    const double d = ph - a.cen[k];
    phi[k] = exp(-((d * d * a.bw[k]) / 2));
    sum += phi[k];
  }
  if (NT > 1) {
    const double r = 1.0 / sum;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const double q = phi[k] * r;
      phi[k] = fma(fma(-q, sum, phi[k]), r, q);
    }
  }
  return ph;
}

// The same in float32, op for op what the library's torch tensors compute (mp_pytorch NormalizedRBFBasisGenerator.basis:
// tmp = (phase - centers)^2 * bandwidth; exp(-tmp / 2); / sum; ExpDecayPhaseGenerator.phase: exp(-alpha * z)): ~9x fewer
// instructions than the float64 evaluation (6 expf instead of 6 double-precision exp per time point), which turns the
// kernel from instruction bound (28 % fp64 pipe) into HBM bound.  coef[] = the float32 table entry of basis function kk.
template <int MPK, int NT>
__device__ __forceinline__ void eval_coef32(const PhaseArgs& a, float z, float (&coef)[NT]) {
  const float ph = a.phase_kind ? expf(__fmul_rn(-(float)a.alpha_phase, z)) : z;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    const float d = __fsub_rn(ph, a.cen32[k]);
    coef[k] = expf(__fmul_rn(-__fmul_rn(__fmul_rn(d, d), a.bw32[k]), 0.5f));       // (/ 2 == * 0.5 exactly)
    sum = __fadd_rn(sum, coef[k]);
  }
  const float r_sum = __frcp_rn(sum);
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    if (NT > 1) coef[k] = div_by(coef[k], sum, r_sum);       // the IEEE quotient (fg_device.cuh)
    // ProMP: basis * weights_scale; DMP: canonical x * basis (* the scale when it sits on the basis)
    coef[k] = (MPK == FG_MP_PROMP) ? __fmul_rn(coef[k], a.wscale) : __fmul_rn(__fmul_rn(ph, coef[k]), (float)a.basis_scale);
  }
}

// MPK: the MP type (one instantiation per type keeps the register footprint of each at its own needs);
// NT: ProMP / DMP number of RBFs incl. zero padding; ProDMP number of weighted columns K + 1 (weights + goal)
template <int MPK, int NT>
__global__ void __launch_bounds__(256, FG_PHASE_MINB) k_trajgen_phase(const __grid_constant__ PhaseArgs a) {
  extern __shared__ float sm[];
  const int N = a.N, TM = a.T, K = a.K;      // TM: rows of the output buffers (the longest plan)
  const int KP = (MPK == FG_MP_PROMP) ? K : K + 1;
  float* s_pos = sm;                 // [TM, N]  (DMP: the forcing term)
  float* s_vel = s_pos + TM * N;     // [TM, N]  (DMP: unused)
  float* s_w = s_vel + TM * N;       // [N, KP] this env's parameters
  const long long b = blockIdx.x;
  // ragged plans: this env's own number of points and its own time grid (rows beyond it are never read by the rollout)
  const int T = a.n_steps_env ? min(max(a.n_steps_env[b], 2), TM) : TM;
  const float* times = a.n_steps_env ? a.times_table + (long long)T * a.times_stride : a.times;
  const float tau = a.tau[b], delay = a.delay[b];
  for (int i = threadIdx.x; i < N * KP; i += blockDim.x) {
    float p = a.params[b * N * KP + i];
    if (MPK == FG_MP_DMP) p = __fmul_rn(p, (i % KP < K) ? a.wscale : a.gscale);
    if (MPK == FG_MP_PRODMP && a.rel_goal && (i % KP) == K) p = __fadd_rn(p, a.bc_pos[b * N + i / KP]);
    s_w[i] = p;
  }
  __syncthreads();
  if constexpr (MPK == FG_MP_PRODMP) {
    // ProDMP (App. B.7): per env only the LOOKUP into the pre-integrated bases changes with tau / delay.  Indices are
    // rounded in float32 like the library, the blend with the boundary condition is float64 rounded once to float32 —
    // the arithmetic of the host-built shared tables (fancy_gym_b200/mp/mp.py ProDMP.tables).
    auto index_of = [&](float t32) {
      const float z = fmaxf(__fdiv_rn(__fsub_rn(t32, delay), tau), 0.f);
      const int i = (int)rintf(__fdiv_rn(z, a.scaled_dt));
      return min(max(i, 0), a.n_pc - 1);
    };
    constexpr int KG = NT;        // == K + 1 (checked by the launcher)
    const int ib = index_of(a.init_time);
    const double y1b = a.pc_y[ib * 4], y2b = a.pc_y[ib * 4 + 1], dy1b = a.pc_y[ib * 4 + 2], dy2b = a.pc_y[ib * 4 + 3];
    const double det = y1b * dy2b - y2b * dy1b;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const int ix = index_of(times[t]);
      const double y1 = a.pc_y[ix * 4], y2 = a.pc_y[ix * 4 + 1], dy1 = a.pc_y[ix * 4 + 2], dy2 = a.pc_y[ix * 4 + 3];
      const double xi1 = dy2b / det * y1 - dy1b / det * y2, xi2 = y1b / det * y2 - y2b / det * y1;
      const double xi3 = dy2b / det * dy1 - dy1b / det * dy2, xi4 = y1b / det * dy2 - y2b / det * dy1;
      float hp[NT], hv[NT];         // KG == NT, a compile-time trip count: registers, not the local-memory stack
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        hp[k] = (float)((a.pc_pos[ix * KG + k] - xi1 * a.pc_pos[ib * KG + k] - xi2 * a.pc_vel[ib * KG + k]) * a.scale[k]);
        hv[k] = (float)((a.pc_vel[ix * KG + k] - xi3 * a.pc_pos[ib * KG + k] - xi4 * a.pc_vel[ib * KG + k]) * a.scale[k]);
      }
      const float x1 = (float)xi1, x2 = (float)xi2, x3 = (float)xi3, x4 = (float)xi4;
      for (int d = 0; d < N; ++d) {
        const float yb = a.bc_pos[b * N + d], vb = __fmul_rn(a.bc_vel[b * N + d], tau);
        float ap = fmaf(x2, vb, fmaf(x1, yb, 0.f)), av = fmaf(x4, vb, fmaf(x3, yb, 0.f));     // table columns 0, 1
#pragma unroll
        for (int k = 0; k < NT; ++k) {
          ap = fmaf(hp[k], s_w[d * KP + k], ap);
          av = fmaf(hv[k], s_w[d * KP + k], av);
        }
        s_pos[t * N + d] = ap;
        s_vel[t * N + d] = __fdiv_rn(av, tau);
      }
    }
  } else if constexpr (NT <= kMaxRbf) {
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const float un = __fdiv_rn(__fsub_rn(times[t], delay), tau);       // float32 elementwise ops of the library
      const float z = fminf(fmaxf(un, 0.f), (a.phase_kind && !a.exp_right_clip) ? INFINITY : 1.f);
      // coefficient of weighted basis function k = kk - first: float32 like the library, or float64 rounded once like the
      // host-built tables (eval_f64)
      float coef[NT];
      if (a.eval_f64) {
        double phi[NT];
        const double x = eval_basis<NT>(a, z, phi);
#pragma unroll
        for (int kk = 0; kk < NT; ++kk)
          coef[kk] = (MPK == FG_MP_PROMP) ? __fmul_rn((float)phi[kk], a.wscale) : (float)(x * phi[kk] * a.basis_scale);
      } else {
        eval_coef32<MPK, NT>(a, z, coef);
      }
      for (int d = 0; d < N; ++d) {
        float acc = 0.f;
#pragma unroll
        for (int kk = 0; kk < NT; ++kk)          // FMA chain in index order over the weighted functions
          if (kk >= a.first && kk < a.first + K) acc = fmaf(coef[kk], s_w[d * KP + kk - a.first], acc);
        s_pos[t * N + d] = acc;
      }
    }
  }
  __syncthreads();
  float* gp = a.pos + b * TM * N;
  float* gv = a.vel + b * TM * N;
  if constexpr (MPK == FG_MP_DMP) {
    // the forcing term goes to the velocity buffer; k_dmp_integrate_phase turns it into the trajectory in place
    for (int i = threadIdx.x; i < T * N; i += blockDim.x) gv[i] = s_pos[i];
  } else {
    if constexpr (MPK == FG_MP_PROMP) {
      for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int ts = (t < T - 1) ? t : T - 2;                 // vel[T-1] = vel[T-2]
        for (int d = 0; d < N; ++d)
          s_vel[t * N + d] = (T > 1) ? __fdiv_rn(__fsub_rn(s_pos[(ts + 1) * N + d], s_pos[ts * N + d]),
                                                 __fsub_rn(times[ts + 1], times[ts])) : 0.f;
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < T * N; i += blockDim.x) {
      gp[i] = s_pos[i];
      gv[i] = s_vel[i];
    }
  }
  for (int i = T * N + threadIdx.x; i < TM * N; i += blockDim.x) gp[i] = gv[i] = 0.f;        // ragged: rows past this env's plan
}

// The same evaluation with ONE WARP per env (the default; the block-per-env kernel above stays as the fallback behind
// FG_PHASE_BLOCK=1 and as the reference the two are tested against each other with).  A block per env spends most of its
// time in its own latency chain — parameters -> barrier -> float64 basis -> barrier -> difference -> barrier -> store — with
// 4 blocks per SM to overlap; a warp per env needs no block barrier, 32 time points per pass go through a 1.3 KB per-warp
// stage, and 24-32 independent envs are in flight per SM.  ProMP's forward difference needs the NEXT point: a pass advances
// by 31 points and lane 31 only supplies that neighbour (it is re-evaluated as lane 0 of the next pass); the last point
// copies the velocity of the one before it, across a pass boundary through a carried row.
constexpr int kPhaseWarps = 4;
constexpr int kPhaseRows = 32;
constexpr int kPhaseWarpFloats = FG_MAX_DOF * 17 + (2 * kPhaseRows + 1) * FG_MAX_DOF;

template <int MPK, int NT>
__global__ void __launch_bounds__(kPhaseWarps * 32, (MPK == FG_MP_PRODMP) ? 6 : 8)
k_trajgen_phase_warp(const __grid_constant__ PhaseArgs a, const long long B) {
  __shared__ float s_all[kPhaseWarps][kPhaseWarpFloats];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * kPhaseWarps + warp;
  if (b >= B) return;                          // whole warps leave together; the kernel has no block barrier
  const int N = a.N, TM = a.T, K = a.K;
  const int KP = (MPK == FG_MP_PROMP) ? K : K + 1;
  float* s_w = s_all[warp];                    // [N, KP] this env's parameters
  float* st_p = s_w + FG_MAX_DOF * 17;         // [32, N] positions of this pass (DMP: the forcing term)
  float* st_v = st_p + kPhaseRows * FG_MAX_DOF;   // [32, N] velocities of this pass
  float* st_c = st_v + kPhaseRows * FG_MAX_DOF;   // [N]     velocity row carried across a pass boundary (ProMP)
  const int T = a.n_steps_env ? min(max(a.n_steps_env[b], 2), TM) : TM;
  const float* times = a.n_steps_env ? a.times_table + (long long)T * a.times_stride : a.times;
  const float tau = a.tau[b], delay = a.delay[b];
  const float r_tau = __frcp_rn(tau);
  for (int i = lane; i < N * KP; i += 32) {
    float p = a.params[b * N * KP + i];
    if (MPK == FG_MP_DMP) p = __fmul_rn(p, (i % KP < K) ? a.wscale : a.gscale);
    if (MPK == FG_MP_PRODMP && a.rel_goal && (i % KP) == K) p = __fadd_rn(p, a.bc_pos[b * N + i / KP]);
    s_w[i] = p;
  }
  if (lane < N) st_c[lane] = 0.f;
  __syncwarp();

  // ProDMP: boundary-condition row of the pre-integrated bases (same for every time point of the env)
  int ib = 0;
  double y1b = 0, y2b = 0, dy1b = 0, dy2b = 0, det = 1;
  auto index_of = [&](float t32) {
    const float z = fmaxf(__fdiv_rn(__fsub_rn(t32, delay), tau), 0.f);
    const int i = (int)rintf(__fdiv_rn(z, a.scaled_dt));
    return min(max(i, 0), a.n_pc - 1);
  };
  if constexpr (MPK == FG_MP_PRODMP) {
    ib = index_of(a.init_time);
    y1b = a.pc_y[ib * 4]; y2b = a.pc_y[ib * 4 + 1]; dy1b = a.pc_y[ib * 4 + 2]; dy2b = a.pc_y[ib * 4 + 3];
    det = y1b * dy2b - y2b * dy1b;
  }

  constexpr int STEP = (MPK == FG_MP_PROMP) ? kPhaseRows - 1 : kPhaseRows;
  float* gp = a.pos + b * TM * N;
  float* gv = a.vel + b * TM * N;
  for (int t0 = 0; t0 < T; t0 += STEP) {
    const int t = t0 + lane;
    const bool valid = t < T;
    if (valid) {
      if constexpr (MPK == FG_MP_PRODMP) {
        constexpr int KG = NT;        // == K + 1 (checked by the launcher)
        const int ix = index_of(times[t]);
        const double y1 = a.pc_y[ix * 4], y2 = a.pc_y[ix * 4 + 1], dy1 = a.pc_y[ix * 4 + 2], dy2 = a.pc_y[ix * 4 + 3];
        const double xi1 = dy2b / det * y1 - dy1b / det * y2, xi2 = y1b / det * y2 - y2b / det * y1;
        const double xi3 = dy2b / det * dy1 - dy1b / det * dy2, xi4 = y1b / det * dy2 - y2b / det * dy1;
        float hp[NT], hv[NT];
#pragma unroll
        for (int k = 0; k < NT; ++k) {
          hp[k] = (float)((a.pc_pos[ix * KG + k] - xi1 * a.pc_pos[ib * KG + k] - xi2 * a.pc_vel[ib * KG + k]) * a.scale[k]);
          hv[k] = (float)((a.pc_vel[ix * KG + k] - xi3 * a.pc_pos[ib * KG + k] - xi4 * a.pc_vel[ib * KG + k]) * a.scale[k]);
        }
        const float x1 = (float)xi1, x2 = (float)xi2, x3 = (float)xi3, x4 = (float)xi4;
        for (int d = 0; d < N; ++d) {
          const float yb = a.bc_pos[b * N + d], vb = __fmul_rn(a.bc_vel[b * N + d], tau);
          float ap = fmaf(x2, vb, fmaf(x1, yb, 0.f)), av = fmaf(x4, vb, fmaf(x3, yb, 0.f));     // table columns 0, 1
#pragma unroll
          for (int k = 0; k < NT; ++k) {
            ap = fmaf(hp[k], s_w[d * KP + k], ap);
            av = fmaf(hv[k], s_w[d * KP + k], av);
          }
          st_p[lane * N + d] = ap;
          st_v[lane * N + d] = __fdiv_rn(av, tau);
        }
      } else if constexpr (NT <= kMaxRbf) {
        const float un = div_by(__fsub_rn(times[t], delay), tau, r_tau);   // float32 elementwise ops of the library
        const float z = fminf(fmaxf(un, 0.f), (a.phase_kind && !a.exp_right_clip) ? INFINITY : 1.f);
        float coef[NT];
        if (a.eval_f64) {
          double phi[NT];
          const double x = eval_basis<NT>(a, z, phi);
#pragma unroll
          for (int kk = 0; kk < NT; ++kk)
            coef[kk] = (MPK == FG_MP_PROMP) ? __fmul_rn((float)phi[kk], a.wscale) : (float)(x * phi[kk] * a.basis_scale);
        } else {
          eval_coef32<MPK, NT>(a, z, coef);
        }
        for (int d = 0; d < N; ++d) {
          float acc = 0.f;
#pragma unroll
          for (int kk = 0; kk < NT; ++kk)          // FMA chain in index order over the weighted functions
            if (kk >= a.first && kk < a.first + K) acc = fmaf(coef[kk], s_w[d * KP + kk - a.first], acc);
          st_p[lane * N + d] = acc;
        }
      }
    }
    __syncwarp();
    if constexpr (MPK == FG_MP_PROMP) {
      if (t < T - 1 && lane < STEP) {
        const float dtt = __fsub_rn(times[t + 1], times[t]), r_dtt = __frcp_rn(dtt);
        for (int d = 0; d < N; ++d)
          st_v[lane * N + d] = div_by(__fsub_rn(st_p[(lane + 1) * N + d], st_p[lane * N + d]), dtt, r_dtt);
      }
      __syncwarp();
      if (t == T - 1 && lane < STEP)                 // vel[T-1] = vel[T-2] (0 for a one-point plan: the carried row starts at 0)
        for (int d = 0; d < N; ++d) st_v[lane * N + d] = (lane > 0) ? st_v[(lane - 1) * N + d] : st_c[d];
      if (lane == STEP - 1 && t < T - 1)
        for (int d = 0; d < N; ++d) st_c[d] = st_v[lane * N + d];
      __syncwarp();
    }
    const int rows = min(STEP, T - t0);
    for (int i = lane; i < rows * N; i += 32) {
      if constexpr (MPK == FG_MP_DMP) {
        gv[t0 * N + i] = st_p[i];      // the forcing term goes to the velocity buffer (k_dmp_integrate_phase works in place)
      } else {
        gp[t0 * N + i] = st_p[i];
        gv[t0 * N + i] = st_v[i];
      }
    }
    __syncwarp();
  }
  for (int i = T * N + lane; i < TM * N; i += 32) gp[i] = gv[i] = 0.f;          // ragged: rows past this env's plan
}

// DMP with a per-env phase, second half: semi-implicit Euler in scaled time, every op rounded separately (the library's
// recurrence).  A lane owns one (env, dof) pair and a warp G = 32 / N envs, like k_trajgen_dmp; the forcing term comes from
// the velocity buffer (written by k_trajgen_phase) and is replaced in place, CH time points at a time through a per-warp
// staging buffer and per-env TMA bulk stores (20 points: small stages keep ~44 warps resident per SM, which this
// latency chain needs more than it needs large copies).  A chunk's forcing values are read before its stores are issued, the
// scaled-time increments h = max((t_{i+1} - delay) / tau, 0) - max((t_i - delay) / tau, 0) are float32 ops off the chain.
constexpr int kPhaseDmpChunk = 20;

template <int N>
__global__ void __launch_bounds__(kDmpThreads)
k_dmp_integrate_phase(const __grid_constant__ PhaseArgs a, const long long B) {
  extern __shared__ __align__(128) float smem[];
  constexpr int G = 32 / N;
  constexpr int CH = kPhaseDmpChunk;
  constexpr int SE = 2 * CH * N;                             // staging floats per env (pos | vel)
  const int TM = a.T, K = a.K, KP = K + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = lane / N, d = lane - e * N;
  const long long b = ((long long)blockIdx.x * kDmpWarps + warp) * G + e;
  const bool valid = lane < G * N && b < B;
  float* st = smem + (warp * G + e) * SE;
  int T = 0;
  const float* times = a.times;
  float tau = 1.f, delay = 0.f, g = 0.f, y = 0.f, yd = 0.f;
  if (valid) {
    T = a.n_steps_env ? min(max(a.n_steps_env[b], 2), TM) : TM;
    if (a.n_steps_env) times = a.times_table + (long long)T * a.times_stride;
    tau = a.tau[b];
    delay = a.delay[b];
    g = __fmul_rn(a.params[(b * N + d) * KP + K], a.gscale);
    y = a.bc_pos[b * N + d];
    yd = __fmul_rn(a.bc_vel[b * N + d], tau);
  }
  int t_warp = T;                                            // the longest plan of this warp's envs
#pragma unroll
  for (int o = 16; o; o >>= 1) t_warp = max(t_warp, __shfl_xor_sync(0xffffffffu, t_warp, o));
  const bool bulk = (TM * N) % 4 == 0;                       // every env's rows start 16-byte aligned
  float* gp = a.pos + b * TM * N;
  float* gv = a.vel + b * TM * N;
  float s_prev = valid ? fmaxf(__fdiv_rn(__fsub_rn(times[0], delay), tau), 0.f) : 0.f;
  bool pending = false;
  for (int c0 = 0; c0 < t_warp; c0 += CH) {
    if (pending) {
      bulk_wait_read0();
      pending = false;
    }
    __syncwarp();
    const int rows = max(0, min(CH, T - c0));
    for (int r0 = 0; r0 < rows; r0 += 8) {
      float ff[8], hh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {                            // loads and phase arithmetic first: nothing of it is on the chain
        const int t = min(c0 + r0 + j, T - 2);
        ff[j] = gv[t * N + d];
        const float s_next = fmaxf(__fdiv_rn(__fsub_rn(times[t + 1], delay), tau), 0.f);
        hh[j] = __fsub_rn(s_next, s_prev);
        if (r0 + j < rows && c0 + r0 + j < T - 1) s_prev = s_next;      // (rows past the chunk belong to the next pass)
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = r0 + j;
        if (r < rows) {
          st[r * N + d] = y;
          st[CH * N + r * N + d] = __fdiv_rn(yd, tau);
          if (c0 + r < T - 1) {
            float acc = __fmul_rn(a.beta, __fsub_rn(g, y));
            acc = __fmul_rn(a.alpha, __fsub_rn(acc, yd));
            acc = __fadd_rn(acc, ff[j]);
            yd = __fadd_rn(yd, __fmul_rn(hh[j], acc));
            y = __fadd_rn(y, __fmul_rn(hh[j], yd));
          }
        }
      }
    }
    fence_proxy_async_smem();                       // generic-proxy smem writes -> visible to the async (TMA) proxy
    __syncwarp();
    if (bulk && (rows * N) % 4 == 0) {
      if (valid && d == 0 && rows > 0) {            // one lane per env issues that env's two bulk stores
        bulk_store_s2g(gp + c0 * N, st, (unsigned)(rows * N * sizeof(float)));
        bulk_store_s2g(gv + c0 * N, st + CH * N, (unsigned)(rows * N * sizeof(float)));
        bulk_commit();
        pending = true;
      }
    } else {
      for (int r = 0; r < rows; ++r) {              // ragged tails / odd sizes: every lane stores what it staged itself
        gp[(c0 + r) * N + d] = st[r * N + d];
        gv[(c0 + r) * N + d] = st[CH * N + r * N + d];
      }
    }
  }
  if (pending) bulk_wait_read0();
}

template <int N>
static cudaError_t launch_dmp_integrate(const PhaseArgs& a, long long B, cudaStream_t stream) {
  constexpr int G = 32 / N;
  const size_t smem = sizeof(float) * (size_t)kDmpWarps * G * 2 * kPhaseDmpChunk * N;
  const long long per_block = (long long)kDmpWarps * G;
  k_dmp_integrate_phase<N><<<(unsigned)((B + per_block - 1) / per_block), kDmpThreads, smem, stream>>>(a, B);
  return cudaGetLastError();
}

template <int MPK, int NT>
static cudaError_t launch_phase_nt(const PhaseArgs& a, long long B, cudaStream_t stream, size_t smem, bool block_per_env) {
  if (!block_per_env) {
    k_trajgen_phase_warp<MPK, NT><<<(unsigned)((B + kPhaseWarps - 1) / kPhaseWarps), kPhaseWarps * 32, 0, stream>>>(a, B);
    return cudaGetLastError();
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_trajgen_phase<MPK, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int threads = ((a.T + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  k_trajgen_phase<MPK, NT><<<(unsigned)B, threads, smem, stream>>>(a);
  return cudaGetLastError();
}

template <int MPK>
static cudaError_t launch_phase_mp(const PhaseArgs& a, long long B, cudaStream_t stream, size_t smem, bool block_per_env,
                                   const char** why) {
  const int nt = (MPK == FG_MP_PRODMP) ? a.K + 1 : a.n_total;
#define FG_NT(n) case n: return launch_phase_nt<MPK, n>(a, B, stream, smem, block_per_env);
  switch (nt) {
    FG_NT(1) FG_NT(2) FG_NT(3) FG_NT(4) FG_NT(5) FG_NT(6) FG_NT(7) FG_NT(8) FG_NT(9) FG_NT(10) FG_NT(11) FG_NT(12)
    FG_NT(13) FG_NT(14) FG_NT(15) FG_NT(16)
    case 17: if constexpr (MPK == FG_MP_PRODMP) return launch_phase_nt<MPK, 17>(a, B, stream, smem, block_per_env);
  }
#undef FG_NT
  *why = "at most 16 basis functions";
  return cudaSuccess;
}

cudaError_t launch_trajgen_phase(const PhaseArgs& a, long long B, cudaStream_t stream, int max_smem_optin, const char** why) {
  const int KP = (a.mp_kind == FG_MP_PROMP) ? a.K : a.K + 1;
  const size_t smem = sizeof(float) * ((size_t)2 * a.T * a.N + (size_t)a.N * KP);
  // one warp per env by default; FG_PHASE_BLOCK=1 selects the block-per-env kernel (kept for A/B tests; it also needs the
  // whole trajectory in shared memory, which the warp kernel does not)
  static const bool block_per_env = [] { const char* e = getenv("FG_PHASE_BLOCK"); return e && atoi(e) > 0; }();
  if (block_per_env && smem > (size_t)max_smem_optin) {
    *why = "trajectory too long for the block-per-env phase kernel";
    return cudaSuccess;
  }
  cudaError_t e = cudaSuccess;
  switch (a.mp_kind) {
    case FG_MP_PROMP: e = launch_phase_mp<FG_MP_PROMP>(a, B, stream, smem, block_per_env, why); break;
    case FG_MP_DMP: e = launch_phase_mp<FG_MP_DMP>(a, B, stream, smem, block_per_env, why); break;
    case FG_MP_PRODMP: e = launch_phase_mp<FG_MP_PRODMP>(a, B, stream, smem, block_per_env, why); break;
    default: *why = "unknown mp_kind"; return cudaSuccess;
  }
  if (*why) return cudaSuccess;
  if (e != cudaSuccess || a.mp_kind != FG_MP_DMP) return e;
#define FG_N(n) case n: return launch_dmp_integrate<n>(a, B, stream);
  switch (a.N) {
    FG_N(1) FG_N(2) FG_N(3) FG_N(4) FG_N(5) FG_N(6) FG_N(7) FG_N(8)
  }
#undef FG_N
  *why = "n_dof out of range";
  return cudaSuccess;
}

}  // namespace fg
