// Trajectory generation with a PER-ENV phase (learned tau / delay): the basis can no longer come from
// shared tables, so it is evaluated in the kernel — float32 linear phase with the library's elementwise ops, float64
// transcendental part rounded once to float32, i.e. exactly what the host does when it builds the shared tables
// (fancy_gym_b200/mp/basis_gn.py) and what oracle/mp.py 'mirror' mode specifies.  One block owns one env, a thread owns a
// time point; positions are exchanged through shared memory for ProMP's finite-difference velocity, DMP runs its serial
// Euler recurrence on one thread per dof after the block has evaluated the forcing term for all time points in parallel.
// The fused rollout consumes the result through FG_MP_TRAJ (8 KB per env from HBM: far below its compute time).
#include <cuda_runtime.h>

#include "fg_device.cuh"
#include "fg_dispatch.h"

namespace fg {

// normalised RBFs at linear phase z (float32), evaluated like basis_gn.basis64: float64, returns the phase x as well
__device__ __forceinline__ double eval_basis(const PhaseArgs& a, float z, double (&phi)[16]) {
  const double ph = a.phase_kind ? exp(-a.alpha_phase * (double)z) : (double)z;
  double sum = 0.0;
  for (int k = 0; k < a.n_total; ++k) {
    const double d = ph - a.cen[k];
    phi[k] = exp(-((d * d * a.bw[k]) / 2));
    sum += phi[k];
  }
  if (a.n_total > 1)
    for (int k = 0; k < a.n_total; ++k) phi[k] = phi[k] / sum;
  return ph;
}

__global__ void k_trajgen_phase(const __grid_constant__ PhaseArgs a) {
  extern __shared__ float sm[];
  const int N = a.N, TM = a.T, K = a.K;      // TM: rows of the output buffers (the longest plan)
  const int KP = (a.mp_kind == FG_MP_PROMP) ? K : K + 1;
  float* s_pos = sm;                 // [TM, N]
  float* s_vel = s_pos + TM * N;     // [TM, N]
  float* s_f = s_vel + TM * N;       // DMP: forcing [TM, N]
  float* s_h = s_f + TM * N;         // DMP: scaled-time increments [TM]
  float* s_w = s_h + TM;             // [N, KP] this env's parameters
  const long long b = blockIdx.x;
  // ragged plans: this env's own number of points and its own time grid (rows beyond it are never read by the rollout)
  const int T = a.n_steps_env ? min(max(a.n_steps_env[b], 2), TM) : TM;
  const float* times = a.n_steps_env ? a.times_table + (long long)T * a.times_stride : a.times;
  const float tau = a.tau[b], delay = a.delay[b];
  for (int i = threadIdx.x; i < N * KP; i += blockDim.x) {
    float p = a.params[b * N * KP + i];
    if (a.mp_kind == FG_MP_DMP) p = __fmul_rn(p, (i % KP < K) ? a.wscale : a.gscale);
    if (a.mp_kind == FG_MP_PRODMP && a.rel_goal && (i % KP) == K) p = __fadd_rn(p, a.bc_pos[b * N + i / KP]);
    s_w[i] = p;
  }
  __syncthreads();
  if (a.mp_kind == FG_MP_PRODMP) {
    // ProDMP (App. B.7): per env only the LOOKUP into the pre-integrated bases changes with tau / delay.  Indices are
    // rounded in float32 like the library, the blend with the boundary condition is float64 rounded once to float32 —
    // the arithmetic of the host-built shared tables (fancy_gym_b200/mp/mp.py ProDMP.tables).
    auto index_of = [&](float t32) {
      const float z = fmaxf(__fdiv_rn(__fsub_rn(t32, delay), tau), 0.f);
      const int i = (int)rintf(__fdiv_rn(z, a.scaled_dt));
      return min(max(i, 0), a.n_pc - 1);
    };
    const int KG = K + 1;
    const int ib = index_of(a.init_time);
    const double y1b = a.pc_y[ib * 4], y2b = a.pc_y[ib * 4 + 1], dy1b = a.pc_y[ib * 4 + 2], dy2b = a.pc_y[ib * 4 + 3];
    const double det = y1b * dy2b - y2b * dy1b;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const int ix = index_of(times[t]);
      const double y1 = a.pc_y[ix * 4], y2 = a.pc_y[ix * 4 + 1], dy1 = a.pc_y[ix * 4 + 2], dy2 = a.pc_y[ix * 4 + 3];
      const double xi1 = dy2b / det * y1 - dy1b / det * y2, xi2 = y1b / det * y2 - y2b / det * y1;
      const double xi3 = dy2b / det * dy1 - dy1b / det * dy2, xi4 = y1b / det * dy2 - y2b / det * dy1;
      float hp[17], hv[17];
      for (int k = 0; k < KG; ++k) {
        hp[k] = (float)((a.pc_pos[ix * KG + k] - xi1 * a.pc_pos[ib * KG + k] - xi2 * a.pc_vel[ib * KG + k]) * a.scale[k]);
        hv[k] = (float)((a.pc_vel[ix * KG + k] - xi3 * a.pc_pos[ib * KG + k] - xi4 * a.pc_vel[ib * KG + k]) * a.scale[k]);
      }
      const float x1 = (float)xi1, x2 = (float)xi2, x3 = (float)xi3, x4 = (float)xi4;
      for (int d = 0; d < N; ++d) {
        const float yb = a.bc_pos[b * N + d], vb = __fmul_rn(a.bc_vel[b * N + d], tau);
        float ap = fmaf(x2, vb, fmaf(x1, yb, 0.f)), av = fmaf(x4, vb, fmaf(x3, yb, 0.f));     // table columns 0, 1
        for (int k = 0; k < KG; ++k) {
          ap = fmaf(hp[k], s_w[d * KP + k], ap);
          av = fmaf(hv[k], s_w[d * KP + k], av);
        }
        s_pos[t * N + d] = ap;
        s_vel[t * N + d] = __fdiv_rn(av, tau);
      }
    }
    __syncthreads();
    float* gp = a.pos + b * TM * N;
    float* gv = a.vel + b * TM * N;
    for (int i = threadIdx.x; i < T * N; i += blockDim.x) {
      gp[i] = s_pos[i];
      gv[i] = s_vel[i];
    }
    for (int i = T * N + threadIdx.x; i < TM * N; i += blockDim.x) gp[i] = gv[i] = 0.f;      // ragged: rows past this env's plan
    return;
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float un = __fdiv_rn(__fsub_rn(times[t], delay), tau);         // float32 elementwise ops of the library
    const float z = fminf(fmaxf(un, 0.f), 1.f);
    double phi[16];
    const double x = eval_basis(a, z, phi);
    if (a.mp_kind == FG_MP_PROMP) {
      for (int d = 0; d < N; ++d) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(__fmul_rn((float)phi[a.first + k], a.wscale), s_w[d * KP + k], acc);
        s_pos[t * N + d] = acc;
      }
    } else {
      for (int d = 0; d < N; ++d) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf((float)(x * phi[a.first + k]), s_w[d * KP + k], acc);
        s_f[t * N + d] = acc;
      }
      s_h[t] = fmaxf(un, 0.f);       // left-bounded scaled time; differenced below
    }
  }
  __syncthreads();
  if (a.mp_kind == FG_MP_PROMP) {
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const int ts = (t < T - 1) ? t : T - 2;                 // vel[T-1] = vel[T-2]
      for (int d = 0; d < N; ++d)
        s_vel[t * N + d] = (T > 1) ? __fdiv_rn(__fsub_rn(s_pos[(ts + 1) * N + d], s_pos[ts * N + d]),
                                               __fsub_rn(times[ts + 1], times[ts])) : 0.f;
    }
  } else {
    if (threadIdx.x < N) {          // serial semi-implicit Euler in scaled time, every op rounded separately
      const int d = threadIdx.x;
      const float g = s_w[d * KP + K];
      float y = a.bc_pos[b * N + d], yd = __fmul_rn(a.bc_vel[b * N + d], tau);
      for (int t = 0; t < T; ++t) {
        s_pos[t * N + d] = y;
        s_vel[t * N + d] = __fdiv_rn(yd, tau);
        if (t < T - 1) {
          const float h = __fsub_rn(s_h[t + 1], s_h[t]);
          float acc = __fmul_rn(a.beta, __fsub_rn(g, y));
          acc = __fmul_rn(a.alpha, __fsub_rn(acc, yd));
          acc = __fadd_rn(acc, s_f[t * N + d]);
          yd = __fadd_rn(yd, __fmul_rn(h, acc));
          y = __fadd_rn(y, __fmul_rn(h, yd));
        }
      }
    }
  }
  __syncthreads();
  float* gp = a.pos + b * TM * N;
  float* gv = a.vel + b * TM * N;
  for (int i = threadIdx.x; i < T * N; i += blockDim.x) {
    gp[i] = s_pos[i];
    gv[i] = s_vel[i];
  }
  for (int i = T * N + threadIdx.x; i < TM * N; i += blockDim.x) gp[i] = gv[i] = 0.f;        // ragged: rows past this env's plan
}

cudaError_t launch_trajgen_phase(const PhaseArgs& a, long long B, cudaStream_t stream, int max_smem_optin, const char** why) {
  const int KP = (a.mp_kind == FG_MP_PROMP) ? a.K : a.K + 1;
  const size_t smem = sizeof(float) * ((size_t)3 * a.T * a.N + a.T + (size_t)a.N * KP);
  if (smem > (size_t)max_smem_optin) {
    *why = "trajectory too long for the per-env-phase kernel";
    return cudaSuccess;
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_trajgen_phase, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int threads = ((a.T + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  k_trajgen_phase<<<(unsigned)B, threads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace fg
