// Host-side launch dispatch: picks the kernel instantiation for (env, mp, controller dtype, dof).
#pragma once
#include <cuda_runtime.h>

#include "fg_device.cuh"

namespace fg {

// Returns the CUDA error of the launch; *why is set (and nothing is launched) for a combination
// that is not instantiated.
// queue: two zeroed device words (work queue of a persistent grid, reset by the kernel itself) or null
cudaError_t launch_rollout(const DevCfg& c, int env_kind, int mp_kind, const fg_rollout_io& io, long long B,
                           int seg_steps, cudaStream_t stream, int max_smem_optin, const char** why, unsigned* queue,
                           int sm_count, const PhaseConst* pc);

cudaError_t launch_trajgen(const DevCfg& c, int mp_kind, const float* params, const float* bc_pos, const float* bc_vel,
                           float* pos_out, float* vel_out, long long B, cudaStream_t stream, int max_smem_optin,
                           int sm_count, const char** why);

// trajectory covariance (fg_cov.cu); path 1 = CUDA cores, 2 = tcgen05
struct CovArgs {
  const float* basis;   // device [T, ld] float32, columns c0 .. c0+Kc-1 are Bm
  int ld, c0, Kc, T, N; // N = dof
  const float* L;       // [B, D, D]
  float* cov;           // [B, N*T, N*T] or null
  float* stdv;          // [B, T, N] or null
  float* diag;          // scratch [B, N*T]
  float* envmax;        // scratch [B]
  float* gmax;          // scratch [1]
  float reg;
  int batch_scope;      // 1: regulariser uses the max over the whole batch (mp_pytorch's torch.max over the batched tensor)
  int rows_per_block;
};

cudaError_t launch_traj_cov(const CovArgs& a, long long B, int path, cudaStream_t stream, int max_smem_optin, const char** why);

// trajectory generation with a per-env phase (fg_trajgen_phase.cu)
struct PhaseArgs {
  int mp_kind, N, T, K;             // K weighted basis functions per dof
  int phase_kind, n_total, first;   // phase 0 linear / 1 exp; RBF count incl. zero padding; first weighted RBF
  int exp_right_clip;               // exponential phase of the clipped (1) or only left-bounded (0) linear phase
  int eval_f64;                     // basis in float64 rounded once (1) or float32 elementwise like the library (0)
  float cen32[16], bw32[16];        // float32 copies of the centres / bandwidths (eval_f64 == 0)
  double alpha_phase, basis_scale;  // basis_scale: DMP forcing-basis factor (1 unless weights_scale sits on the basis)
  double cen[16], bw[16];
  RbfRec rec;                       // ProMP with a linear phase: two exp() per time point (fg_device.cuh)
  float wscale, gscale, alpha, beta;
  const float* times;               // [T] float32 time grid (device)
  const float* dts;                 // [T-1] its increments (ProMP velocity), device
  const float* tau;                 // [B]
  const float* delay;               // [B]
  const float* params;              // [B, N * (K or K+1)]
  const float* bc_pos;              // [B, N] (DMP)
  const float* bc_vel;
  float* pos;                       // [B, T, N]
  float* vel;
  // ProDMP: pre-integrated bases (float64, device), index rounding step, boundary time, parameter scales
  const double* pc_pos;
  const double* pc_vel;
  const double* pc_y;
  int n_pc, rel_goal;
  float scaled_dt, init_time;
  double scale[17];
  // ragged plans: per-env number of points and the table of time grids (one row per length), or null
  const int* n_steps_env;
  const float* times_table;
  int times_stride;
};

cudaError_t launch_trajgen_phase(const PhaseArgs& a, long long B, cudaStream_t stream, int max_smem_optin, const char** why);

// per-env translation units (compiled in parallel)
#define FG_DECL_ENV_LAUNCH(name)                                                                              \
  cudaError_t name(const DevCfg& c, int mp_kind, const fg_rollout_io& io, long long B, int seg_steps,        \
                   cudaStream_t stream, int max_smem_optin, const char** why, unsigned* queue, int sm_count,        \
                   const PhaseConst* pc)
#define FG_DECL_ENV_DOFS(env)                                                                              \
  FG_DECL_ENV_LAUNCH(launch_rollout_##env##_2); FG_DECL_ENV_LAUNCH(launch_rollout_##env##_3);                 \
  FG_DECL_ENV_LAUNCH(launch_rollout_##env##_4); FG_DECL_ENV_LAUNCH(launch_rollout_##env##_5);                 \
  FG_DECL_ENV_LAUNCH(launch_rollout_##env##_6); FG_DECL_ENV_LAUNCH(launch_rollout_##env##_7);
FG_DECL_ENV_DOFS(hole)
FG_DECL_ENV_DOFS(viapoint)
FG_DECL_ENV_DOFS(simple)
FG_DECL_ENV_LAUNCH(launch_rollout_toy);

}  // namespace fg
