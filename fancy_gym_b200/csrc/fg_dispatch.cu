// Kernel selection for the two entry points; trajectory-generation instantiations live here.
#include <cstdlib>

#include "fg_dispatch.h"
#include "fg_trajgen.cuh"

namespace fg {

cudaError_t launch_rollout(const DevCfg& c, int env_kind, int mp_kind, const fg_rollout_io& io, long long B,
                           int seg_steps, cudaStream_t stream, int max_smem_optin, const char** why, unsigned* queue,
                           int sm_count, const PhaseConst* pc) {
#define FG_DOF_SWITCH(env)                                                                                          \
  switch (c.n_dof) {                                                                                                \
    case 2: return launch_rollout_##env##_2(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);             \
    case 3: return launch_rollout_##env##_3(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);             \
    case 4: return launch_rollout_##env##_4(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);             \
    case 5: return launch_rollout_##env##_5(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);             \
    case 6: return launch_rollout_##env##_6(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);             \
    case 7: return launch_rollout_##env##_7(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);             \
    default: *why = "n_links not instantiated for the fused rollout (available: 2..7)"; return cudaSuccess;         \
  }
  switch (env_kind) {
    case FG_ENV_HOLE_REACHER: FG_DOF_SWITCH(hole)
    case FG_ENV_VIAPOINT_REACHER: FG_DOF_SWITCH(viapoint)
    case FG_ENV_SIMPLE_REACHER: FG_DOF_SWITCH(simple)
    case FG_ENV_TOY: return launch_rollout_toy(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
  }
#undef FG_DOF_SWITCH
  *why = "unknown env_kind";
  return cudaSuccess;
}

namespace {
template <int MP, int N, int KW>
cudaError_t launch_closed(const DevCfg& c, const float* params, const float* bc_pos, const float* bc_vel, float* pos_out,
                          float* vel_out, long long B, cudaStream_t stream, int max_smem_optin, int sm_count,
                          const char** why) {
  const int nq = (c.T + 3) / 4;
  const size_t fl = (size_t)kTrajWarps * 2 * kTrajStageFloats + (size_t)nq * traj_rec4(MP, c.cols_a) * 4 +
                    (KW == 0 ? (size_t)kTrajWarps * N * c.cols_a : 0);
  const size_t smem = fl * sizeof(float);
  if (smem > (size_t)max_smem_optin) {
    *why = "tables exceed the shared memory of one SM";
    return cudaSuccess;
  }
  auto kern = k_trajgen_closed<MP, N, KW>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // blocks of `epb` consecutive envs, dynamically scheduled (see the kernel): large enough to amortise the per-block
  // table staging, small enough for >= ~8 blocks per SM
  int epb = 64;
  while (epb > kTrajWarps && (B + epb - 1) / epb < (long long)sm_count * 8) epb >>= 1;
  const char* env_epb = getenv("FG_TRAJ_EPB");
  if (env_epb && atoi(env_epb) > 0) epb = atoi(env_epb);
  const unsigned blocks = (unsigned)((B + epb - 1) / epb);
  kern<<<blocks, kTrajThreads, smem, stream>>>(c, params, bc_pos, bc_vel, pos_out, vel_out, B, epb);
  return cudaGetLastError();
}

template <int N>
cudaError_t launch_dmp(const DevCfg& c, const float* params, const float* bc_pos, const float* bc_vel, float* pos_out,
                       float* vel_out, long long B, cudaStream_t stream, int max_smem_optin, int sm_count, const char** why) {
  constexpr int G = 32 / N;
  const int RA = (c.K + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)kDmpWarps * G * 2 * kDmpChunk * N + (size_t)c.T * RA + c.T);
  if (smem > (size_t)max_smem_optin) {
    *why = "tables exceed the shared memory of one SM";
    return cudaSuccess;
  }
  auto kern = k_trajgen_dmp<N>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // hardware-scheduled blocks of consecutive envs (see launch_closed): a multiple of the envs one block handles per pass
  int epb = kDmpWarps * G * 4;
  while (epb > kDmpWarps * G && (B + epb - 1) / epb < (long long)sm_count * 8) epb -= kDmpWarps * G;
  const unsigned blocks = (unsigned)((B + epb - 1) / epb);
  kern<<<blocks, kDmpThreads, smem, stream>>>(c, params, bc_pos, bc_vel, pos_out, vel_out, B, epb);
  return cudaGetLastError();
}

template <int MP, int N>
cudaError_t launch_closed_k(const DevCfg& c, const float* params, const float* bc_pos, const float* bc_vel,
                            float* pos_out, float* vel_out, long long B, cudaStream_t stream, int max_smem_optin,
                            int sm_count, const char** why) {
  constexpr int KW5 = (MP == FG_MP_PROMP) ? 5 : 8;   // the registry default num_basis = 5 (registry.py:76-125)
  if (c.cols_a == KW5)
    return launch_closed<MP, N, KW5>(c, params, bc_pos, bc_vel, pos_out, vel_out, B, stream, max_smem_optin, sm_count, why);
  return launch_closed<MP, N, 0>(c, params, bc_pos, bc_vel, pos_out, vel_out, B, stream, max_smem_optin, sm_count, why);
}

template <int MP>
cudaError_t launch_closed_n(const DevCfg& c, const float* params, const float* bc_pos, const float* bc_vel,
                            float* pos_out, float* vel_out, long long B, cudaStream_t stream, int max_smem_optin,
                            int sm_count, const char** why) {
#define FG_N(n) \
  case n: return launch_closed_k<MP, n>(c, params, bc_pos, bc_vel, pos_out, vel_out, B, stream, max_smem_optin, sm_count, why);
  switch (c.n_dof) {
    FG_N(1) FG_N(2) FG_N(3) FG_N(4) FG_N(5) FG_N(6) FG_N(7) FG_N(8)
  }
#undef FG_N
  *why = "n_dof out of range";
  return cudaSuccess;
}
}  // namespace

cudaError_t launch_trajgen(const DevCfg& c, int mp_kind, const float* params, const float* bc_pos, const float* bc_vel,
                           float* pos_out, float* vel_out, long long B, cudaStream_t stream, int max_smem_optin,
                           int sm_count, const char** why) {
  if (mp_kind == FG_MP_PROMP)
    return launch_closed_n<FG_MP_PROMP>(c, params, bc_pos, bc_vel, pos_out, vel_out, B, stream, max_smem_optin, sm_count, why);
  if (mp_kind == FG_MP_PRODMP)
    return launch_closed_n<FG_MP_PRODMP>(c, params, bc_pos, bc_vel, pos_out, vel_out, B, stream, max_smem_optin, sm_count, why);
  if (mp_kind == FG_MP_DMP) {
    if (c.K > 16) {
      *why = "DMP trajectory kernel: at most 16 basis functions";
      return cudaSuccess;
    }
#define FG_DMP(n) \
  case n: return launch_dmp<n>(c, params, bc_pos, bc_vel, pos_out, vel_out, B, stream, max_smem_optin, sm_count, why);
    switch (c.n_dof) {
      FG_DMP(1) FG_DMP(2) FG_DMP(3) FG_DMP(4) FG_DMP(5) FG_DMP(6) FG_DMP(7) FG_DMP(8)
    }
#undef FG_DMP
    *why = "n_dof out of range";
    return cudaSuccess;
  }
  *why = "no trajectory generator for this mp_kind";
  return cudaSuccess;
}

}  // namespace fg
