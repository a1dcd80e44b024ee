// Rollout kernel instantiations for FG_ENV_TOY (one translation unit per env: parallel builds).
#include "fg_rollout_launch.cuh"
namespace fg {
FG_DECL_ENV_LAUNCH(launch_rollout_toy) {
  switch (c.n_dof) {
    case 1: return launch_mp_ctrl<FG_ENV_TOY, 1>(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
    case 2: return launch_mp_ctrl<FG_ENV_TOY, 2>(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
    case 5: return launch_mp_ctrl<FG_ENV_TOY, 5>(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
  }
  *why = "n_dof not instantiated for this env (available: 1 2 5)";
  return cudaSuccess;
}
}  // namespace fg
