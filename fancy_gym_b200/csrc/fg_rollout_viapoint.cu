// Rollout kernel instantiations for FG_ENV_VIAPOINT_REACHER (one translation unit per env: parallel builds).
#include "fg_rollout_launch.cuh"
namespace fg {
FG_DECL_ENV_LAUNCH(launch_rollout_viapoint) {
  switch (c.n_dof) {
    case 5: return launch_mp_ctrl<FG_ENV_VIAPOINT_REACHER, 5>(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why);
  }
  *why = "n_dof not instantiated for this env (available: 5)";
  return cudaSuccess;
}
}  // namespace fg
