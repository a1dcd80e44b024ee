// Rollout kernel instantiations: FG_ENV_VIAPOINT_REACHER with 2 links (one translation unit per (env, dof): parallel builds).
#include "fg_rollout_launch.cuh"
namespace fg {
FG_DECL_ENV_LAUNCH(launch_rollout_viapoint_2) {
  return launch_mp_ctrl<FG_ENV_VIAPOINT_REACHER, 2>(c, mp_kind, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
}
}  // namespace fg
