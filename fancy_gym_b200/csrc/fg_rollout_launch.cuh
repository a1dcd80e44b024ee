// Shared launch helper for the per-env rollout translation units.
#pragma once
#include "fg_dispatch.h"
#include <cstdlib>
#include <cstring>

#include "fg_rollout.cuh"

namespace fg {

template <int ENV, int MP, bool MOTOR, int N, int KC, bool DBG, int NTX = 0>
cudaError_t launch_dbg(const DevCfg& c, const fg_rollout_io& io, long long B, int seg_steps, cudaStream_t stream,
                      int max_smem_optin, const char** why, unsigned* queue = nullptr, int sm_count = 0,
                      const PhaseConst* pc = nullptr) {
  const int pw = N * weight_slots(MP, c.K);
  const size_t smem = rollout_smem_bytes(c.T, c.cols_a, c.rows_b, c.cols_b, pw, SlotLayout<ENV, MP, MOTOR, N, KC>::WORDS,
                                         kRolloutThreads);
  if (smem > (size_t)max_smem_optin) {
    *why = "tables + per-env weights and state exceed the shared memory of one SM (reduce n_steps or n_basis)";
    return cudaSuccess;
  }
  auto kern = k_rollout<ENV, MP, MOTOR, N, KC, DBG, NTX>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  long long blocks = (B + kRolloutThreads - 1) / kRolloutThreads;
  fg_rollout_io iok = io;
  if (iok.n_plans <= 1) {      // one plan: the whole table is that plan
    iok.n_plans = 1;
    iok.plan_T = c.T;
    iok.plan_row0[0] = 0;
  }
  // More envs than the GPU holds at once: a persistent grid (every SM full) whose threads take the next env from a work
  // queue when the one they ran is finished, instead of blocks that drain while their last envs run out.
  unsigned* q = nullptr;
  int warps_lo = kRolloutWarps, blocks_extra = 0;
  if (queue) {
    static thread_local size_t occ_smem = ~(size_t)0;
    static thread_local int occ_blocks = 0;
    if (occ_smem != smem) {
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, kern, kRolloutThreads, smem);
      if (e != cudaSuccess) return e;
      occ_smem = smem;
    }
    const long long resident = (long long)occ_blocks * sm_count;
    static const bool no_queue = getenv("FG_NO_QUEUE") != nullptr;      // (A/B measurements only)
    if (resident > 0 && blocks > 2 * resident && !no_queue) {
      blocks = resident;
      q = queue;
    }
    // (cutting a batch that does not fill the GPU into 592 blocks of 3 or 4 warps, so that every SM holds 13 - 14 warps instead
    //  of 12 or 16, was measured: no change — 0.259 ms either way at 65 536 envs; the warps progress at their own pace)
  }
  PhaseConst pcv;
  if (pc) pcv = *pc; else memset(&pcv, 0, sizeof(pcv));
  kern<<<(unsigned)blocks, kRolloutThreads, smem, stream>>>(c, iok, B, seg_steps, q, warps_lo, blocks_extra, pcv);
  return cudaGetLastError();
}

template <int ENV, int MP, bool MOTOR, int N, int KC>
cudaError_t launch_kc(const DevCfg& c, const fg_rollout_io& io, long long B, int seg_steps, cudaStream_t stream,
                      int max_smem_optin, const char** why, unsigned* queue = nullptr, int sm_count = 0,
                      const PhaseConst* pc = nullptr) {
  // the per-step debug outputs of verbose >= 2 are a separate instantiation: the hot variant carries none of their code.
  // They always use run-time K (KC = 0) to keep the number of instantiations down.
  if (io.dbg_actions || io.dbg_obs || io.dbg_rewards || io.dbg_state)
    return launch_dbg<ENV, MP, MOTOR, N, 0, true>(c, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
  return launch_dbg<ENV, MP, MOTOR, N, KC, false>(c, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
}

template <int ENV, int MP, bool MOTOR, int N>
cudaError_t launch_one(const DevCfg& c, const fg_rollout_io& io, long long B, int seg_steps, cudaStream_t stream,
                       int max_smem_optin, const char** why, unsigned* queue = nullptr, int sm_count = 0,
                      const PhaseConst* pc = nullptr) {
  // num_basis = 5 is the registry default of every MP type (registry.py:76-125): register-resident weights
  // (instantiated for the registered link counts only — 5 links, SimpleReacher's 2 — to keep the library small)
  if (pc && pc->n_total > 0) {
    // per-env phase evaluated inside the rollout: instantiated for the registry's shapes (5 weighted RBFs of 5 or 6 in total,
    // 5 or 2 links, velocity / motor control); everything else goes through fg_trajgen_phase + FG_MP_TRAJ
    if constexpr ((MP == FG_MP_PROMP || MP == FG_MP_DMP) && (N == 5 || N == 2)) {
      const bool ok = c.K == 5 && (MOTOR || c.ctrl == FG_CTRL_VELOCITY) && pc->first == pc->n_total - 5 &&
                      !(io.dbg_actions || io.dbg_obs || io.dbg_rewards || io.dbg_state) && io.n_plans <= 1;
      if (ok && pc->n_total == 5)
        return launch_dbg<ENV, MP, MOTOR, N, 5, false, 5>(c, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
      if (ok && pc->n_total == 6)
        return launch_dbg<ENV, MP, MOTOR, N, 5, false, 6>(c, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
    }
    *why = "per-env phase inside the rollout is not instantiated for this shape (use fg_trajgen_phase + FG_MP_TRAJ)";
    return cudaSuccess;
  }
  if constexpr (MP != FG_MP_TRAJ && (N == 5 || N == 2)) {
    // (without a motor law the KC instantiation assumes velocity control: position control takes the run-time-K variant)
    if (c.K == 5 && (MOTOR || c.ctrl == FG_CTRL_VELOCITY))
      return launch_kc<ENV, MP, MOTOR, N, 5>(c, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
  }
  return launch_kc<ENV, MP, MOTOR, N, 0>(c, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
}

template <int ENV, int N>
cudaError_t launch_mp_ctrl(const DevCfg& c, int mp_kind, const fg_rollout_io& io, long long B, int seg_steps,
                           cudaStream_t stream, int max_smem_optin, const char** why, unsigned* queue = nullptr, int sm_count = 0,
                      const PhaseConst* pc = nullptr) {
  const bool motor = c.ctrl == FG_CTRL_MOTOR;
#define FG_CASE(MPK)                                                                                             \
  case MPK:                                                                                                      \
    return motor ? launch_one<ENV, MPK, true, N>(c, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc)  \
                 : launch_one<ENV, MPK, false, N>(c, io, B, seg_steps, stream, max_smem_optin, why, queue, sm_count, pc);
  switch (mp_kind) {
    FG_CASE(FG_MP_PROMP)
    FG_CASE(FG_MP_DMP)
    FG_CASE(FG_MP_PRODMP)
    FG_CASE(FG_MP_TRAJ)
  }
#undef FG_CASE
  *why = "unknown mp_kind";
  return cudaSuccess;
}

}  // namespace fg
