// Fused black-box rollout kernel (K1-K3 of SURVEY.md §2.1): one thread == one environment.
//
// Replaces, for B envs at once, the Python loop of BlackBoxWrapper.step
// (fancy_gym/black_box/black_box_wrapper.py:150-217): trajectory evaluation (get_trajectory
// :96-120 -> mp_pytorch), controller (:176-177), np.clip (:178-179), env.step
// (base_reacher_direct.py:20-38 / base_reacher_torque.py:20-37), reward, termination, TimeLimit
// truncation and reward aggregation (:215-216).
#pragma once
#include "fg_device.cuh"

namespace fg {

#ifndef FG_ROLLOUT_THREADS
#define FG_ROLLOUT_THREADS 128
#endif
constexpr int kRolloutThreads = FG_ROLLOUT_THREADS;

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }

// per-dof weight slots: ProMP K, DMP K + goal, ProDMP [y_b, tau*dy_b, w_0..w_K-1, g]
__host__ __device__ constexpr int weight_slots(int mp, int K) {
  return mp == FG_MP_PROMP ? K : mp == FG_MP_DMP ? K + 1 : mp == FG_MP_PRODMP ? K + 3 : 0;
}

// shared memory layout (floats):
//   [tab_a T * pad4(cols_a)] [tab_b rows_b * pad4(cols_b)] [1/tab_b pad4(rows_b)] [s_m 100] [w  slots * n_dof * blockDim]
__host__ __device__ inline size_t rollout_smem_floats(int T, int cols_a, int rows_b, int cols_b, int w_per_thread,
                                                      int threads) {
  return (size_t)T * pad4(cols_a) + (size_t)rows_b * pad4(cols_b) + pad4(rows_b) + kLinePoints +
         (size_t)w_per_thread * threads;
}

// KC > 0: the number of weighted basis functions is a compile-time constant (the registry default 5): the per-env
//         weights live in REGISTERS and the (float4-padded) table rows are fetched with vector broadcast loads.
// KC == 0: run-time K; weights stay in shared memory (k-major, thread-minor: conflict free).
// DBG: the verbose>=2 variant that also writes the per-step actions / observations / rewards (black_box_wrapper.py:208-213).
#ifndef FG_ROLLOUT_MINB
#define FG_ROLLOUT_MINB 4   // <= 128 registers: 4 blocks of 128 threads per SM (65 536 envs need 443 resident threads per SM)
#endif
template <int ENV, int MP, bool MOTOR, int N, int KC, bool DBG>
__global__ void __launch_bounds__(kRolloutThreads, FG_ROLLOUT_MINB)
k_rollout(const __grid_constant__ DevCfg c, const __grid_constant__ fg_rollout_io io, const long long B,
          const int seg_steps) {
  extern __shared__ __align__(16) float smem[];
  const int T = c.T;
  const int K = (KC > 0) ? KC : c.K;
  const int CA = (KC > 0) ? weight_slots(MP, KC) : c.cols_a;   // table columns == weight slots per dof
  const int RA = pad4(CA), RB = pad4(c.cols_b);
  float* tabA = smem;
  float* tabB = tabA + T * RA;
  float* tabR = tabB + c.rows_b * RB;          // ProMP: reciprocals of the time increments
  float* s_m = tabR + pad4(c.rows_b);
  float* wsm = s_m + kLinePoints;
  const int tid = threadIdx.x, BD = blockDim.x;

  // ---- stage the shared tables (coalesced reads, rows zero-padded to float4) ----
  for (int i = tid; i < T * RA; i += BD) {
    const int r = i / max(RA, 1), col = i - r * RA;
    tabA[i] = (col < c.cols_a) ? c.tab_a[r * c.cols_a + col] : 0.f;
  }
  for (int i = tid; i < c.rows_b * RB; i += BD) {
    const int r = i / max(RB, 1), col = i - r * RB;
    tabB[i] = (col < c.cols_b) ? c.tab_b[r * c.cols_b + col] : 0.f;
  }
  if constexpr (MP == FG_MP_PROMP)
    for (int i = tid; i < c.rows_b; i += BD) tabR[i] = __frcp_rn(c.tab_b[i]);
  for (int i = tid; i < kLinePoints; i += BD)   // float32(numpy.linspace(0,1,100)): i*(1/99) in float64, last forced to 1
    s_m[i] = (i == kLinePoints - 1) ? 1.0f : (float)((double)i * (1.0 / 99.0));

  const long long b0 = (long long)blockIdx.x * BD;
  const long long b = b0 + tid;
  const bool valid = b < B;

  // ---- stage this block's MP parameters: coalesced read of [BD, P], stored slot-major / thread-minor ----
  constexpr bool HAS_W = (MP != FG_MP_TRAJ);
  const int KP = (MP == FG_MP_PROMP) ? K : K + 1;         // params per dof
  const int P = N * KP;
  const int WS = weight_slots(MP, K);                     // slots per dof
  if constexpr (HAS_W) {
    const long long nblk = min((long long)BD, B - b0);
    for (long long f = tid; f < nblk * P; f += BD) {
      const int th = (int)(f / P), idx = (int)(f % P);
      const int d = idx / KP, k = idx % KP;
      float val = io.params[b0 * P + f];
      if constexpr (MP == FG_MP_DMP) val = __fmul_rn(val, (k < K) ? c.wscale : c.gscale);
      const int slot = d * WS + ((MP == FG_MP_PRODMP) ? k + 2 : k);
      wsm[slot * BD + th] = val;
    }
  }
  __syncthreads();
  if (!valid) return;

  if (io.done[b]) {   // episode already over: frozen (oracle/blackbox.py keeps such envs untouched)
    io.ret[b] = 0.0;
    io.length[b] = 0;
    io.flags[b] = 0;
    if (io.flag_bytes)
      for (int i = 0; i < 4; ++i) io.flag_bytes[i * B + b] = 0;
    if (io.prev_obs)
      for (int j = 0; j < c.n_obs_out; ++j) io.obs[b * c.n_obs_out + j] = io.prev_obs[b * c.n_obs_out + j];
    if (io.prev_info)
      for (int j = 0; j < 4; ++j) io.info[b * 4 + j] = io.prev_info[b * 4 + j];
    return;
  }

  // ---- per-env state ----
  // Velocity state.  In the velocity-controlled envs with a float32 action (velocity / position controller) the joint
  // velocity IS the last float32 action (or the zeros of reset), so it is carried as float32: no conversions per step and
  // ten registers less; it is float64 only at the HBM boundary.  PD-controlled / torque envs keep the float64 value.
  constexpr bool VF = !MOTOR && (ENV == FG_ENV_HOLE_REACHER || ENV == FG_ENV_VIAPOINT_REACHER);
  // The register-resident-weights instantiation (KC > 0) is dispatched for velocity / motor control only (position control
  // takes the run-time-K variant, fg_rollout_launch.cuh): without a motor law the action IS the desired velocity, so the desired
  // position is dead weight in the loop — no copies of it, no select per joint, for ProDMP no position contraction at all.
  constexpr bool VEL_ONLY = (KC > 0) && !MOTOR;
  double q[N], v[N];
  float vf[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if constexpr (ENV == FG_ENV_TOY) {   // ToyWrapper: current_pos = 1, current_vel = 0 (test_black_box.py:48-56)
      q[i] = 1.0;
      v[i] = 0.0;
      vf[i] = 0.f;
    } else {
      q[i] = io.q[b * N + i];
      v[i] = io.v[b * N + i];
      vf[i] = (float)v[i];
    }
  }
  int steps = io.steps[b];

  // env context
  Hole hole{};
  double cx0 = 0, cx1 = 0, cx2 = 0, cx3 = 0;
  if constexpr (ENV != FG_ENV_TOY) {
    cx0 = io.ctx[b * 4 + 0]; cx1 = io.ctx[b * 4 + 1]; cx2 = io.ctx[b * 4 + 2]; cx3 = io.ctx[b * 4 + 3];
  }
  double ee180x = 0.0, ee180y = 0.0;     // rew_fct "unbounded": end effector latched at step 180 (hr_unbounded_reward.py:35-36)
  if constexpr (ENV == FG_ENV_HOLE_REACHER) {
    hole.xl = (float)(cx0 - cx1 / 2);    // hole_reacher.py:152 (x - width/2), rounded once to float32
    hole.xr = (float)(cx0 + cx1 / 2);
    hole.nd = (float)(-cx2);
    if (c.rew_fct == 2 && steps > 0) {   // a later plan segment of the same episode: the latch lives in info[2..3]
      ee180x = io.info[b * 4 + 2];
      ee180y = io.info[b * 4 + 3];
    }
  }

  // full step observation (float64 trig, cast to float32 like _get_obs); returns its width
  auto build_obs = [&](float* obs, double& ex, double& ey) -> int {
    int no = 0;
    if constexpr (ENV == FG_ENV_TOY) {
      obs[no++] = -1.0f;
      ex = ey = 0.0;
    } else {
      double th[N];
      th[0] = q[0];
#pragma unroll
      for (int i = 1; i < N; ++i) th[i] = th[i - 1] + q[i];
      end_effector64<N>(th, ex, ey);
#pragma unroll
      for (int i = 0; i < N; ++i) obs[i] = (float)cos(q[i]);
#pragma unroll
      for (int i = 0; i < N; ++i) obs[N + i] = (float)sin(q[i]);
#pragma unroll
      for (int i = 0; i < N; ++i) obs[2 * N + i] = VF ? vf[i] : (float)v[i];
      no = 3 * N;
      if constexpr (ENV == FG_ENV_HOLE_REACHER) {
        obs[no++] = (float)cx1;
        obs[no++] = (float)(ex - cx0);
        obs[no++] = (float)(ey - (-cx2));
      } else if constexpr (ENV == FG_ENV_VIAPOINT_REACHER) {
        obs[no++] = (float)(ex - cx0);
        obs[no++] = (float)(ey - cx1);
        obs[no++] = (float)(ex - cx2);
        obs[no++] = (float)(ey - cx3);
      } else {
        obs[no++] = (float)(ex - cx0);
        obs[no++] = (float)(ey - cx1);
      }
      obs[no++] = (float)steps;
    }
    if (c.time_aware) obs[no++] = (float)((double)steps / (double)c.max_steps);
    return no;
  };

  // boundary condition of the plan (black_box_wrapper.py:110-114), float32 like the library
  float ybc[N], vbc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (io.use_cond) {
      ybc[i] = io.cond_pos[b * N + i];
      vbc[i] = io.cond_vel[b * N + i];
    } else {
      ybc[i] = (float)q[i];
      vbc[i] = VF ? vf[i] : (float)v[i];
    }
  }

  // ---- per-env weights: registers (KC > 0) or shared memory (KC == 0) ----
  constexpr int WSC = (KC > 0) ? weight_slots(MP, KC) : 1;
  constexpr int RAC = pad4(WSC) > 0 ? pad4(WSC) : 4;
  float wreg[N][(KC > 0 && HAS_W) ? WSC : 1];
#define WSM(d, k) wsm[((d) * WS + (k)) * BD + tid]
  if constexpr (MP == FG_MP_PRODMP) {
#pragma unroll
    for (int d = 0; d < N; ++d) {
      WSM(d, 0) = ybc[d];
      WSM(d, 1) = __fmul_rn(vbc[d], c.tau);
      if (c.rel_goal) WSM(d, K + 2) = __fadd_rn(WSM(d, K + 2), ybc[d]);
    }
  }
  if constexpr (KC > 0 && HAS_W) {
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
      for (int k = 0; k < WSC; ++k) wreg[d][k] = WSM(d, k);
  }
  // dot(table row, weights of dof d): FMA chain in index order, accumulator starts at 0 (oracle 'mirror' mode)
  auto dot_row = [&](const float* row, int d, int n) -> float {
    float acc = 0.f;
    if constexpr (KC > 0) {
      float r[RAC];
#pragma unroll
      for (int j = 0; j < RAC / 4; ++j) {
        const float4 x = reinterpret_cast<const float4*>(row)[j];
        r[4 * j] = x.x; r[4 * j + 1] = x.y; r[4 * j + 2] = x.z; r[4 * j + 3] = x.w;
      }
#pragma unroll
      for (int k = 0; k < WSC; ++k)
        if (k < n) acc = fmaf(r[k], wreg[d][k], acc);
    } else {
      for (int k = 0; k < n; ++k) acc = fmaf(row[k], WSM(d, k), acc);
    }
    return acc;
  };
  auto weight = [&](int d, int k) -> float {
    if constexpr (KC > 0) return wreg[d][k]; else return WSM(d, k);
  };

  // ---- MP set-up ----
  float dmp_y[N], dmp_yd[N];        // DMP integrator state (scaled-time velocity)
  float pos_next[N];                // ProMP: pos[t+1] carried to the next step
  float pos[N], vel[N];             // desired position / velocity of the current step
  const float r_tau = __frcp_rn(c.tau), r_dt = __frcp_rn(c.dt_f);
  if constexpr (MP == FG_MP_DMP) {
#pragma unroll
    for (int d = 0; d < N; ++d) {
      dmp_y[d] = ybc[d];
      dmp_yd[d] = __fmul_rn(vbc[d], c.tau);
    }
  }
  if constexpr (MP == FG_MP_PROMP) {
#pragma unroll
    for (int d = 0; d < N; ++d) {
      pos_next[d] = dot_row(tabA, d, K);
      vel[d] = 0.f;                 // (a one-point plan has zero velocity; otherwise overwritten at t = 0)
    }
  }

  double ret = 0.0;
  int t = 0;
  bool terminated = false, truncated = false, success = false, collided = false;
  double info0 = 0, info1 = 0;

  const int my_steps = io.seg_steps_env ? min(seg_steps, io.seg_steps_env[b]) : seg_steps;      // ragged sub-trajectories
  for (; t < my_steps; ++t) {
    // ------------------------------------------------------------------ desired pos / vel at point t
    if constexpr (MP == FG_MP_PROMP) {
#pragma unroll
      for (int d = 0; d < N; ++d) pos[d] = pos_next[d];
      if (t < T - 1) {
        const float* row = tabA + (t + 1) * RA;
        const float dtt = tabB[t * RB], rdt = tabR[t];
#pragma unroll
        for (int d = 0; d < N; ++d) {
          const float acc = dot_row(row, d, K);
          pos_next[d] = acc;
          vel[d] = div_by(__fsub_rn(acc, pos[d]), dtt, rdt);     // (pos[t+1]-pos[t]) / (times[t+1]-times[t])
        }
      }                     // t == T - 1: vel[T-1] = vel[T-2] — the registers simply keep the previous step's values
    } else if constexpr (MP == FG_MP_DMP) {
#pragma unroll
      for (int d = 0; d < N; ++d) {
        pos[d] = dmp_y[d];
        vel[d] = div_by(dmp_yd[d], c.tau, r_tau);
      }
      if (t < T - 1) {   // semi-implicit Euler in scaled time (oracle/mp.py DMP._integrate)
        const float* row = tabA + t * RA;
        const float h = tabB[t * RB];
#pragma unroll
        for (int d = 0; d < N; ++d) {
          const float f = dot_row(row, d, K);
          const float g = weight(d, K);
          float a = __fmul_rn(c.beta, __fsub_rn(g, dmp_y[d]));
          a = __fmul_rn(c.alpha, __fsub_rn(a, dmp_yd[d]));
          a = __fadd_rn(a, f);
          dmp_yd[d] = __fadd_rn(dmp_yd[d], __fmul_rn(h, a));
          dmp_y[d] = __fadd_rn(dmp_y[d], __fmul_rn(h, dmp_yd[d]));
        }
      }
    } else if constexpr (MP == FG_MP_PRODMP) {
      const float* rp = tabA + t * RA;
      const float* rv = tabB + t * RB;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        pos[d] = dot_row(rp, d, K + 3);
        vel[d] = div_by(dot_row(rv, d, K + 3), c.tau, r_tau);
      }
    } else {
#pragma unroll
      for (int d = 0; d < N; ++d) {
        pos[d] = io.traj_pos[(b * T + t) * N + d];
        vel[d] = io.traj_vel[(b * T + t) * N + d];
      }
    }

    // ------------------------------------------------------------------ controller + clip + dynamics
    double a64[N];
    float a32[N];
    double acc_cost = 0.0;    // sum(acc^2) (direct envs), in the reference's dtype
    if constexpr (MOTOR) {    // pd_controller.py:28: float64 because c_pos / c_vel are float64
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double tq = __dadd_rn(__dmul_rn(c.p[i], (double)pos[i] - q[i]), __dmul_rn(c.d[i], (double)vel[i] - v[i]));
        a64[i] = fmin(fmax(tq, -(double)c.act_lim), (double)c.act_lim);
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const float des = (VEL_ONLY || c.ctrl == FG_CTRL_VELOCITY) ? vel[i] : pos[i];
        a32[i] = fminf(fmaxf(des, -c.act_lim), c.act_lim);
        a64[i] = (double)a32[i];
      }
    }

    if constexpr (ENV == FG_ENV_HOLE_REACHER || ENV == FG_ENV_VIAPOINT_REACHER) {
      // base_reacher_direct.py:25-27
      if (MOTOR || steps == 0) {      // v is float64 (zeros at reset / float64 actions): float64 arithmetic
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double ac = (a64[i] - (VF ? (double)vf[i] : v[i])) / c.dt;
          acc_cost += ac * ac;
        }
      } else {                        // float32 action and float32 velocity
        float s32 = 0.f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const float ac = div_by(__fsub_rn(a32[i], VF ? vf[i] : (float)v[i]), c.dt_f, r_dt);
          s32 = __fadd_rn(s32, __fmul_rn(ac, ac));
        }
        acc_cost = (double)s32;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if constexpr (VF) vf[i] = a32[i]; else v[i] = a64[i];
        q[i] += MOTOR ? __dmul_rn(c.dt, a64[i]) : (double)__fmul_rn(c.dt_f, a32[i]);
      }
    } else if constexpr (ENV == FG_ENV_SIMPLE_REACHER) {
      // base_reacher_torque.py:25-26
#pragma unroll
      for (int i = 0; i < N; ++i) {
        v[i] += MOTOR ? __dmul_rn(c.dt, a64[i]) : (double)__fmul_rn(c.dt_f, a32[i]);
        q[i] += __dmul_rn(c.dt, v[i]);
      }
    }

    // ------------------------------------------------------------------ geometry, collisions, reward
    double reward = 0.0;
    if constexpr (ENV == FG_ENV_TOY) {
      reward = 1.0;
    } else {
      double th[N];
      th[0] = q[0];
#pragma unroll
      for (int i = 1; i < N; ++i) th[i] = th[i - 1] + q[i];

      if constexpr (ENV == FG_ENV_HOLE_REACHER) {
        float cs[N], sn[N];
#pragma unroll
        for (int i = 0; i < N; ++i) sincos_reduced(th[i], sn[i], cs[i]);
        bool selfc = false, wallc = false;
        if (!c.allow_self) selfc = self_collision<N>(q, th, cs, sn);
        if (!c.allow_wall) wallc = wall_collision<N>(s_m, cs, sn, hole, c.wall_mode);
        collided = selfc | wallc;
        success = false;
        if (c.rew_fct == 0) {
          // hr_simple_reward.py:35-53
          // ordinary steps: (-0.0 + x) + -0.0 == x for x = acc_cost * -5e-8 <= 0, so only the product is formed
          reward = __dmul_rn(acc_cost, -5e-8);
          if (steps == 199 || collided) {
            double ex, ey;
            end_effector64<N>(th, ex, ey);
            const double dx = ex - cx0, dy = ey - (-cx2);
            const double dist = sqrt(dx * dx + dy * dy);
            const double dist_cost = dist * dist;
            const double coll_cost = collided ? 1.0 : 0.0;
            success = (dist < 0.005) && !collided;
            reward = __dadd_rn(__dadd_rn(__dmul_rn(dist_cost, -1.0), reward), __dmul_rn(coll_cost, -c.penalty));
          }
        } else if (c.rew_fct == 1) {
          // hr_dist_vel_acc_reward.py:20-60: distance / collision terms only on step 199 (a collision ends the episode, so
          // the latched flag and collision_dist are this step's); factors (-1, -1e-4, -1e-6, -penalty, 0)
          double dist_cost = 0.0, coll_cost = 0.0;
          if (steps == 199) {
            double ex, ey;
            end_effector64<N>(th, ex, ey);
            const double dx = ex - cx0, dy = ey - (-cx2);
            const double dist = sqrt(dx * dx + dy * dy);
            dist_cost = dist * dist;
            coll_cost = collided ? dist * dist : 0.0;
            success = (dist < 0.005) && !collided;
          }
          double vel_cost;
          if constexpr (MOTOR) {
            vel_cost = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) vel_cost += a64[i] * a64[i];
          } else {      // float32 action: np.sum(v ** 2) is a float32 sum
            float s32 = 0.f;
#pragma unroll
            for (int i = 0; i < N; ++i) s32 = __fadd_rn(s32, __fmul_rn(a32[i], a32[i]));
            vel_cost = (double)s32;
          }
          reward = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(dist_cost, -1.0), __dmul_rn(vel_cost, -1e-4)),
                                       __dmul_rn(acc_cost, -1e-6)), __dmul_rn(coll_cost, -c.penalty));
        } else {
          // hr_unbounded_reward.py:17-60: end effector latched at step 180 (or on collision); factors (1, -5e-6)
          double dist_reward = 0.0;
          if (steps == 180 || steps == 199 || collided) {
            double ex, ey;
            end_effector64<N>(th, ex, ey);
            if (steps == 180 || collided) {
              ee180x = ex;
              ee180y = ey;
            }
            if (steps == 199 || collided) {
              const double dx = ee180x - cx0, dy = ee180y - (-cx2);
              const double dist = sqrt(dx * dx + dy * dy);
              if (collided) dist_reward = 0.25 * exp(-dist);
              else if (ey > 0) dist_reward = exp(-dist);
              else dist_reward = 1 - ee180y;
              success = !collided;
            }
          }
          reward = __dadd_rn(dist_reward, __dmul_rn(acc_cost, -5e-6));
        }
        terminated = collided;
      } else if constexpr (ENV == FG_ENV_VIAPOINT_REACHER) {
        // viapoint_reacher.py:79-107 (App. A.6-Q1/Q2: -inf start, `acc` is the action)
        collided = false;
        if (!c.allow_self) {
          float cs[N], sn[N];
#pragma unroll
          for (int i = 0; i < N; ++i) sincos_reduced(th[i], sn[i], cs[i]);
          collided = self_collision<N>(q, th, cs, sn);
        }
        double dist = INFINITY;
        reward = -INFINITY;
        success = false;
        if (!collided) {
          if (steps == 100 || steps == 199) {
            double ex, ey;
            end_effector64<N>(th, ex, ey);
            const double tx = (steps == 100) ? cx0 : cx2, ty = (steps == 100) ? cx1 : cx3;
            dist = sqrt((ex - tx) * (ex - tx) + (ey - ty) * (ey - ty));
          }
          success = dist < 0.005;
        } else {
          double ex, ey;
          end_effector64<N>(th, ex, ey);
          dist = sqrt((ex - cx2) * (ex - cx2) + (ey - cx3) * (ey - cx3));
          reward = -c.penalty;
        }
        reward -= dist * dist;
        if constexpr (MOTOR) {
          double asq = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) asq += a64[i] * a64[i];
          reward -= 5e-8 * asq;
        } else {   // float32 action: np.sum(acc**2) is float32 and 5e-8 * float32 stays float32
          float s32 = 0.f;
#pragma unroll
          for (int i = 0; i < N; ++i) s32 = __fadd_rn(s32, __fmul_rn(a32[i], a32[i]));
          reward -= (double)__fmul_rn(5e-8f, s32);
        }
        terminated = collided;
      } else {   // SIMPLE_REACHER: simple_reacher.py:56-70 (collision flag is computed but unused)
        double rdist = 0.0;
        if (steps >= 199) {
          double ex, ey;
          end_effector64<N>(th, ex, ey);
          rdist = -sqrt((ex - cx0) * (ex - cx0) + (ey - cx1) * (ey - cx1));
        }
        double rctrl;
        if constexpr (MOTOR) {
          rctrl = 0.0;
#pragma unroll
          for (int i = 0; i < N; ++i) rctrl += a64[i] * a64[i];
          reward = rdist - rctrl;
        } else {   // float32 action: reward_ctrl float32; `0 - float32` stays float32 before steps>=199
          float s32 = 0.f;
#pragma unroll
          for (int i = 0; i < N; ++i) s32 = __fadd_rn(s32, __fmul_rn(a32[i], a32[i]));
          rctrl = (double)s32;
          reward = (steps >= 199) ? rdist - rctrl : (double)(0.f - s32);
        }
        info0 = rdist;
        info1 = rctrl;
        terminated = false;
      }
    }
    steps += 1;
    truncated = steps >= c.max_steps;       // gymnasium TimeLimit (App. A.6-Q12)
    ret += reward;

    if constexpr (DBG) {
      if (io.dbg_rewards) io.dbg_rewards[b * T + t] = reward;
      if (io.dbg_actions) {
#pragma unroll
        for (int i = 0; i < N; ++i) io.dbg_actions[(b * T + t) * N + i] = a64[i];
      }
      if (io.dbg_obs) {
        float so[FG_MAX_OBS];
        double ex, ey;
        const int n_full = build_obs(so, ex, ey);
        for (int j = 0; j < n_full; ++j) io.dbg_obs[(b * T + t) * n_full + j] = so[j];
      }
    }
    if (terminated || truncated) {
      ++t;
      break;
    }
  }
  const int len = t;          // executed steps
  const bool stopped = terminated || truncated;

  // ---- write back state and results ----
  if (!io.keep_state) {
    if constexpr (ENV != FG_ENV_TOY) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        io.q[b * N + i] = q[i];
        io.v[b * N + i] = VF ? (double)vf[i] : v[i];
      }
    }
    io.steps[b] = steps;
    io.done[b] = stopped ? 1 : 0;
  }
  if (io.write_cond && len > 0 && (stopped || io.write_cond == 2)) {
    if constexpr (VEL_ONLY && (MP == FG_MP_PROMP || MP == FG_MP_PRODMP)) {
      // the desired position is not carried through the loop in this instantiation: the same FMA chain, once, here
      const float* row = tabA + (len - 1) * RA;
#pragma unroll
      for (int i = 0; i < N; ++i) pos[i] = dot_row(row, i, MP == FG_MP_PROMP ? K : K + 3);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      io.cond_pos[b * N + i] = pos[i];
      io.cond_vel[b * N + i] = vel[i];
    }
  }
  io.ret[b] = ret;
  io.length[b] = len;
  io.flags[b] = (terminated ? FG_FLAG_TERMINATED : 0u) | (truncated ? FG_FLAG_TRUNCATED : 0u) |
                (success ? FG_FLAG_SUCCESS : 0u) | (collided ? FG_FLAG_COLLIDED : 0u);
  if (io.flag_bytes) {
    io.flag_bytes[b] = terminated;
    io.flag_bytes[B + b] = truncated;
    io.flag_bytes[2 * B + b] = success;
    io.flag_bytes[3 * B + b] = collided;
  }

  // observation after the last executed step
  float obs[FG_MAX_OBS];
  {
    double ex, ey;
    build_obs(obs, ex, ey);
    if constexpr (ENV == FG_ENV_HOLE_REACHER || ENV == FG_ENV_VIAPOINT_REACHER) {
      info0 = ex;
      info1 = ey;
    }
  }
  for (int j = 0; j < c.n_obs_out; ++j) io.obs[b * c.n_obs_out + j] = obs[c.obs_index[j]];
  io.info[b * 4 + 0] = info0;
  io.info[b * 4 + 1] = info1;
  io.info[b * 4 + 2] = ee180x;
  io.info[b * 4 + 3] = ee180y;
#undef WSM
}

}  // namespace fg
