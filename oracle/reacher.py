"""numpy restatement of the reference's classic_control reacher envs, batched over envs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED against the reference's own files (run
unmodified via oracle/ref_loader.py) by tests/golden/make_golden.py -> tests/golden/env_*.npz.

Every method cites the reference lines it restates (paths relative to
/root/reference/fancy_gym/envs/classic_control/).  Arithmetic keeps the reference's *mixed*
precision: actions arrive as float32 (they come out of torch), `dt * v`
is rounded to float32 by numpy's weak-scalar promotion, joint angles / FK / collision tests /
rewards are float64, observations are cast to float32.  Writing the same numpy expressions on
arrays with a leading batch axis reproduces those promotions automatically.

Besides the reference outputs the oracle reports a *decision margin* per step: an L-infinity
estimate of how far the geometric quantities are from flipping a collision / joint-limit
decision.  The fp32 CUDA path is required to reproduce flags and step counts exactly except
where this margin is below a documented epsilon ("boundary ties").
"""
from __future__ import annotations

import numpy as np

DT = 0.01                      # base_reacher/base_reacher.py:21
N_LINE_POINTS = 100            # hole_reacher/hole_reacher.py:149
CCW_EPS = 1e-12                # utils.py:2


# ----------------------------------------------------------------------------------------------
# reset samplers (numpy-exact: same Generator construction and draw order as the reference)
# ----------------------------------------------------------------------------------------------
def _rng(seed):
    # gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def sample_first_joint(rng):
    # base_reacher/base_reacher.py:82
    return rng.uniform(np.pi / 4, 3 * np.pi / 4)


def sample_hole_context(seed, hole_width=None, hole_x=None, hole_depth=1, random_start=True):
    """hole_reacher/hole_reacher.py:60-71 (reset), :79-112 (_generate_hole); then
    base_reacher/base_reacher.py:73-93 on the same stream (reset called with seed=None)."""
    rng = _rng(seed)
    width = rng.uniform(0.15, 0.5) if hole_width is None else float(hole_width)
    if hole_x is None:
        direction = rng.choice([-1, 1])
        x = direction * rng.uniform(width / 2, 3.5)
    else:
        x = float(hole_x)
    depth = rng.uniform(1, 1) if hole_depth is None else hole_depth
    q0 = sample_first_joint(rng) if random_start else np.pi / 2
    return dict(x=float(x), width=float(width), depth=float(depth), q0=float(q0))


def sample_viapoint_context(seed, n_links=5, via_target=None, target=None, random_start=False):
    """viapoint_reacher/viapoint_reacher.py:45-77.  reset() = _generate_goal() [stale stream,
    discarded] -> seeded reset (draws the start angle if random_start) -> _generate_goal() [continues
    that stream] -> seeded reset again, so the start angle is the *first* variate of the stream
    and the goal is drawn from the variates after it (App. A.6-Q4)."""
    total = float(n_links)
    rng = _rng(seed)
    if random_start:
        sample_first_joint(rng)     # first seeded reset already consumed one variate
    if via_target is None:
        via = np.array([total, total])
        while np.linalg.norm(via) >= 0.5 * total:
            via = rng.uniform(low=-0.5 * total, high=0.5 * total, size=2)
    else:
        via = np.array(via_target, dtype=np.float64)
    if target is None:
        goal = np.array([total, total])
        while np.linalg.norm(goal) >= total or np.linalg.norm(goal) <= 0.5 * total:
            goal = rng.uniform(low=-total, high=total, size=2)
    else:
        goal = np.array(target, dtype=np.float64)
    q0 = sample_first_joint(_rng(seed)) if random_start else np.pi / 2
    return dict(via=via, goal=goal, q0=float(q0))


def sample_simple_context(seed, n_links=2, target=None, random_start=True):
    """simple_reacher/simple_reacher.py:46-54, :85-96 (same double-seeding as ViaPoint)."""
    total = float(n_links)
    rng = _rng(seed)
    if random_start:
        sample_first_joint(rng)     # first seeded reset already consumed one variate
    if target is None:
        goal = np.array([total, total])
        while np.linalg.norm(goal) >= total:
            goal = rng.uniform(low=-total, high=total, size=2)
    else:
        goal = np.array(target, dtype=np.float64)
    # random_start=False: simple_reacher.py:29 sets _start_pos = zeros
    q0 = sample_first_joint(_rng(seed)) if random_start else 0.0
    return dict(goal=goal, q0=float(q0))


# ----------------------------------------------------------------------------------------------
# geometry shared by all three envs
# ----------------------------------------------------------------------------------------------
def forward_kinematics(q):
    """base_reacher/base_reacher.py:95-103 (_update_joints); link lengths are all 1 (:19).
    q [B,n] -> joints [B,n+1,2]."""
    angles = np.cumsum(q, axis=1)
    B, n = q.shape
    J = np.zeros((B, n + 1, 2))
    J[:, 1:, 0] = np.cumsum(1.0 * np.cos(angles), axis=1)
    J[:, 1:, 1] = np.cumsum(1.0 * np.sin(angles), axis=1)
    return J


def _cross(A, B, C):
    # utils.py:1-2 (ccw) without the threshold:  (C_y-A_y)(B_x-A_x) - (B_y-A_y)(C_x-A_x)
    return (C[:, 1] - A[:, 1]) * (B[:, 0] - A[:, 0]) - (B[:, 1] - A[:, 1]) * (C[:, 0] - A[:, 0])


def self_collision(q, J, allow_self_collision=False, margins=True):
    """base_reacher/base_reacher.py:105-119 + utils.py:1-9.
    Returns (collided [B] bool, margin [B]).

    margin: how far the decision is from flipping, in the error model of a float32 evaluation.
    Every orientation test value is (in exact arithmetic) a sum of sines of relative link angles,
      ccw(J_i,J_i+1,J_m) = sum_{l=i+1}^{m-1} sin(th_l - th_i),  ccw(J_a,J_j,J_j+1) = sum_{l=a}^{j-1} sin(th_j - th_l),
    so a float32 evaluation has an error *relative* to sum |sin|; the margin of one test is
    |value - 1e-12| / sum|sin| (inf when all its terms vanish: exactly collinear links are decided
    identically by both sides).  Per link pair the margin is what it takes to flip `intersect`
    (both XORs must hold: max of the two "needs" when it is False, min of the two "breaks" when True);
    the joint-limit margin |q| - pi is absolute (q itself is bit-identical on both sides)."""
    B, n = q.shape
    if allow_self_collision:
        return np.zeros(B, bool), np.full(B, np.inf)
    limit = np.any(q > np.pi, axis=1) | np.any(q < -np.pi, axis=1)
    hit = np.zeros(B, bool)
    if not margins and B == 1:
        # one env, like the reference: plain Python float arithmetic on the joint coordinates (same IEEE operations as the
        # array expressions below, without numpy's per-call overhead on 1-element arrays) — used by bench.py's CPU arm
        P = J[0].tolist()

        def ccw(a, b, c):
            return (c[1] - a[1]) * (b[0] - a[0]) - (b[1] - a[1]) * (c[0] - a[0]) > CCW_EPS
        h = False
        for i in range(n):
            for j in range(i + 2, n):
                a, b_, c, d = P[i], P[i + 1], P[j], P[j + 1]
                if ccw(a, c, d) != ccw(b_, c, d) and ccw(a, b_, c) != ccw(a, b_, d):
                    h = True
        return limit | np.array([h]), np.full(B, np.inf)
    if not margins:     # decision only, exactly the reference's arithmetic
        for i in range(n):
            for j in range(i + 2, n):
                A, Bp, C, D = J[:, i], J[:, i + 1], J[:, j], J[:, j + 1]
                hit |= ((_cross(A, C, D) > CCW_EPS) != (_cross(Bp, C, D) > CCW_EPS)) & \
                       ((_cross(A, Bp, C) > CCW_EPS) != (_cross(A, Bp, D) > CCW_EPS))
        return limit | hit, np.full(B, np.inf)
    margin = np.min(np.abs(np.abs(q) - np.pi), axis=1)
    th = np.cumsum(q, axis=1)
    absin = np.abs(np.sin(th[:, None, :] - th[:, :, None]))       # [B, i, l] = |sin(th_l - th_i)|

    def rel(c, scale):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(scale > 0, np.abs(c - CCW_EPS) / scale, np.inf)

    for i in range(n):
        for j in range(i + 2, n):
            A, Bp, C, D = J[:, i], J[:, i + 1], J[:, j], J[:, j + 1]
            c1, c2, c3, c4 = _cross(A, C, D), _cross(Bp, C, D), _cross(A, Bp, C), _cross(A, Bp, D)
            x12 = (c1 > CCW_EPS) != (c2 > CCW_EPS)
            x34 = (c3 > CCW_EPS) != (c4 > CCW_EPS)
            pair_hit = x12 & x34
            hit |= pair_hit
            s3 = absin[:, i, i + 1:j].sum(axis=1)
            s4 = s3 + absin[:, i, j]
            s2 = absin[:, i + 1:j, j].sum(axis=1)
            s1 = s2 + absin[:, i, j]
            m12 = np.minimum(rel(c1, s1), rel(c2, s2))       # perturbation that flips XOR(c1, c2)
            m34 = np.minimum(rel(c3, s3), rel(c4, s4))
            need = np.maximum(np.where(x12, 0.0, m12), np.where(x34, 0.0, m34))
            brk = np.minimum(m12, m34)
            margin = np.minimum(margin, np.where(pair_hit, brk, need))
    return limit | hit, margin


_S_M = np.linspace(0, 1, N_LINE_POINTS)    # hole_reacher/hole_reacher.py:129


def line_points(q):
    """hole_reacher/hole_reacher.py:126-143 (_get_line_points, 100 points per link).
    q [B,n] -> pts [B,n,100,2]."""
    B, n = q.shape
    acc = np.cumsum(q, axis=1)[:, :, None]
    x = np.cos(acc) * 1.0 * _S_M
    y = np.sin(acc) * 1.0 * _S_M
    pts = np.zeros((B, n, N_LINE_POINTS, 2))
    pts[:, 0, :, 0] = x[:, 0]
    pts[:, 0, :, 1] = y[:, 0]
    for i in range(1, n):
        pts[:, i, :, 0] = x[:, i] + pts[:, i - 1, -1, 0][:, None]
        pts[:, i, :, 1] = y[:, i] + pts[:, i - 1, -1, 1][:, None]
    return pts


def wall_collision(q, hole_x, hole_w, hole_d, allow_wall_collision=False, margins=True):
    """hole_reacher/hole_reacher.py:148-179 (check_wall_collision).
    Returns (collided [B], margin [B]).  margin: for a free state the L-inf distance of the
    nearest sampled point to the forbidden region; for a colliding state the largest L-inf
    depth of a sampled point inside it."""
    B = q.shape[0]
    if allow_wall_collision:
        return np.zeros(B, bool), np.full(B, np.inf)
    pts = line_points(q).reshape(B, -1, 2)
    px, py = pts[:, :, 0], pts[:, :, 1]
    # (sample 0 of link 0 is the arm base, exactly (0, 0) on both sides: it sits on the region boundary by
    # construction and is excluded from the margin, not from the decision)
    xl = (hole_x - hole_w / 2)[:, None]
    xr = (hole_x + hole_w / 2)[:, None]
    d = hole_d[:, None]
    r1 = (px < xl) & (py < 0)
    r2 = (px > xr) & (py < 0)
    r3 = (px > xl) & (px < xr) & (py < -d)
    inside = r1 | r2 | r3
    hit = inside.any(axis=1)
    if not margins:
        return hit, np.full(B, np.inf)
    pos = lambda a: np.maximum(a, 0.0)   # noqa: E731
    dist1 = np.maximum(pos(px - xl), pos(py))
    dist2 = np.maximum(pos(xr - px), pos(py))
    dist3 = np.maximum(np.maximum(pos(xl - px), pos(px - xr)), pos(py + d))
    free_margin = np.minimum(np.minimum(dist1, dist2), dist3)[:, 1:].min(axis=1)
    depth1 = np.minimum(xl - px, -py)
    depth2 = np.minimum(px - xr, -py)
    depth3 = np.minimum(np.minimum(px - xl, xr - px), -d - py)
    depth = np.where(r1, depth1, np.where(r2, depth2, np.where(r3, depth3, 0.0)))
    hit_margin = depth.max(axis=1)
    return hit, np.where(hit, hit_margin, free_margin)


# ----------------------------------------------------------------------------------------------
# the envs
# ----------------------------------------------------------------------------------------------
class BatchedReacher:
    """B independent copies of one of HoleReacherEnv / ViaPointReacherEnv / SimpleReacherEnv.

    kind: 'hole' | 'viapoint' | 'simple'.  Constructor keyword names follow the reference
    constructors (hole_reacher.py:18-20, viapoint_reacher.py:15-16, simple_reacher.py:20-21).
    """

    def __init__(self, kind, n_links, random_start=None, allow_self_collision=False,
                 allow_wall_collision=False, collision_penalty=None, hole_x=None, hole_depth=None,
                 hole_width=1.0, via_target=None, target=None, rew_fct="simple"):
        assert kind in ("hole", "viapoint", "simple")
        self.kind = kind
        self.n_links = n_links
        if random_start is None:   # class defaults: hole False (:19), viapoint False (:15), simple True (:20)
            random_start = kind == "simple"
        self.random_start = random_start
        self.allow_self_collision = allow_self_collision
        self.allow_wall_collision = allow_wall_collision
        if collision_penalty is None:
            collision_penalty = 1000
        self.collision_penalty = collision_penalty
        self.initial_x, self.initial_depth, self.initial_width = hole_x, hole_depth, hole_width
        self.initial_via_target, self.initial_target = via_target, target
        if rew_fct not in ("simple", "vel_acc", "unbounded"):
            raise ValueError("Unknown reward function {}".format(rew_fct))        # hole_reacher.py:57-58
        self.rew_fct = rew_fct
        self.dt = DT
        self.compute_margins = True     # bench.py's CPU baseline switches the (non-reference) margin bookkeeping off
        self.double_collision_eval = False   # ... and evaluates the collision tests twice per step like the reference (App. A.6-Q3)
        self.torque = kind == "simple"
        # action bounds: base_reacher_direct.py:16-18 (2*pi), base_reacher_torque.py:16-18 (1000);
        # Box default dtype float32
        bound = 1000.0 if self.torque else 2 * np.pi
        self.action_low = (-np.ones(n_links) * bound).astype(np.float32)
        self.action_high = (np.ones(n_links) * bound).astype(np.float32)
        self.obs_dim = 3 * n_links + {"hole": 4, "viapoint": 5, "simple": 3}[kind]

    # -- reset ---------------------------------------------------------------------------------
    def sample_contexts(self, seeds):
        ctxs = []
        for s in seeds:
            s = int(s)
            if self.kind == "hole":
                ctxs.append(sample_hole_context(s, self.initial_width, self.initial_x,
                                                self.initial_depth, self.random_start))
            elif self.kind == "viapoint":
                ctxs.append(sample_viapoint_context(s, self.n_links, self.initial_via_target,
                                                    self.initial_target, self.random_start))
            else:
                ctxs.append(sample_simple_context(s, self.n_links, self.initial_target,
                                                  self.random_start))
        return ctxs

    def reset(self, seeds=None, contexts=None):
        """contexts: list of dicts as returned by sample_*_context (explicit) or seeds."""
        if contexts is None:
            contexts = self.sample_contexts(seeds)
        B, n = len(contexts), self.n_links
        self.B = B
        self.q = np.zeros((B, n))
        self.q[:, 0] = [c["q0"] for c in contexts]
        self.v = np.zeros((B, n))                 # _start_vel: float64 zeros (base_reacher.py:35)
        self.acc = None
        self.steps = np.zeros(B, dtype=np.int64)
        if self.kind == "hole":
            self.hole_x = np.array([c["x"] for c in contexts], dtype=np.float64)
            self.hole_w = np.array([c["width"] for c in contexts], dtype=np.float64)
            self.hole_d = np.array([c["depth"] for c in contexts], dtype=np.float64)
            self.goal = np.stack([self.hole_x, -self.hole_d], axis=1)   # hole_reacher.py:100
            self.ee_latch = np.zeros((B, 2))                            # hr_unbounded_reward.py:35-36 (end_eff_pos)
        else:
            self.goal = np.stack([c["goal"] for c in contexts]).astype(np.float64)
            if self.kind == "viapoint":
                self.via = np.stack([c["via"] for c in contexts]).astype(np.float64)
        self.J = forward_kinematics(self.q)
        return self.get_obs()

    # -- properties used by the black-box loop (raw_interface_wrapper.py:24-53) ------------------
    @property
    def current_pos(self):
        return self.q.copy()

    @property
    def current_vel(self):
        return self.v.copy()

    @property
    def end_effector(self):
        return self.J[:, self.n_links]

    def context_mask(self):
        """*/mp_wrapper.py context_mask for the three envs."""
        n, rs = self.n_links, self.random_start
        if self.kind == "hole":
            m = [rs] * n + [rs] * n + [rs] * n + [self.initial_width is None] + [True] * 2 + [False]
        elif self.kind == "viapoint":
            m = [rs] * n + [rs] * n + [rs] * n + [self.initial_via_target is None] * 2 + [True] * 2 + [False]
        else:
            m = [rs] * n + [rs] * n + [rs] * n + [True] * 2 + [False]
        return np.array(m, dtype=bool)

    # -- observation ---------------------------------------------------------------------------
    def get_obs(self):
        """hole_reacher.py:114-124 / viapoint_reacher.py:112-121 / simple_reacher.py:75-83."""
        q, ee = self.q, self.end_effector
        parts = [np.cos(q), np.sin(q), self.v]
        if self.kind == "hole":
            parts += [self.hole_w[:, None]]
        elif self.kind == "viapoint":
            parts += [ee - self.via]
        parts += [ee - self.goal, self.steps[:, None]]
        return np.hstack(parts).astype(np.float32)

    # -- step ----------------------------------------------------------------------------------
    def step(self, action):
        """base_reacher_direct.py:20-38 / base_reacher_torque.py:20-37.
        action [B,n].  Returns obs [B,O] f32, reward [B] f64, terminated [B] bool, info dict of
        arrays (+ 'margin')."""
        if self.torque:
            self.v = self.v + self.dt * action
            self.q = self.q + self.dt * self.v
        else:
            self.acc = (action - self.v) / self.dt
            self.v = action
            self.q = self.q + self.dt * self.v
        self.J = forward_kinematics(self.q)

        mg = self.compute_margins
        selfc, margin = self_collision(self.q, self.J, self.allow_self_collision, mg)
        if self.kind != "simple" and self.double_collision_eval:   # the reference evaluates the tests again in the reward (Q3)
            self_collision(self.q, self.J, self.allow_self_collision, mg)
        if self.kind == "hole":
            wallc, wmargin = wall_collision(self.q, self.hole_x, self.hole_w, self.hole_d,
                                            self.allow_wall_collision, mg)
            if self.double_collision_eval:
                wall_collision(self.q, self.hole_x, self.hole_w, self.hole_d, self.allow_wall_collision, mg)
            collided = selfc | wallc
            margin = np.minimum(margin, wmargin)
            reward, info = {"simple": self._reward_hole, "vel_acc": self._reward_hole_vel_acc,
                            "unbounded": self._reward_hole_unbounded}[self.rew_fct](collided)
        elif self.kind == "viapoint":
            collided = selfc
            reward, info = self._reward_viapoint(action, collided)
        else:
            collided = selfc
            reward, info = self._reward_simple(action)
        info["margin"] = margin
        self.steps = self.steps + 1
        terminated = np.zeros(self.B, bool) if self.torque else collided.copy()
        return self.get_obs(), reward, terminated, info

    # hole_reacher/hr_simple_reward.py:19-53
    def _reward_hole(self, collided):
        ee = self.end_effector
        last = (self.steps == 199) | collided
        diff = ee - self.goal
        dist = np.sqrt(diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1])
        dist_cost = np.where(last, dist ** 2, 0.0)
        collision_cost = np.where(last, collided.astype(np.float64), 0.0)
        success = last & (dist < 0.005) & ~collided
        acc_cost = np.sum(self.acc ** 2, axis=1)
        factors = np.array((-1, -5e-8, -self.collision_penalty), dtype=np.float64)
        reward = dist_cost * factors[0] + acc_cost.astype(np.float64) * factors[1] + collision_cost * factors[2]
        return reward, dict(is_success=success, is_collided=collided.copy(), end_effector=ee.copy())

    def _goal_dist_rowwise(self, ee):
        """np.linalg.norm(ee - goal) per env, with the 1-D call the reference makes (BLAS nrm2 and the batched
        sqrt(sum(x*x)) may differ in the last bit, which exp() would amplify)"""
        return np.array([np.linalg.norm(ee[b] - self.goal[b]) for b in range(self.B)])

    # hole_reacher/hr_dist_vel_acc_reward.py:20-60.  The reward object latches the first collision, but a collision
    # also terminates the episode (hole_reacher.py:76-77), so inside an episode the latch equals this step's test and
    # collision_dist is this step's distance.  Distance / collision terms exist on step 199 only.
    def _reward_hole_vel_acc(self, collided):
        ee = self.end_effector
        last = self.steps == 199
        dist = self._goal_dist_rowwise(ee)
        dist_cost = np.where(last, dist ** 2, 0.0)
        collision_cost = np.where(last, collided * dist ** 2, 0.0)
        success = last & (dist < 0.005) & ~collided
        vel_cost = np.sum(self.v ** 2, axis=1).astype(np.float64)
        acc_cost = np.sum(self.acc ** 2, axis=1).astype(np.float64)
        f = np.array((-1, -1e-4, -1e-6, -self.collision_penalty, 0), dtype=np.float64)
        reward = dist_cost * f[0] + vel_cost * f[1] + acc_cost * f[2] + collision_cost * f[3]
        return reward, dict(is_success=success, is_collided=collided.copy(), end_effector=ee.copy())

    # hole_reacher/hr_unbounded_reward.py:17-60
    def _reward_hole_unbounded(self, collided):
        ee = self.end_effector
        latch = (self.steps == 180) | collided
        self.ee_latch = np.where(latch[:, None], ee, self.ee_latch)
        last = (self.steps == 199) | collided
        dist = self._goal_dist_rowwise(self.ee_latch)
        dist_reward = np.where(collided, 0.25 * np.exp(-dist), np.where(ee[:, 1] > 0, np.exp(-dist), 1 - self.ee_latch[:, 1]))
        dist_reward = np.where(last, dist_reward, 0.0)
        success = last & ~collided
        acc_cost = np.sum(self.acc ** 2, axis=1).astype(np.float64)
        reward = dist_reward * 1.0 + acc_cost * -5e-6
        return reward, dict(is_success=success, is_collided=collided.copy(), end_effector=ee.copy(), joints=self.q.copy())

    # viapoint_reacher/viapoint_reacher.py:79-107  (App. A.6-Q1: starts from -inf; Q2: `acc` is the action)
    def _reward_viapoint(self, action, collided):
        ee = self.end_effector
        dvia = np.linalg.norm(ee - self.via, axis=1)
        dgoal = np.linalg.norm(ee - self.goal, axis=1)
        dist = np.full(self.B, np.inf)
        dist = np.where(self.steps == 100, dvia, dist)
        dist = np.where(self.steps == 199, dgoal, dist)
        success = (dist < 0.005) & ~collided
        dist = np.where(collided, dgoal, dist)
        reward = np.where(collided, -float(self.collision_penalty), -np.inf)
        with np.errstate(invalid="ignore"):
            reward = reward - dist ** 2
            reward = reward - 5e-8 * np.sum(action ** 2, axis=1)
        return reward, dict(is_success=success, is_collided=collided.copy(), end_effector=ee.copy())

    # simple_reacher/simple_reacher.py:56-70
    def _reward_simple(self, action):
        diff = self.end_effector - self.goal
        reward_dist = np.where(self.steps >= 199, -np.linalg.norm(diff, axis=1), 0.0)
        reward_ctrl = (action ** 2).sum(axis=1)
        reward = reward_dist - reward_ctrl
        return reward, dict(reward_dist=reward_dist, reward_ctrl=reward_ctrl)
