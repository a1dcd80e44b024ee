"""Basis generators (host side): the role of mp_pytorch.basis_gn.{NormalizedRBF,
ZeroPaddingNormalizedRBF, ProDMP}BasisGenerator behind
fancy_gym/black_box/factory/basis_generator_factory.py:8-23.

They build the small float32 tables the CUDA kernels stage in shared memory.  Transcendental
parts are evaluated in float64 and rounded once (SURVEY.md §7 "ProDMP numerics"); everything the
library computes with plain float32 elementwise ops (time grid, linear phase, time increments)
is computed with the same float32 ops here, so the tables carry the library's rounding pattern.
"""
from __future__ import annotations

import numpy as np

from .phase_gn import ExpDecayPhaseGenerator, PhaseGenerator


class NormalizedRBFBasisGenerator:
    def __init__(self, phase_generator: PhaseGenerator, num_basis: int = 10, basis_bandwidth_factor: float = 3,
                 num_basis_outside: int = 0, **kwargs):
        self.phase_generator = phase_generator
        self._num_basis = int(num_basis)
        self.basis_bandwidth_factor = basis_bandwidth_factor
        self.num_basis_outside = int(num_basis_outside)
        self.centers_p, self.bandwidth = self._centres()

    # learnable / total number of basis functions (zero padding makes them differ)
    @property
    def num_basis(self) -> int:
        return self._num_basis

    @property
    def total_num_basis(self) -> int:
        return self._num_basis

    @property
    def first_learnable(self) -> int:
        return 0

    def _centres(self):
        K = self.total_num_basis
        if K <= 1:
            return np.array([0.5]), np.array([3.0])
        pg = self.phase_generator
        tau, delay = pg._tau0, pg._delay0
        dist = tau / (K - 2 * self.num_basis_outside - 1)
        centres_t = np.linspace(-self.num_basis_outside * dist + delay, tau + self.num_basis_outside * dist + delay, K)
        cp = pg.centre_phase64_of_time(centres_t)
        spacing = np.concatenate([cp[1:] - cp[:-1], cp[-1:] - cp[-2:-1]])
        return cp, float(self.basis_bandwidth_factor) / spacing ** 2

    def basis64(self, lin_phase) -> np.ndarray:
        """normalised RBFs [..., K_total] (float64) at the given *linear* phase values"""
        ph = self.phase_generator.phase64(lin_phase)
        b = np.exp(-((ph[..., None] - self.centers_p) ** 2 * self.bandwidth) / 2)
        if self.total_num_basis > 1:
            b = b / b.sum(axis=-1, keepdims=True)
        return b

    def learnable_basis32(self, times32: np.ndarray) -> np.ndarray:
        lin = self.phase_generator.phase_argument32(times32)
        b = self.basis64(lin).astype(np.float32)
        z0 = self.first_learnable
        return np.ascontiguousarray(b[..., z0:z0 + self.num_basis])


class ZeroPaddingNormalizedRBFBasisGenerator(NormalizedRBFBasisGenerator):
    def __init__(self, phase_generator: PhaseGenerator, num_basis: int = 10, num_basis_zero_start: int = 2,
                 num_basis_zero_goal: int = 0, basis_bandwidth_factor: float = 3, **kwargs):
        self.num_basis_zero_start = int(num_basis_zero_start)
        self.num_basis_zero_goal = int(num_basis_zero_goal)
        self._learnable = int(num_basis)
        super().__init__(phase_generator, num_basis=int(num_basis) + self.num_basis_zero_start + self.num_basis_zero_goal,
                         basis_bandwidth_factor=basis_bandwidth_factor, num_basis_outside=0)

    @property
    def num_basis(self) -> int:
        return self._learnable

    @property
    def total_num_basis(self) -> int:
        return self._num_basis

    @property
    def first_learnable(self) -> int:
        return self.num_basis_zero_start


class ProDMPBasisGenerator(NormalizedRBFBasisGenerator):
    """Pre-integrated position / velocity bases of the ProDMP ODE solution on the scaled-time grid
    z_j = j * dt / tau, j = 0 .. factor * round(tau / dt), cumulative trapezoid."""

    def __init__(self, phase_generator: PhaseGenerator, num_basis: int = 10, basis_bandwidth_factor: float = 3,
                 num_basis_outside: int = 0, dt: float = 0.01, alpha: float = 25, pre_compute_length_factor: int = 6,
                 **kwargs):
        assert isinstance(phase_generator, ExpDecayPhaseGenerator)
        super().__init__(phase_generator, num_basis, basis_bandwidth_factor, num_basis_outside)
        self.alpha = float(alpha)
        self.dt = float(dt)
        self.pre_compute_length_factor = int(pre_compute_length_factor)
        self._pre_compute()

    @property
    def num_basis_g(self) -> int:
        return self._num_basis + 1

    def _pre_compute(self):
        pg = self.phase_generator
        tau = pg._tau0
        self.scaled_dt = self.dt / tau
        n_pc = self.pre_compute_length_factor * int(round(1.0 / self.scaled_dt)) + 1
        z = np.linspace(0, self.pre_compute_length_factor, n_pc)
        a = self.alpha
        y1 = np.exp(-0.5 * a * z)
        y2 = z * y1
        dy1 = -0.5 * a * y1
        dy2 = -0.5 * a * y2 + y1
        q1 = (0.5 * a * z - 1) * np.exp(0.5 * a * z) + 1
        q2 = 0.5 * a * (np.exp(0.5 * a * z) - 1)
        # grid times map back to linear phase z (clipped for x unless switch exp_phase_right_clip is off)
        lin = np.clip(z, 0, 1 if pg.assume["exp_phase_right_clip"] else None)
        x = pg.phase64(lin)
        b = self.basis64(lin)
        e = np.exp(a * z / 2)
        dp1 = (z * e * x)[:, None] * b
        dp2 = (e * x)[:, None] * b
        dz = np.diff(z)[:, None]
        zero = np.zeros((1, b.shape[1]))
        p1 = np.concatenate([zero, np.cumsum(0.5 * (dp1[1:] + dp1[:-1]) * dz, axis=0)])
        p2 = np.concatenate([zero, np.cumsum(0.5 * (dp2[1:] + dp2[:-1]) * dz, axis=0)])
        pos_w = p2 * y2[:, None] - p1 * y1[:, None]
        pos_g = q2 * y2 - q1 * y1
        vel_w = p2 * dy2[:, None] - p1 * dy1[:, None]
        vel_g = q2 * dy2 - q1 * dy1
        self.pc_pos_basis = np.concatenate([pos_w, pos_g[:, None]], axis=1)
        self.pc_vel_basis = np.concatenate([vel_w, vel_g[:, None]], axis=1)
        self.pc_y = np.stack([y1, y2, dy1, dy2], axis=1)
        self.auto_basis_scale_factors = 1.0 / np.abs(self.pc_pos_basis).max(axis=0)

    def indices(self, times) -> np.ndarray:
        # float32 like the library: torch.round(left_bound_linear_phase(times) / scaled_dt), scaled_dt = dt / tau0 a float32
        # tensor; ties (k * tau0 / tau = x.5, e.g. tau0 1.5 and a learned tau of 0.8) depend on that rounding
        pg = self.phase_generator
        f32 = np.float32
        z = np.maximum((np.asarray(times).astype(f32) - f32(pg.scalar_delay())) / f32(pg.scalar_tau()), f32(0)).astype(f32)
        if z.size and z.max() > self.pre_compute_length_factor:
            raise RuntimeError("Time is beyond the pre-computation range.")
        sd32 = f32(f32(self.dt) / f32(pg._tau0))
        return np.rint((z / sd32).astype(f32)).astype(np.int64)

    def lookup(self, table: np.ndarray, times) -> np.ndarray:
        """rows of a pre-computed table at the given times: nearest grid index (the default) or, switch prodmp_interpolate,
        linear interpolation in the index z / scaled_dt"""
        pg = self.phase_generator
        if not pg.assume["prodmp_interpolate"]:
            return table[self.indices(times)]
        f32 = np.float32      # the scaled time is a float32 tensor in the library; the interpolation weight is formed from it
        z = np.maximum((np.asarray(times, dtype=np.float64) - f32(pg.scalar_delay())) / f32(pg.scalar_tau()), 0.0)
        z = z.astype(f32).astype(np.float64)
        if z.size and z.max() > self.pre_compute_length_factor:
            raise RuntimeError("Time is beyond the pre-computation range.")
        idx = z / self.scaled_dt
        i0 = np.clip(np.floor(idx).astype(np.int64), 0, table.shape[0] - 2)
        fr = (idx - i0)[..., None] if table.ndim > 1 else (idx - i0)
        return table[i0] * (1 - fr) + table[i0 + 1] * fr
