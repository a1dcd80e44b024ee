"""Named switches for the points where mp_pytorch's behaviour could not be checked against the package itself.

The reference delegates the movement-primitive arithmetic to `mp_pytorch<=0.1.3` (pyproject.toml:30; call sites
fancy_gym/black_box/factory/{phase,basis,trajectory}_generator_factory.py).  That package is neither vendored in the
reference nor installed in this image, so every reading of it that is not fixed by the reference's own tests is kept behind a
switch (SURVEY.md App. B.8 i-vi, plus the right clip of the exponential phase and the ProDMP table lookup).  The DEFAULT is
the reading this implementation ships; `with assume(switch=value): env = fancy_gym.make(...)` builds an env under the other
reading.  A generator captures the switches when it is constructed.  oracle/mp.py carries the same table (a test keeps the
two in step); DESIGN.md §2 lists how much every BASELINE env moves when a switch flips (tools/mp_sensitivity.py).
"""
from __future__ import annotations

ASSUMPTIONS = {
    # B.8 (i)   DMP: the first grid point t0 + dt carries the initial state and the Euler recurrence starts there
    #           (set_duration(include_init_time=False)).  False: the recurrence starts at t0 and t0's row is dropped.
    "dmp_init_on_first_grid_point": True,
    # B.8 (ii)  weights_scale / goal_scale multiply the basis (ProMP, ProDMP) resp. the parameters (DMP).  False: the other
    #           way round (algebraically identical, float32 rounding differs).
    "scale_on_library_side": True,
    # B.8 (iii) alpha_phase of the exponential phase when a config gives none (registry.py:112-115: every fancy_ProDMP id)
    "alpha_phase_default": 3.0,
    # B.8 (iv)  RBF centres are mapped through the unbounded phase.  False: through the bounded one.
    "centres_through_unbounded_phase": True,
    # B.8 (v)   covariance regulariser reg * max(diag): per sample.  True: over the whole batch.
    "cov_reg_batch_global": False,
    # B.8 (vi)  goal_offset (kwarg, default 0, set by no classic_control config) is added after goal_scale.  False: before.
    "goal_offset_after_scale": True,
    # x = exp(-alpha_phase * z) with z the linear phase clipped to [0, 1].  False: z = max((t - delay) / tau, 0) only, x keeps
    # decaying after delay + tau (matters when tau < duration: fancy_ProDMP/* have tau = 1.5, duration 2.0).
    "exp_phase_right_clip": True,
    # ProDMP: the pre-integrated bases are read at the nearest grid index.  True: linear interpolation between grid points.
    "prodmp_interpolate": False,
}


class assume:
    """`with assume(exp_phase_right_clip=False): env = fancy_gym.make(...)`"""

    def __init__(self, **switches):
        unknown = set(switches) - set(ASSUMPTIONS)
        if unknown:
            raise KeyError(f"unknown assumption switch(es): {sorted(unknown)}")
        self.switches = switches

    def __enter__(self):
        self.saved = dict(ASSUMPTIONS)
        ASSUMPTIONS.update(self.switches)
        return self

    def __exit__(self, *exc):
        ASSUMPTIONS.clear()
        ASSUMPTIONS.update(self.saved)
        return False
