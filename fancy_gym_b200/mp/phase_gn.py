"""Phase generators (host side): the role of mp_pytorch.phase_gn.{Linear,ExpDecay}PhaseGenerator
behind fancy_gym/black_box/factory/phase_generator_factory.py:9-23.

The phase itself is evaluated while the basis tables for the CUDA kernels are built
(fancy_gym_b200/mp/basis_gn.py); this class carries tau / delay, their learnable-parameter
bookkeeping (params are consumed from the front of the vector: tau first, delay second,
test/test_black_box.py:168-193) and the "finalize" rule (tau/delay only change on the first
set_params after reset(), test/test_replanning_sequencing.py:285-335).
"""
from __future__ import annotations

import numpy as np
import torch

from .assumptions import ASSUMPTIONS


class PhaseGenerator:
    kind = "linear"

    def __init__(self, tau: float = 3.0, delay: float = 0.0, learn_tau: bool = False, learn_delay: bool = False,
                 tau_bound=None, delay_bound=None, **kwargs):
        self.assume = dict(ASSUMPTIONS)      # the readings of mp_pytorch this generator is built under (mp/assumptions.py)
        self._tau0 = float(tau)
        self._delay0 = float(delay)
        self.learn_tau = bool(learn_tau)
        self.learn_delay = bool(learn_delay)
        self.tau_bound = list(tau_bound) if tau_bound is not None else [1e-5, float("inf")]
        self.delay_bound = list(delay_bound) if delay_bound is not None else [0.0, float("inf")]
        self.reset()

    # -- state --------------------------------------------------------------------------------
    def reset(self):
        """un-finalize: the next set_params() may change tau / delay again"""
        self.tau = torch.tensor(self._tau0, dtype=torch.float32)
        self.delay = torch.tensor(self._delay0, dtype=torch.float32)
        self.is_finalized = False

    def finalize(self):
        self.is_finalized = True

    @property
    def num_params(self) -> int:
        return int(self.learn_tau) + int(self.learn_delay)

    def set_params(self, params: torch.Tensor) -> torch.Tensor:
        """Consumes [tau][delay] from the front of params[..., P]; returns the remaining columns.  For a batch the values
        stay on the params' device as [B] float32 tensors (per-env phase, no host synchronisation); the caller has
        already clipped them to tau_bound / delay_bound (black_box_wrapper.py:104-105)."""
        i = 0
        for flag, name in ((self.learn_tau, "tau"), (self.learn_delay, "delay")):
            if not flag:
                continue
            if not self.is_finalized:
                v = params[..., i].detach().to(torch.float32)
                if v.numel() <= 1:          # single env: a host scalar, shared tables
                    v = v.reshape(()).cpu()
                    if name == "tau" and not float(v) > 0:
                        raise AssertionError("tau must be positive")
                    if name == "delay" and not float(v) >= 0:
                        raise AssertionError("delay must be non-negative")
                setattr(self, name, v)
            i += 1
        self.finalize()
        return params[..., i:]

    def get_params_bounds(self):
        lo, hi = [], []
        if self.learn_tau:
            lo.append(self.tau_bound[0]); hi.append(self.tau_bound[1])
        if self.learn_delay:
            lo.append(self.delay_bound[0]); hi.append(self.delay_bound[1])
        return lo, hi

    # -- numerics (float32 linear phase exactly as the library's elementwise torch ops) ---------
    def uniform(self) -> bool:
        """True when tau and delay are scalars shared by every env of the batch (shared tables); per-env values ([B]
        tensors) are handled by the per-env-phase kernels without looking at them on the host."""
        return self.tau.dim() == 0 and self.delay.dim() == 0

    def scalar_tau(self) -> float:
        """tau for the shared tables; with a per-env phase the construction-time value (tables are nominal then)"""
        return float(self.tau) if self.tau.dim() == 0 else self._tau0

    def scalar_delay(self) -> float:
        return float(self.delay) if self.delay.dim() == 0 else self._delay0

    def collapse_if_equal(self) -> bool:
        """If every env of the batch carries the same tau and the same delay, store them as shared scalars (table-driven
        kernels).  Reads the values back to the host: used where the host needs them anyway (sub-trajectory lengths)
        or where no per-env kernel exists (ProDMP).  Returns uniform()."""
        for name in ("tau", "delay"):
            x = getattr(self, name)
            if x.dim() > 0:
                xc = x.detach().cpu().reshape(-1)
                if bool((xc == xc[0]).all()):
                    setattr(self, name, xc[0].clone())
        return self.uniform()

    def per_env(self, num_envs: int, device):
        """(tau [B], delay [B]) float32 on the device"""
        def expand(x):
            x = x.to(device, torch.float32)
            return x.expand(num_envs).contiguous() if x.dim() == 0 else x.contiguous()
        return expand(self.tau), expand(self.delay)

    def linear_phase32(self, times32: np.ndarray, clip_hi: bool = True) -> np.ndarray:
        tau, delay = np.float32(self.scalar_tau()), np.float32(self.scalar_delay())
        z = (times32.astype(np.float32) - delay) / tau
        return np.clip(z, 0, 1 if clip_hi else None).astype(np.float32)

    def phase_argument32(self, times32: np.ndarray) -> np.ndarray:
        """the scaled time the canonical phase is a function of: the linear phase clipped to [0, 1]"""
        return self.linear_phase32(times32)

    def phase64(self, lin) -> np.ndarray:
        """canonical phase in float64 from a linear phase"""
        return np.asarray(lin, dtype=np.float64)

    def unbound_phase64_of_time(self, t64):
        """canonical phase (unbounded) of absolute times with the construction-time tau / delay"""
        return (np.asarray(t64, dtype=np.float64) - self._delay0) / self._tau0

    def centre_phase64_of_time(self, t64):
        """where an RBF centre placed at time t sits in phase space (switch centres_through_unbounded_phase)"""
        if self.assume["centres_through_unbounded_phase"]:
            return self.unbound_phase64_of_time(t64)
        z = np.clip((np.asarray(t64, dtype=np.float64) - self._delay0) / self._tau0, 0, 1)
        return self.phase64(z)


class LinearPhaseGenerator(PhaseGenerator):
    kind = "linear"


class ExpDecayPhaseGenerator(PhaseGenerator):
    kind = "exp"

    def __init__(self, tau: float = 3.0, delay: float = 0.0, alpha_phase: float = None, learn_tau: bool = False,
                 learn_delay: bool = False, learn_alpha_phase: bool = False, **kwargs):
        if learn_alpha_phase:
            raise NotImplementedError("learn_alpha_phase is not used by any fancy_gym config")
        super().__init__(tau, delay, learn_tau, learn_delay, **kwargs)
        self.alpha_phase = float(self.assume["alpha_phase_default"] if alpha_phase is None else alpha_phase)

    def phase_argument32(self, times32: np.ndarray) -> np.ndarray:
        """switch exp_phase_right_clip: x = exp(-alpha z) of the clipped linear phase, or of the left-bounded one"""
        return self.linear_phase32(times32, clip_hi=bool(self.assume["exp_phase_right_clip"]))

    def phase64(self, lin) -> np.ndarray:
        return np.exp(-self.alpha_phase * np.asarray(lin, dtype=np.float64))

    def unbound_phase64_of_time(self, t64):
        return np.exp(-self.alpha_phase * super().unbound_phase64_of_time(t64))
