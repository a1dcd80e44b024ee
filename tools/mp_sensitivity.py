"""Sensitivity of every classic_control black-box id to every switchable reading of mp_pytorch (oracle/mp.py ASSUMPTIONS).

    python tools/mp_sensitivity.py [--all] [--envs 256]

For each env and each flipped switch: max |delta position|, max |delta velocity| of the planned trajectories, max |delta return|
over the episodes whose return is finite in both runs, and how many of the B episodes change length / termination — all in
the oracle's 'shipped' float32 mode (the reference's arithmetic), on B seeded episodes with sigma = 0.5 parameters.  The
table goes into DESIGN.md §2.  CPU only (test infrastructure; nothing here is product code).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import mp as omp  # noqa: E402
from oracle.blackbox import RESOLVED, make_oracle  # noqa: E402
from tests.golden.make_golden import n_params_of  # noqa: E402

FLIPS = [("dmp_init_on_first_grid_point", False), ("scale_on_library_side", False), ("alpha_phase_default", 2.0),
         ("centres_through_unbounded_phase", False), ("goal_offset_after_scale", False), ("exp_phase_right_clip", False),
         ("prodmp_interpolate", True)]
BASELINE_IDS = ["fancy_ProMP/HoleReacher-v0", "fancy_DMP/ViaPointReacher-v0", "fancy_ProDMP/SimpleReacher-v0"]


def run(env_id, B, flips, replan=False):
    kw = dict(replanning_schedule=lambda p, v, o, a, t: t % 25 == 0, max_planning_times=4) if replan else {}
    with omp.assume(**flips):
        orc = make_oracle(env_id, mode="shipped", **kw)
    orc.reset(seeds=np.arange(B))
    rng = np.random.default_rng(0)
    pos, vel, ret, length, term = [], [], np.zeros(B), np.zeros(B, np.int64), np.zeros(B, bool)
    for _ in range(4 if replan else 1):
        params = (0.5 * rng.standard_normal((B, n_params_of(env_id)))).astype(np.float32)
        p, v = orc.get_trajectory(params)
        o, r, te, tr, info = orc.step(params)
        pos.append(np.broadcast_to(p, (B, *p.shape[-2:])).copy())
        vel.append(np.broadcast_to(v, (B, *v.shape[-2:])).copy())
        ret += r
        length += info["trajectory_length"]
        term |= te
    return np.stack(pos), np.stack(vel), ret, length, term


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", action="store_true", help="all twelve ids instead of the three BASELINE ones")
    ap.add_argument("--envs", type=int, default=256)
    a = ap.parse_args()
    ids = list(RESOLVED) if a.all else BASELINE_IDS
    B = a.envs
    print(f"| env | switch flipped | max abs dpos | max abs dvel | max abs dreturn | episodes (of {B}) whose length / termination changes |")
    print("|---|---|---|---|---|---|")
    cases = [(i, False) for i in ids] + [("fancy_ProDMP/SimpleReacher-v0", True)]
    for env_id, replan in cases:
        base = run(env_id, B, {}, replan)
        for sw, val in FLIPS:
            alt = run(env_id, B, {sw: val}, replan)
            dpos = np.abs(alt[0] - base[0]).max()
            dvel = np.abs(alt[1] - base[1]).max()
            fin = np.isfinite(base[2]) & np.isfinite(alt[2])
            dret = f"{np.abs(alt[2] - base[2])[fin].max():.3g}" if fin.any() else "n/a (-inf)"
            changed = int(((alt[3] != base[3]) | (alt[4] != base[4])).sum())
            name = env_id + (" replanning t%25, 4 plans" if replan else "")
            print(f"| {name} | {sw} = {val} | {dpos:.3g} | {dvel:.3g} | {dret} | {changed} |")


if __name__ == "__main__":
    main()
