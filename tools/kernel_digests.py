"""GPU: SHA-256 fingerprints of everything the fused rollout returns, over a matrix of configurations.

    python tools/kernel_digests.py write tests/golden/rollout_digests.json     # record (run with the kernel to be trusted)
    python tools/kernel_digests.py check tests/golden/rollout_digests.json     # compare the current build against the record

A kernel rewrite that is meant to change scheduling only (re-packing of live envs, plans looped inside one launch, ...)
must reproduce every digest bit for bit; tests/test_gpu_digests.py runs the check.  The digests depend on the CUDA toolkit the
library is built with (device libm), which the record names.
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.pop("FG_PHASE_F32", None)      # the record was taken with the float64 per-env basis (the default)

import numpy as np  # noqa: E402
import torch  # noqa: E402

_R25 = dict(replanning_schedule=lambda p, v, o, a, t: t % 25 == 0, max_planning_times=4)
IDS = [f"fancy_{mp}/{name}" for name in ("HoleReacher-v0", "ViaPointReacher-v0", "SimpleReacher-v0", "LongSimpleReacher-v0")
       for mp in ("ProMP", "DMP", "ProDMP")]
# name, env id, B, sigma, black-box kwargs, env kwargs, mp_config_override sections, number of step() calls, tau column
CASES = [(f"{i}-s{s}", i, 4133, s, {}, {}, {}, 1, None) for i in IDS for s in (0.25, 1.0)]
CASES += [
    ("hole-vel_acc", "fancy_ProMP/HoleReacher-v0", 4133, 0.5, {}, dict(rew_fct="vel_acc"), {}, 1, None),
    ("hole-unbounded", "fancy_ProMP/HoleReacher-v0", 4133, 0.5, {}, dict(rew_fct="unbounded"), {}, 1, None),
    ("hole-allow-self", "fancy_ProMP/HoleReacher-v0", 4133, 1.0, {}, dict(allow_self_collision=True), {}, 1, None),
    ("hole-allow-wall", "fancy_ProMP/HoleReacher-v0", 4133, 1.0, {}, dict(allow_wall_collision=True, hole_x=1.0, hole_width=0.4, hole_depth=0.7), {}, 1, None),
    ("hole-wall-mode1", "fancy_ProMP/HoleReacher-v0", 4133, 1.0, dict(wall_mode=1), {}, {}, 1, None),
    ("viapoint-free-random-start", "fancy_DMP/ViaPointReacher-v0", 4133, 1.0, {}, dict(allow_self_collision=True, random_start=True), {}, 1, None),
    ("simple-prodmp-replan25", "fancy_ProDMP/SimpleReacher-v0", 4133, 1.0, dict(_R25, condition_on_desired=False), {}, {}, 4, None),
    ("simple-prodmp-replan25-cod", "fancy_ProDMP/SimpleReacher-v0", 4133, 1.0, dict(_R25, condition_on_desired=True), {}, {}, 4, None),
    ("hole-promp-replan50", "fancy_ProMP/HoleReacher-v0", 4133, 0.5, dict(replanning_schedule=lambda p, v, o, a, t: t % 50 == 0), {}, {}, 4, None),
    ("hole-dmp-replan40-cod", "fancy_DMP/HoleReacher-v0", 4133, 0.3, dict(replanning_schedule=lambda p, v, o, a, t: t % 40 == 0, condition_on_desired=True), {}, {}, 5, None),
    ("hole-prodmp-unbounded-replan60-cod", "fancy_ProDMP/HoleReacher-v0", 4133, 0.5,
     dict(replanning_schedule=lambda p, v, o, a, t: t % 60 == 0, condition_on_desired=True), dict(rew_fct="unbounded"), {}, 4, None),
    ("simple-prodmp-subtraj-ragged", "fancy_ProDMP/SimpleReacher-v0", 2051, 0.4, dict(learn_sub_trajectories=True), {}, {}, 5, "ragged"),
    ("hole-promp-subtraj-ragged", "fancy_ProMP/HoleReacher-v0", 2051, 0.4, dict(learn_sub_trajectories=True), {}, {}, 5, "ragged"),
    ("viapoint-dmp-subtraj-shared", "fancy_DMP/ViaPointReacher-v0", 2051, 0.4, dict(learn_sub_trajectories=True), {}, {}, 4, "shared"),
    ("hole-promp-learn-tau-delay", "fancy_ProMP/HoleReacher-v0", 2051, 0.4, {}, {}, {"phase_generator_kwargs": dict(phase_generator_type="linear", learn_tau=True, learn_delay=True)}, 1, "tau-delay"),
    ("hole-promp-position", "fancy_ProMP/HoleReacher-v0", 4133, 0.5, {}, {}, {"controller_kwargs": dict(controller_type="position")}, 1, None),
    ("hole-promp-k7", "fancy_ProMP/HoleReacher-v0", 4133, 0.5, {}, {}, {"basis_generator_kwargs": dict(basis_generator_type="zero_rbf", num_basis=7, num_basis_zero_start=1, basis_bandwidth_factor=3.0)}, 1, None),
    ("simple-promp-3links-k6", "fancy_ProMP/SimpleReacher-v0", 4133, 0.5, {}, dict(n_links=3), {"basis_generator_kwargs": dict(basis_generator_type="zero_rbf", num_basis=6, num_basis_zero_start=1, basis_bandwidth_factor=3.0)}, 1, None),
    ("hole-config2", "fancy_ProMP/HoleReacher-v0", 65536, 0.25, {}, {}, {}, 1, None),
    ("hole-config2-sigma1", "fancy_ProMP/HoleReacher-v0", 65536, 1.0, {}, {}, {}, 1, None),
    ("viapoint-config3", "fancy_DMP/ViaPointReacher-v0", 262144, 1.0, {}, {}, {}, 1, None),
]


def digest_case(fancy_gym, case, dev):
    name, env_id, B, sigma, bbk, env_kw, over, n_calls, lead = case
    override = dict(over)
    if bbk:
        override["black_box_kwargs"] = dict(bbk)
    env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override=override, **env_kw)
    env.reset(seed=7)
    gen = torch.Generator(device=dev).manual_seed(3)
    P = env.action_space.shape[0]
    h = hashlib.sha256()
    steps = 0
    for call in range(n_calls):
        params = sigma * torch.randn(B, P, generator=gen, device=dev)
        if lead == "ragged":
            params[:, 0] = 0.05 + 0.85 * torch.rand(B, generator=gen, device=dev)
        elif lead == "shared":
            params[:, 0] = (0.37, 0.8, 0.55, 2.0)[call]
        elif lead == "tau-delay":
            params[:, 0] = 0.3 + 2.0 * torch.rand(B, generator=gen, device=dev)
            params[:, 1] = 0.5 * torch.rand(B, generator=gen, device=dev)
        obs, ret, te, tr, info = env.step(params)
        torch.cuda.synchronize()
        for x in (obs, ret, te, tr, info["trajectory_length"]):
            h.update(x.cpu().numpy().tobytes())
        for k in sorted(info):
            if k != "trajectory_length":
                h.update(info[k].cpu().numpy().tobytes())
        steps += int(info["trajectory_length"].sum())
    env.close()
    return h.hexdigest(), steps


def main():
    import fancy_gym_b200 as fancy_gym
    mode, path = sys.argv[1], sys.argv[2]
    only = sys.argv[3] if len(sys.argv) > 3 else None
    dev = torch.device("cuda", 0)
    out = {}
    for case in CASES:
        if only and only not in case[0]:
            continue
        d, steps = digest_case(fancy_gym, case, dev)
        out[case[0]] = dict(sha256=d, env_steps=steps)
        print(case[0], d[:16], steps, flush=True)
    if mode == "write":
        with open(path, "w") as f:
            json.dump(dict(cuda=torch.version.cuda, note="outputs of the fused rollout (obs, return, flags, lengths, infos) per case; "
                           "recorded with the round-1 kernel (one thread per env, no re-packing, one launch per plan)", cases=out), f, indent=1)
    else:
        with open(path) as f:
            want = json.load(f)["cases"]
        bad = [k for k in out if want[k]["sha256"] != out[k]["sha256"]]
        print("MISMATCH" if bad else "all digests match", bad)
        sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
