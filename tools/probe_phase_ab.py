"""GPU: bit-level fingerprint of the per-env-phase trajectories (run once per kernel variant: FG_PHASE_BLOCK=0/1) + timing."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fancy_gym_b200 as fancy_gym
dev = torch.device("cuda", 0)
for env_id, phase, B in (("fancy_ProMP/HoleReacher-v0", dict(phase_generator_type="linear", learn_tau=True, learn_delay=True), 4099),
                         ("fancy_DMP/ViaPointReacher-v0", dict(phase_generator_type="exp", alpha_phase=2, learn_tau=True), 4099),
                         ("fancy_ProDMP/HoleReacher-v0", dict(learn_tau=True, learn_delay=True), 4099),
                         ("fancy_ProMP/HoleReacher-v0", "ragged", 4099), ("fancy_DMP/ViaPointReacher-v0", "ragged", 4099)):
    if phase == "ragged":
        env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override={"black_box_kwargs": {"learn_sub_trajectories": True}})
    else:
        env = fancy_gym.make(env_id, num_envs=B, device=dev, mp_config_override={"phase_generator_kwargs": phase})
    env.reset(seed=1)
    g = torch.Generator(device=dev).manual_seed(5)
    P = env.action_space.shape[0]
    p = 0.4 * torch.randn(B, P, generator=g, device=dev)
    p[:, 0] = 0.05 + 1.9 * torch.rand(B, generator=g, device=dev)
    if phase != "ragged" and phase.get("learn_delay"):
        p[:, 1] = 0.5 * torch.rand(B, generator=g, device=dev)
    pos, vel = env.get_trajectory(p)
    torch.cuda.synchronize()
    h = hashlib.sha256(pos.cpu().numpy().tobytes() + vel.cpu().numpy().tobytes()).hexdigest()[:16]
    print(env_id, "ragged" if phase == "ragged" else "phase", tuple(pos.shape), h, flush=True)
