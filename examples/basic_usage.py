"""The reference's black-box contract, scalar and batched (cf. fancy_gym/examples/examples_movement_primitives.py).

    python examples/basic_usage.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fancy_gym_b200 as fancy_gym  # noqa: E402


def scalar(env_id="fancy_ProMP/HoleReacher-v0", seed=1):
    """num_envs=1 with numpy actions: (obs[O], float, bool, bool, dict), exactly the reference's step()"""
    env = fancy_gym.make(env_id, num_envs=1, device="cuda:0")
    obs, _ = env.reset(seed=seed)
    action = env.action_space.sample()                       # the MP parameters of one whole trajectory
    obs, ret, terminated, truncated, info = env.step(action)
    print(f"{env_id}: return {ret:.4f}, {info['trajectory_length']} env steps, terminated={terminated}, "
          f"truncated={truncated}, collided={info.get('is_collided')}")
    env.close()


def batched(env_id="fancy_ProMP/HoleReacher-v0", num_envs=65536, seed=1):
    """one parameter vector per env, one fused launch for all episodes; everything stays on the device"""
    env = fancy_gym.make(env_id, num_envs=num_envs, device="cuda:0")
    obs, _ = env.reset(seed=seed)                            # env i samples its context like the reference's reset(seed + i)
    params = 0.25 * torch.randn(num_envs, env.action_space.shape[0], device="cuda:0")
    obs, ret, terminated, truncated, info = env.step(params)
    print(f"{env_id} x {num_envs}: mean return {ret.mean().item():.4f}, "
          f"{int(info['trajectory_length'].sum())} env steps, {int(terminated.sum())} terminated early")
    env.close()


def replanning(env_id="fancy_ProDMP/SimpleReacher-v0", num_envs=4096):
    """re-plan every 25 steps, conditioning each plan on the desired state of the previous one"""
    env = fancy_gym.make(env_id, num_envs=num_envs, device="cuda:0", mp_config_override={
        "black_box_kwargs": {"replanning_schedule": lambda pos, vel, obs, action, t: t % 25 == 0,
                             "condition_on_desired": True}})
    obs, _ = env.reset(seed=0)
    total = torch.zeros(num_envs, dtype=torch.float64, device="cuda:0")
    plans = 0
    while True:
        params = torch.randn(num_envs, env.action_space.shape[0], device="cuda:0")
        obs, ret, terminated, truncated, info = env.step(params)
        total += ret
        plans += 1
        if bool((terminated | truncated).all()):
            break
    print(f"{env_id} x {num_envs}: {plans} plans per episode, mean return {total.mean().item():.3f}")
    env.close()


if __name__ == "__main__":
    assert torch.cuda.is_available(), "fancy_gym_b200 runs on a CUDA device (there is no CPU fallback)"
    np.random.seed(0)
    scalar()
    batched()
    replanning()
