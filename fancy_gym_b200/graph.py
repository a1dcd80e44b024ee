"""CUDA-graph replay of a whole black-box episode for fixed-shape throughput loops (black-box optimisers evaluating one
population after another):   reset (next context of every env's stream)  ->  H2D of the parameters from pinned host memory
->  fused rollout  ->  D2H of returns / lengths / flags into pinned host memory   — captured once, replayed with one
cudaGraphLaunch per episode batch, so the per-step host cost is a single driver call.

    runner = GraphedEpisode(env)                 # env from fancy_gym_b200.make(...), already reset(seed=...) once
    runner.host_params[:] = population           # pinned [B, P] float32
    ret, length, terminated = runner.run()       # pinned host tensors, valid after the call returns

Only for envs that plan once per episode (no replanning / sub-trajectories: those change the plan between steps).
"""
from __future__ import annotations

import torch


class GraphedEpisode:
    def __init__(self, env, warmup: int = 3, copy_obs: bool = False):
        if env.do_replanning or env.learn_sub_trajectories:
            raise NotImplementedError("GraphedEpisode captures one plan per episode")
        if not env._fast_reset:
            raise NotImplementedError("GraphedEpisode needs the device-side reset (context_sampler='device')")
        self.env = env
        self.copy_obs = bool(copy_obs)      # also bring the context observation of every episode to the host
        dev = env.device
        B, P = env.num_envs, env.action_space.shape[0]
        self.host_params = torch.zeros(B, P, dtype=torch.float32).pin_memory()
        self.host_ret = torch.zeros(B, dtype=torch.float64).pin_memory()
        self.host_len = torch.zeros(B, dtype=torch.int32).pin_memory()
        self.host_terminated = torch.zeros(B, dtype=torch.bool).pin_memory()
        self.host_obs = torch.zeros(B, env.observation_space.shape[0], dtype=torch.float32).pin_memory()
        self._params = torch.zeros(B, P, dtype=torch.float32, device=dev)
        if env.unwrapped._rng_state is None:
            env.reset(seed=None)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):           # warm-up off the capture: handle creation, lazy module loading
            for _ in range(max(1, warmup)):
                self._episode()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # thread-local capture mode: CUDA calls of other threads (e.g. NCCL's watchdog) must not invalidate the capture
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._episode()

    def _episode(self):
        env = self.env
        obs0, _ = env.reset(seed=None, options={"as_numpy": False})
        self._params.copy_(self.host_params, non_blocking=True)
        _obs, ret, terminated, _trunc, info = env.step(self._params)
        if self.copy_obs:
            self.host_obs.copy_(obs0, non_blocking=True)              # the context observation the parameters answer to
        self.host_ret.copy_(ret, non_blocking=True)
        self.host_len.copy_(info["trajectory_length"], non_blocking=True)
        self.host_terminated.copy_(terminated, non_blocking=True)

    def run(self, sync: bool = True):
        """replays the captured episode on the current stream; with sync=True the pinned results are ready on return"""
        self.graph.replay()
        if sync:
            torch.cuda.current_stream(self.env.device).synchronize()
        return self.host_ret, self.host_len, self.host_terminated
