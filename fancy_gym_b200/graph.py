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


class EpisodePipeline:
    """Double-buffered episode batches for throughput loops that can keep two populations in flight (the
    step_async / step_wait pattern of vector envs): while the fused rollout of batch k runs, the parameters of batch
    k + 1 cross PCIe on a copy stream and the results of batch k - 1 return on another, so a step costs
    max(rollout, H2D, D2H) instead of their sum.

        pipe = EpisodePipeline(env)                       # env from fancy_gym_b200.make(...), reset(seed=...) once
        pipe.host_params[0][:] = population_0; pipe.submit(0)
        pipe.host_params[1][:] = population_1; pipe.submit(1)
        ret, length, terminated = pipe.wait(0)            # pinned host tensors of slot 0, valid until its next submit()
        pipe.host_params[0][:] = population_2; pipe.submit(0) ...

    Every submit() is reset (next context of every env's stream) -> H2D -> rollout -> D2H.  Every slot owns its env state
    (`env.new_state_set()`: q, v, steps, done, ctx) and its compute stream: the resets run in submission order on one stream
    (the envs' context streams advance exactly as with sequential reset() / step() calls), the ROLLOUTS of consecutive batches
    overlap — the second batch fills the SMs the sub-wave grid of the first leaves idle and covers its tail of long
    episodes.  Results are, batch for batch, those of sequential reset() / step() calls; the env's own state buffers are
    not touched.  Like GraphedEpisode: one plan per episode only."""

    SLOTS = 2      # == the wrapper's alternating result sets: the results of batch k stay valid while batch k + 1 runs

    def __init__(self, env, slots: int = 2, graphs: bool = False):
        """slots: batches in flight (2 by default).  More slots absorb copies that are as long as the rollout itself (8 ranks
        sharing the host's PCIe bandwidth: the H2D of one batch takes about as long as its rollout); the env needs as many
        result sets (`black_box_kwargs={'result_sets': slots}`).
        graphs: capture, per slot, the reset and the H2D -> rollout -> D2H chain as two CUDA graphs; a submit() then costs two
        graph launches and three event calls on the host instead of ~0.2 ms of Python (a 65 536-env rollout takes 0.24 ms:
        the eager pipeline is host bound once the rollouts of consecutive batches overlap)."""
        if env.do_replanning or env.learn_sub_trajectories:
            raise NotImplementedError("EpisodePipeline runs one plan per episode")
        if not env._fast_reset:
            raise NotImplementedError("EpisodePipeline needs the device-side reset (context_sampler='device')")
        self.SLOTS = int(slots)
        if self.SLOTS < 2:
            raise ValueError("EpisodePipeline needs at least 2 slots")
        if len(env._out_sets) < self.SLOTS:
            raise ValueError(f"{self.SLOTS} batches in flight need {self.SLOTS} result sets: make the env with "
                             f"mp_config_override={{'black_box_kwargs': {{'result_sets': {self.SLOTS}}}}} (it has {len(env._out_sets)})")
        self.env = env
        dev = env.device
        B, P = env.num_envs, env.action_space.shape[0]
        n = self.SLOTS
        self.host_params = [torch.zeros(B, P, dtype=torch.float32).pin_memory() for _ in range(n)]
        self.host_ret = [torch.zeros(B, dtype=torch.float64).pin_memory() for _ in range(n)]
        self.host_len = [torch.zeros(B, dtype=torch.int32).pin_memory() for _ in range(n)]
        self.host_terminated = [torch.zeros(B, dtype=torch.bool).pin_memory() for _ in range(n)]
        self._params = [torch.zeros(B, P, dtype=torch.float32, device=dev) for _ in range(n)]
        self._s_in, self._s_reset, self._s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self._s_run = [torch.cuda.Stream(device=dev) for _ in range(n)]      # one compute stream per slot: rollouts overlap
        self._state = [env.new_state_set() for _ in range(n)]
        self._ev_in = [torch.cuda.Event() for _ in range(n)]          # parameters of the slot are on the device
        self._ev_reset = [torch.cuda.Event() for _ in range(n)]       # the slot's env state is reset
        self._ev_run = [torch.cuda.Event() for _ in range(n)]         # rollout of the slot finished
        self._ev_out = [torch.cuda.Event() for _ in range(n)]         # results of the slot are on the host
        self._next = 0
        self._pending = [False] * n
        if env.unwrapped._rng_state is None:
            env.reset(seed=None)
        for s in (self._s_in, self._s_reset, self._s_out, *self._s_run):
            s.wait_stream(torch.cuda.current_stream(dev))
        self._g_reset, self._g_run = None, None
        if graphs:
            self._capture()

    def _capture(self):
        env, dev = self.env, self.env.device
        base = env.unwrapped
        rng = base._rng_state.clone()           # warm-up and capture must not advance the envs' context streams
        side = self._s_run[0]
        with torch.cuda.stream(side):           # off the capture: handle creation, lazy module loading
            env.reset_into(self._state[0])
            env.run_episode(self._params[0], self._state[0])
        side.synchronize()
        self._g_reset, self._g_run = [], []
        for slot in range(self.SLOTS):
            g = torch.cuda.CUDAGraph()
            # thread-local capture mode: CUDA calls of other threads (e.g. NCCL's watchdog) must not invalidate the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                env.reset_into(self._state[slot])
            self._g_reset.append(g)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._params[slot].copy_(self.host_params[slot], non_blocking=True)
                _obs, ret, terminated, _trunc, info = env.run_episode(self._params[slot], self._state[slot])   # its own result set
                self.host_ret[slot].copy_(ret, non_blocking=True)
                self.host_len[slot].copy_(info["trajectory_length"], non_blocking=True)
                self.host_terminated[slot].copy_(terminated, non_blocking=True)
            self._g_run.append(g)
        torch.cuda.synchronize(dev)
        base._rng_state.copy_(rng)
        torch.cuda.synchronize(dev)

    @property
    def next_slot(self) -> int:
        """the slot the next submit() must use (slots alternate)"""
        return self._next

    def submit(self, slot: int):
        """enqueues one episode batch with the parameters in host_params[slot]; returns immediately"""
        if slot != self._next:
            raise ValueError(f"slots are used in turn: expected {self._next}, got {slot}")
        if self._pending[slot]:
            raise RuntimeError(f"slot {slot} still holds results that were not collected with wait()")
        env = self.env
        if self._g_run is not None:
            self._s_reset.wait_event(self._ev_out[slot])   # the slot's previous batch is through (state, parameters, result set)
            with torch.cuda.stream(self._s_reset):         # resets in submission order: every env's context stream advances once
                self._g_reset[slot].replay()
                self._ev_reset[slot].record(self._s_reset)
            run = self._s_run[slot]
            run.wait_event(self._ev_reset[slot])
            with torch.cuda.stream(run):
                self._g_run[slot].replay()
                self._ev_out[slot].record(run)
            self._pending[slot] = True
            self._next = (slot + 1) % self.SLOTS
            return
        self._s_in.wait_event(self._ev_run[slot])          # the slot's device parameters are no longer being read
        with torch.cuda.stream(self._s_in):
            self._params[slot].copy_(self.host_params[slot], non_blocking=True)
            self._ev_in[slot].record(self._s_in)
        self._s_reset.wait_event(self._ev_run[slot])       # ... nor is the slot's env state
        with torch.cuda.stream(self._s_reset):             # resets in submission order: every env's context stream advances once
            env.reset_into(self._state[slot])
            self._ev_reset[slot].record(self._s_reset)
        run = self._s_run[slot]
        run.wait_event(self._ev_in[slot])
        run.wait_event(self._ev_reset[slot])
        run.wait_event(self._ev_out[slot])                 # the result set this step overwrites has been copied out
        with torch.cuda.stream(run):
            _obs, ret, terminated, _trunc, info = env.run_episode(self._params[slot], self._state[slot])
            self._ev_run[slot].record(run)
        self._s_out.wait_event(self._ev_run[slot])
        with torch.cuda.stream(self._s_out):
            self.host_ret[slot].copy_(ret, non_blocking=True)
            self.host_len[slot].copy_(info["trajectory_length"], non_blocking=True)
            self.host_terminated[slot].copy_(terminated, non_blocking=True)
            self._ev_out[slot].record(self._s_out)
        self._pending[slot] = True
        self._next = (slot + 1) % self.SLOTS

    def wait(self, slot: int):
        """blocks until the results of the batch submitted in `slot` are in pinned host memory and returns them"""
        if not self._pending[slot]:
            raise RuntimeError(f"nothing was submitted in slot {slot}")
        self._ev_out[slot].synchronize()
        self._pending[slot] = False
        return self.host_ret[slot], self.host_len[slot], self.host_terminated[slot]

    def drain(self):
        """waits for everything in flight (results stay readable)"""
        for s in (self._s_in, self._s_reset, self._s_out, *self._s_run):
            s.synchronize()
