"""Bit-level regression of the fused rollout: every output of 45 configurations (all twelve ids at two parameter scales,
reward functions, constructor options, replanning +- condition_on_desired, ragged sub-trajectories, learned tau / delay,
position control, run-time basis counts, the BASELINE sizes) must hash to the digests recorded with the round-1 kernel
(tests/golden/rollout_digests.json, tools/kernel_digests.py).  Scheduling changes inside the kernel — re-packing of live
envs, plans looped inside one launch — are only accepted bit-identical."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from tools.kernel_digests import CASES, digest_case  # noqa: E402


@pytest.fixture(scope="module")
def recorded(golden_dir):
    with open(os.path.join(golden_dir, "rollout_digests.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_rollout_outputs_hash_to_the_recorded_digests(case, recorded):
    import fancy_gym_b200 as fancy_gym
    if recorded["cuda"] != torch.version.cuda:
        pytest.skip("digests were recorded with another CUDA toolkit (device libm)")
    digest, steps = digest_case(fancy_gym, case, torch.device("cuda", 0))
    want = recorded["cases"][case[0]]
    assert steps == want["env_steps"], (case[0], steps, want["env_steps"])
    assert digest == want["sha256"], case[0]
