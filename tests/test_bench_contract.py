"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys; the argument
parser accepts the driver's flags."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_ncu_numbers_json_is_what_the_committed_summaries_say():
    """bench.py takes every ncu-derived number (dram bytes, executed instructions, pipe shares) from profiles/ncu_numbers.json;
    that file must be exactly what tools/ncu_to_json.py derives from the committed profiles/*_ncu_summary.txt, and bench.py
    itself must not carry such literals"""
    sys.path.insert(0, ROOT)
    from tools import ncu_to_json
    with open(os.path.join(ROOT, "profiles", "ncu_numbers.json")) as f:
        stored = json.load(f)
    assert stored == json.loads(json.dumps(ncu_to_json.build()))
    assert "rollout_config2" in stored and stored["rollout_config2"]["warp_instructions"] > 0
    assert stored["rollout_config2"]["dram_bytes"] > 0 and stored["trajgen_promp"]["dram_bytes"] > 0
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "NCU_TRAFFIC" not in src and "NCU_WARP_INSTRUCTIONS" not in src and "NCU_PIPES" not in src
