"""profiles/*_ncu_summary.txt (written by profiles/ncu_summary.py from `ncu --set full` captures) -> profiles/ncu_numbers.json,
the ONLY place bench.py takes ncu-derived numbers from (dram bytes per launch, executed warp instructions, pipe
utilisations).  tests/test_bench_contract.py re-derives the JSON from the text files, so the two cannot drift apart.

    python tools/ncu_to_json.py            # rewrites profiles/ncu_numbers.json
"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# key in the JSON -> summary file (one kernel each; the first kernel block of the file is used)
SOURCES = {
    "rollout_config2": "r2_rollout_ncu_summary.txt",
    "rollout_config2_sigma1": "r2_rollout_sigma1_ncu_summary.txt",
    "rollout_config5_1m_envs": "r2_rollout_1m_envs_ncu_summary.txt",
    "rollout_1m_envs_sigma1": "r2_rollout_1m_sigma1_ncu_summary.txt",
    "rollout_learned_tau_delay": "r2_rollout_learned_tau_ncu_summary.txt",
    "rollout_config3_viapoint_dmp": "r2_rollout_viapoint_dmp_ncu_summary.txt",
    "rollout_config4_simple_prodmp_plans": "r2_rollout_simple_prodmp_plans_ncu_summary.txt",
    "trajgen_promp": "r2_trajgen_promp_ncu_summary.txt",
    "trajgen_dmp": "r2_trajgen_dmp_ncu_summary.txt",
    "trajgen_prodmp": "r2_trajgen_prodmp_ncu_summary.txt",
    "trajgen_phase_promp": "r2_trajgen_phase_promp_ncu_summary.txt",
    "dmp_integrate_phase": "r2_dmp_integrate_phase_ncu_summary.txt",
    "reset": "r2_reset_ncu_summary.txt",
    "cov_simt": "r2_cov_simt_ncu_summary.txt",
    "cov_umma": "r2_cov_umma_ncu_summary.txt",
}
_UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
_TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def parse_summary(path):
    """first kernel block of a summary file -> dict"""
    out, seen = {}, False
    for line in open(path):
        if line.startswith("kernel:"):
            if seen:
                break
            seen = True
            out["kernel"] = line[len("kernel:"):].split("|")[0].strip()
            continue
        m = re.match(r"\s+(\S+)\s+([-\d.eE+]+)\s*(\S*)", line)
        if m:
            out[m.group(1)] = (float(m.group(2)), m.group(3))
    return out


def numbers_of(path, name):
    d = parse_summary(path)

    def val(key, table=None):
        if key not in d:
            return None
        v, u = d[key]
        return v * (table or {}).get(u, 1)

    rd, wr = val("dram__bytes_read.sum", _UNIT), val("dram__bytes_write.sum", _UNIT)
    pipes = {k: val(f"sm__inst_executed_pipe_{k}.avg.pct_of_peak_sustained_active") for k in ("fma", "alu", "fp64", "xu", "lsu")}
    pipes["issue_active"] = val("smsp__issue_active.avg.pct_of_peak_sustained_active")
    pipes["tensor"] = val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    return dict(kernel=d.get("kernel"), kernel_ms=val("gpu__time_duration.sum", _TIME),
                dram_bytes=None if rd is None else int(round(rd + (wr or 0))), dram_read_bytes=rd, dram_write_bytes=wr,
                warp_instructions=None if val("smsp__inst_executed.sum") is None else int(val("smsp__inst_executed.sum")),
                threads_per_instruction=val("smsp__thread_inst_executed_per_inst_executed.ratio"),
                registers=val("launch__registers_per_thread"), pipes=pipes, source=f"profiles/{name}")


def build():
    out = {}
    for key, name in SOURCES.items():
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            out[key] = numbers_of(path, name)
    return out


if __name__ == "__main__":
    with open(os.path.join(ROOT, "profiles", "ncu_numbers.json"), "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote profiles/ncu_numbers.json:", ", ".join(build()))
