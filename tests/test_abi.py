"""The C-ABI boundary without a GPU: the shared library loads, exports every function include/fancy_gym_b200.h
declares, and the ctypes mirror of the structs has the layout a C compiler gives the header."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fancy_gym_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fancy_gym_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 9
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} is declared in the header but not exported by {_lib.LIB_PATH}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    assert _lib.lib.fg_abi_version() == 2


def test_header_is_plain_c_and_struct_layout_matches_ctypes(tmp_path):
    from fancy_gym_b200 import _lib
    prog = tmp_path / "layout.c"
    fields_cfg = [f[0] for f in _lib.FgConfig._fields_]
    fields_io = [f[0] for f in _lib.FgRolloutIO._fields_]
    fields_rc = [f[0] for f in _lib.FgResetCfg._fields_]
    fields_ri = [f[0] for f in _lib.FgResetIO._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){',
             'printf("fg_config %zu\\n", sizeof(fg_config));', 'printf("fg_rollout_io %zu\\n", sizeof(fg_rollout_io));',
             'printf("fg_reset_cfg %zu\\n", sizeof(fg_reset_cfg));', 'printf("fg_reset_io %zu\\n", sizeof(fg_reset_io));']
    lines += [f'printf("fg_config.{f} %zu\\n", offsetof(fg_config, {f}));' for f in fields_cfg]
    lines += [f'printf("fg_rollout_io.{f} %zu\\n", offsetof(fg_rollout_io, {f}));' for f in fields_io]
    lines += [f'printf("fg_reset_cfg.{f} %zu\\n", offsetof(fg_reset_cfg, {f}));' for f in fields_rc]
    lines += [f'printf("fg_reset_io.{f} %zu\\n", offsetof(fg_reset_io, {f}));' for f in fields_ri]
    lines += ['return 0;}']
    prog.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", str(prog), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    assert int(out["fg_config"]) == C.sizeof(_lib.FgConfig)
    assert int(out["fg_rollout_io"]) == C.sizeof(_lib.FgRolloutIO)
    assert int(out["fg_reset_cfg"]) == C.sizeof(_lib.FgResetCfg) and int(out["fg_reset_io"]) == C.sizeof(_lib.FgResetIO)
    for f in fields_rc:
        assert int(out[f"fg_reset_cfg.{f}"]) == getattr(_lib.FgResetCfg, f).offset, f
    for f in fields_ri:
        assert int(out[f"fg_reset_io.{f}"]) == getattr(_lib.FgResetIO, f).offset, f
    for f in fields_cfg:
        assert int(out[f"fg_config.{f}"]) == getattr(_lib.FgConfig, f).offset, f
    for f in fields_io:
        assert int(out[f"fg_rollout_io.{f}"]) == getattr(_lib.FgRolloutIO, f).offset, f


def test_invalid_arguments_are_reported_not_crashed():
    """argument validation happens before any CUDA call, so it is testable without a device"""
    from fancy_gym_b200 import _lib
    hp = C.c_void_p()
    cfg = _lib.FgConfig()
    cfg.struct_size = 3                                   # wrong size: ABI mismatch must be refused
    st = _lib.lib.fg_create(C.byref(cfg), 0, C.byref(hp))
    assert st == _lib.ERR_INVALID and b"struct_size" in _lib.lib.fg_last_error()
    with pytest.raises(ValueError):
        _lib.check(st)
    assert _lib.lib.fg_destroy(None) in (_lib.OK, _lib.ERR_INVALID)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under fancy_gym_b200/ may reference it"""
    bad = []
    for r, _, fs in os.walk(os.path.join(ROOT, "fancy_gym_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(r, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    bad.append(os.path.join(r, f))
    assert not bad, bad


def test_hot_kernels_keep_their_register_budget():
    """Occupancy guard (no GPU needed): the fused HoleReacher/ProMP rollout must stay <= 128 registers per thread (4 blocks
    of 128 threads per SM: 65,536 envs need 443 resident threads per SM), the covariance kernel <= 64."""
    import shutil
    from fancy_gym_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    regs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    hot = [n for n in regs if "k_rolloutILi0ELi0ELb0ELi5ELi5ELb0" in n]
    assert hot and all(regs[n] <= 128 for n in hot), {n: regs[n] for n in hot}
    assert all(v <= 128 for n, v in regs.items() if "k_rollout" in n), "a rollout instantiation exceeds 128 registers"
    cov = [n for n in regs if "k_cov_simtILi5" in n]
    assert cov and all(regs[n] <= 64 for n in cov)


def test_integration_stub_lists_the_abi_structs_field_for_field():
    """the ctypes stub printed in INTEGRATION.md (what a fancy_gym maintainer would paste) has to stay the ABI"""
    import re
    from fancy_gym_b200 import _lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class FgConfig\(C\.Structure\):.*?_fields_ = \[(.*?)\]\n\nclass", doc, re.S)
    assert re.findall(r'\("(\w+)"', m.group(1)) == [f[0] for f in _lib.FgConfig._fields_]
    m = re.search(r"class FgRolloutIO\(C\.Structure\):.*?\n\ndef check", doc, re.S)
    assert re.findall(r'"(\w+)"', m.group(0)) == [f[0] for f in _lib.FgRolloutIO._fields_]
