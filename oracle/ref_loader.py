"""Load the reference's *own* hot-path files, unmodified, from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference cannot be imported as a
package here (gymnasium, mp_pytorch, mujoco, matplotlib are not installed), so
this loader
  * puts oracle/refstub (a ~100-line stand-in for gymnasium / matplotlib / mp_pytorch type
    stubs) on sys.path, and
  * pre-seeds `fancy_gym` and its sub-packages in sys.modules as empty namespace modules whose
    __path__ points at the real directories, so the heavy package __init__ files are skipped
    while every *sub-module* resolves to the reference's file.
Nothing is copied: the code that runs is the file under /root/reference.  /root/reference does
not exist on the GPU box, so this module is only used here to pin the oracle restatement and to
generate tests/golden/*.npz (tests/golden/make_golden.py).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("FANCY_GYM_REFERENCE", "/root/reference")
_STUB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refstub")

_NAMESPACE_PKGS = [
    "fancy_gym",
    "fancy_gym.envs",
    "fancy_gym.envs.classic_control",
    "fancy_gym.envs.classic_control.base_reacher",
    "fancy_gym.black_box",
    "fancy_gym.black_box.factory",
    "fancy_gym.black_box.controller",
    "fancy_gym.utils",
]
# sub-packages whose own __init__ is light-weight and must run (they define `MPWrapper`)
_REAL_PKGS = [
    "fancy_gym.envs.classic_control.hole_reacher",
    "fancy_gym.envs.classic_control.viapoint_reacher",
    "fancy_gym.envs.classic_control.simple_reacher",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "fancy_gym", "envs", "classic_control"))


_loaded = False


def _install():
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if "gymnasium" in sys.modules and not getattr(sys.modules["gymnasium"], "__file__", "").startswith(_STUB_DIR):
        raise RuntimeError("a real gymnasium is already imported; ref_loader expects its own stub")
    sys.path.insert(0, _STUB_DIR)
    for name in _NAMESPACE_PKGS:
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
        mod.__package__ = name
        sys.modules[name] = mod
    _loaded = True


def load():
    """Returns a namespace with the reference classes of the hot path."""
    _install()
    ns = types.SimpleNamespace()
    ns.gym = importlib.import_module("gymnasium")
    hr = importlib.import_module("fancy_gym.envs.classic_control.hole_reacher.hole_reacher")
    vp = importlib.import_module("fancy_gym.envs.classic_control.viapoint_reacher.viapoint_reacher")
    sr = importlib.import_module("fancy_gym.envs.classic_control.simple_reacher.simple_reacher")
    ns.HoleReacherEnv = hr.HoleReacherEnv
    ns.ViaPointReacherEnv = vp.ViaPointReacherEnv
    ns.SimpleReacherEnv = sr.SimpleReacherEnv
    ns.MPWrapper_HoleReacher = importlib.import_module(
        "fancy_gym.envs.classic_control.hole_reacher.mp_wrapper").MPWrapper
    ns.MPWrapper_ViaPoint = importlib.import_module(
        "fancy_gym.envs.classic_control.viapoint_reacher.mp_wrapper").MPWrapper
    ns.MPWrapper_SimpleReacher = importlib.import_module(
        "fancy_gym.envs.classic_control.simple_reacher.mp_wrapper").MPWrapper
    ns.BlackBoxWrapper = importlib.import_module("fancy_gym.black_box.black_box_wrapper").BlackBoxWrapper
    ns.get_controller = importlib.import_module("fancy_gym.black_box.factory.controller_factory").get_controller
    ns.TimeLimit = ns.gym.TimeLimit
    return ns


# registration kwargs of the three BASELINE envs, transcribed from
# fancy_gym/envs/__init__.py:38-87 (the registry itself needs the real gymnasium)
REF_ENV_KWARGS = {
    "HoleReacher-v0": dict(n_links=5, random_start=True, allow_self_collision=False,
                           allow_wall_collision=False, hole_width=None, hole_depth=1, hole_x=None,
                           collision_penalty=100),
    "ViaPointReacher-v0": dict(n_links=5, allow_self_collision=False, collision_penalty=1000),
    "SimpleReacher-v0": dict(n_links=2),
    "LongSimpleReacher-v0": dict(n_links=5),
}


def make_step_env(ns, name, **overrides):
    cls = {"HoleReacher-v0": ns.HoleReacherEnv, "ViaPointReacher-v0": ns.ViaPointReacherEnv,
           "SimpleReacher-v0": ns.SimpleReacherEnv, "LongSimpleReacher-v0": ns.SimpleReacherEnv}[name]
    return cls(**{**REF_ENV_KWARGS[name], **overrides})
