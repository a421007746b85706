"""``maro`` import shim — lets code written against the reference run **unchanged** on the CUDA core.

    import maro_b200.shim; maro_b200.shim.install()      # or: python -m maro_b200.shim script.py [args...]
    from maro.simulator import Env                       # -> maro_b200.simulator.Env
    from maro.simulator.scenarios.cim.common import Action, ActionType, DecisionEvent
    from maro.vector_env import VectorEnv

Only the modules on the ``Env.step`` path are replaced (``maro.simulator``, ``maro.simulator.core`` / ``abs_core``,
``maro.simulator.scenarios.{cim,citi_bike,vm_scheduling}[.common]``, ``maro.vector_env``).  When a real ``maro``
distribution is importable (e.g. for ``maro.rl``), it stays in place underneath: every other submodule — ``maro.rl``,
``maro.utils``, ``maro.simulator.utils`` … — resolves to it, and ``maro.rl.rollout.AbsEnvSampler`` (which does
``from maro.simulator import Env``, maro/rl/rollout/env_sampler.py:18) picks up the CUDA-backed ``Env``.
Without one, a bare ``maro`` namespace is created.

Reference surfaces mirrored: maro/simulator/__init__.py:5-9, maro/vector_env/__init__.py,
maro/simulator/scenarios/cim/common.py, citi_bike/common.py, vm_scheduling/__init__.py.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
from typing import Dict, List, Optional

_INSTALLED: Dict[str, Optional[types.ModuleType]] = {}


def _real_package_dir(name: str) -> Optional[str]:
    """directory of the real (non-shim) ``maro`` package on sys.path, without importing it"""
    try:
        spec = importlib.util.find_spec(name)
    except (ImportError, ValueError):
        return None
    if spec is None or not spec.submodule_search_locations:
        return None
    return list(spec.submodule_search_locations)[0]


def _module(name: str, attrs: dict, path: Optional[List[str]], doc: str) -> types.ModuleType:
    m = types.ModuleType(name, doc)
    m.__dict__.update(attrs)
    if path is not None:
        m.__path__ = path  # a package: unknown submodules fall through to the real distribution's directory (if any)
    m.__maro_b200_shim__ = True
    return m


def install() -> None:
    """Register the CUDA-backed modules under the reference's import names (idempotent)."""
    if _INSTALLED:
        return
    from . import simulator as sim, vector_env as venv
    from .scenarios.cim import common as cim_common
    from .scenarios.citi_bike import common as bike_common
    from .scenarios import vm_scheduling as vm_pkg
    from .scenarios.vm_scheduling import common as vm_common

    existing = sys.modules.get("maro")
    if existing is not None and not getattr(existing, "__maro_b200_shim__", False):
        real_dir = os.path.dirname(getattr(existing, "__file__", "") or "") or None
    else:
        real_dir = _real_package_dir("maro")

    def sub(*parts) -> List[str]:
        d = os.path.join(real_dir, *parts) if real_dir else None
        return [d] if d and os.path.isdir(d) else []

    mods: Dict[str, types.ModuleType] = {}
    if existing is None or getattr(existing, "__maro_b200_shim__", False):
        if real_dir:  # import the real top-level package (cheap: its __init__ only sets __version__ / data paths)
            try:
                importlib.import_module("maro")
            except Exception:
                mods["maro"] = _module("maro", {}, [real_dir], "maro namespace (maro_b200 shim)")
        else:
            mods["maro"] = _module("maro", {}, [], "maro namespace (maro_b200 shim)")

    # AbsEnv (maro/simulator/abs_core.py:25) is the isinstance anchor of callers: the CUDA-backed Env plays both roles
    sim_attrs = {"Env": sim.Env, "DecisionMode": sim.DecisionMode, "AbsEnv": sim.Env,
                 "__all__": ["AbsEnv", "Env", "DecisionMode"]}
    mods["maro.simulator"] = _module("maro.simulator", sim_attrs, sub("simulator"), "maro_b200.simulator under the reference's name")
    mods["maro.simulator.core"] = _module("maro.simulator.core", {"Env": sim.Env}, None, "")
    mods["maro.simulator.abs_core"] = _module("maro.simulator.abs_core", {"AbsEnv": sim.Env, "DecisionMode": sim.DecisionMode}, None, "")
    mods["maro.simulator.scenarios"] = _module("maro.simulator.scenarios", {}, sub("simulator", "scenarios"), "")

    def scenario(name: str, common_mod, extra: dict):
        attrs = {k: getattr(common_mod, k) for k in dir(common_mod) if not k.startswith("_")}
        full = f"maro.simulator.scenarios.{name}"
        mods[full + ".common"] = _module(full + ".common", attrs, None, common_mod.__doc__ or "")
        pkg = _module(full, dict(extra), sub("simulator", "scenarios", name), "")
        pkg.common = mods[full + ".common"]
        mods[full] = pkg
        setattr(mods["maro.simulator.scenarios"], name, pkg)

    scenario("cim", cim_common, {})
    scenario("citi_bike", bike_common, {})
    scenario("vm_scheduling", vm_common, {k: getattr(vm_pkg, k) for k in vm_pkg.__all__})
    mods["maro.simulator"].scenarios = mods["maro.simulator.scenarios"]
    mods["maro.simulator"].core = mods["maro.simulator.core"]
    mods["maro.simulator"].abs_core = mods["maro.simulator.abs_core"]
    mods["maro.vector_env"] = _module("maro.vector_env", {"VectorEnv": venv.VectorEnv, "__all__": ["VectorEnv"]}, [], "")
    mods["maro.vector_env.vector_env"] = _module("maro.vector_env.vector_env", {"VectorEnv": venv.VectorEnv}, None, "")

    for name, m in mods.items():
        _INSTALLED[name] = sys.modules.get(name)
        sys.modules[name] = m
    top = sys.modules["maro"]
    top.simulator = mods["maro.simulator"]
    top.vector_env = mods["maro.vector_env"]
    _INSTALLED.setdefault("maro", None)


def uninstall() -> None:
    """Undo ``install`` (tests)."""
    for name, prev in list(_INSTALLED.items()):
        cur = sys.modules.get(name)
        if cur is not None and getattr(cur, "__maro_b200_shim__", False):
            if prev is None:
                del sys.modules[name]
            else:
                sys.modules[name] = prev
    top = sys.modules.get("maro")
    if top is not None:
        for attr in ("simulator", "vector_env"):
            if getattr(getattr(top, attr, None), "__maro_b200_shim__", False):
                delattr(top, attr)
    _INSTALLED.clear()


def installed() -> bool:
    return bool(_INSTALLED)


def main(argv: List[str]) -> int:
    """``python -m maro_b200.shim script.py [args...]`` — run an unmodified reference script on the CUDA core."""
    import runpy

    if not argv:
        print(__doc__)
        return 2
    install()
    sys.argv = list(argv)
    sys.path.insert(0, os.path.dirname(os.path.abspath(argv[0])))
    runpy.run_path(argv[0], run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
