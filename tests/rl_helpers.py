"""numpy restatement of the reference's CIM RL shaping (examples/cim/rl/env_sampler.py:15-36, 66-80) over raw
snapshot rows — test infrastructure (the checker of the device shaping kernels)."""
import importlib.util
import os

import numpy as np

from maro_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("gen_cim_rl_golden", os.path.join(HERE, "golden", "gen_cim_rl_golden.py"))
rl_gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(rl_gen)
RL_CASES = rl_gen.CASES


def load_rl_golden(name):
    return np.load(os.path.join(HERE, "golden", f"cim_rl_{name}.npz"))


class SnapshotView:
    """snapshot_list-style reads over `snapshot(frame) -> int32 frame words or None` (static-backend semantics:
    a frame the ring does not hold reads as zeros)."""

    def __init__(self, snapshot_fn, topo, cache=False):
        self.fn, self.topo, self._cache, self._use_cache = snapshot_fn, topo, {}, cache
        self.lay, _ = _abi.frame_layout(topo.n_ports, topo.n_vessels, topo.past_stop_number, topo.future_stop_number)

    def get(self, node, frame, index, attr):
        off, n, slots = self.lay[node][attr]
        if frame < 0:
            row = None
        elif not self._use_cache:
            row = self.fn(int(frame))
        else:  # only valid while the env does not advance
            if int(frame) not in self._cache:
                self._cache[int(frame)] = self.fn(int(frame))
            row = self._cache[int(frame)]
        if row is None:
            return np.zeros(slots, np.float64)
        w = row[off + index * slots: off + (index + 1) * slots]
        return (w.view(np.float32) if attr == "transfer_cost" else w).astype(np.float64)


def state_numpy(view, tick, port, vessel):
    ticks = [max(0, tick - rt) for rt in range(rl_gen.LOOK_BACK - 1)]
    future = view.get("vessels", tick, vessel, "future_stop_list").astype("int")
    out = []
    for t in ticks:
        for p in [port] + list(future):
            for a in rl_gen.PORT_ATTRIBUTES:
                out.append(view.get("ports", t, int(p), a))
    for a in rl_gen.VESSEL_ATTRIBUTES:
        out.append(view.get("vessels", tick, vessel, a))
    return np.concatenate(out)


def reward_numpy(view, port, tick):
    R = rl_gen.REWARD
    ticks = range(tick + 1, tick + 1 + R["time_window"])
    ff = np.asarray([view.get("ports", t, port, "fulfillment")[0] for t in ticks])
    fs = np.asarray([view.get("ports", t, port, "shortage")[0] for t in ticks])
    decay = [R["time_decay"] ** i for i in range(R["time_window"])]
    return np.float32(R["fulfillment_factor"] * np.dot(ff, decay) - R["shortage_factor"] * np.dot(fs, decay))


def action_numpy(view, dec, model_action):
    """env_sampler.py:38-64 over a decision row [tick, port, vessel, scope.load, scope.discharge, early_discharge]"""
    tick, port, vessel, load, discharge = (int(x) for x in dec[:5])
    space = rl_gen.ACTION_SPACE
    vsl_space = view.get("vessels", tick, vessel, "remaining_space")[0] if rl_gen.FINITE_VESSEL_SPACE else float("inf")
    percent = abs(space[model_action])
    if model_action < len(space) / 2:
        return [vessel, port, int(min(round(percent * load), vsl_space)), 0]
    early = view.get("vessels", tick, vessel, "early_discharge")[0] if rl_gen.HAS_EARLY_DISCHARGE else 0
    plan = percent * (discharge + early) - early
    return [vessel, port, int(round(plan) if plan > 0 else round(percent * discharge)), 1]
