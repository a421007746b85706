"""Shared test helpers: golden-case loading and a generic episode driver."""
import importlib.util
import os

import numpy as np

from maro_b200 import _abi
from maro_b200.scenarios.cim.topology import build_topology

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("gen_cim_golden", os.path.join(GOLDEN, "gen_cim_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
CASES = gen.CASES
policy_random_py = gen.policy_random


def case_topology(spec):
    topo = spec["topology"]
    if topo.endswith("_case_cfg"):
        topo = os.path.join(GOLDEN, "cim_case_config.json")
    max_tick = spec.get("start_tick", 0) + spec["durations"]
    t = build_topology(topo, max_tick, seed=spec.get("topo_seed"))
    if spec.get("reset_new_seed"):  # second episode after Env.reset(keep_seed=False)
        from maro_b200.scenarios.cim.topology import next_topology_seed

        t = build_topology(topo, max_tick, seed=next_topology_seed(t))
    return t


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"cim_{name}.npz"))


def drive(step_fn, spec, max_steps=100000):
    """step_fn(actions or None) -> (status, dec[8], metrics[3]).  Returns (rows[n][9], final_metrics, final dec)."""
    rows = []
    st, dec, met = step_fn(None)
    step = 0
    while st == _abi.STATUS_DECISION and step < max_steps:
        rows.append(list(dec[:6]) + list(met))
        if spec["policy"] == 1:
            act = np.asarray([policy_random_py([int(x) for x in dec[:6]], spec.get("pseed", 0),
                                               spec.get("replica", 0), step)], np.int32)
        elif spec["policy"] == 2:  # action list (LOAD then DISCHARGE), every 7th step none
            act = None if step % 7 == 3 else np.asarray(gen.policy_pair([int(x) for x in dec[:6]], spec, step), np.int32)
        else:
            act = None
        step += 1
        st, dec, met = step_fn(act)
    return np.asarray(rows, np.int64).reshape(-1, 9), np.asarray(met, np.int64), dec, st


def named_frames(words_by_frame, topo):
    """dict 'ports/empty' -> array [frames, nodes(, slots)] from raw frame word rows [frames][FW]."""
    lay, fw = _abi.frame_layout(topo.n_ports, topo.n_vessels, topo.past_stop_number, topo.future_stop_number)
    w = np.asarray(words_by_frame, np.int32)
    out = {}
    for node, attrs in lay.items():
        for a, (off, n, slots) in attrs.items():
            x = w[:, off:off + n * slots]
            if node == "ports":
                x = x.reshape(len(w), n)
                if a == "transfer_cost":
                    x = x.view(np.float32)
            elif node == "vessels":
                x = x.reshape(len(w), n, slots)
            else:
                x = x.reshape(len(w), slots)
            out[f"{node}/{a}"] = x
    return out


def assert_snapshots_equal(get_snapshot, gold, topo):
    frames = gold["frames"].tolist()
    rows = []
    for f in frames:
        s = get_snapshot(int(f))
        assert s is not None, f"frame {f} missing from ring"
        rows.append(s)
    named = named_frames(rows, topo)
    for key, val in named.items():
        g = gold[key]
        assert val.shape == g.shape, (key, val.shape, g.shape)
        if not np.array_equal(val, g):
            bad = np.argwhere(val != g)[0]
            raise AssertionError(f"{key} differs first at {bad.tolist()} (frame {frames[bad[0]]}): "
                                 f"got {val[tuple(bad)]} want {g[tuple(bad)]}")
