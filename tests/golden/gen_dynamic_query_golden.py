"""Golden snapshot-query results of the UNMODIFIED reference under its dynamic (RawBackend) backend.

    bash oracle/build_ref.sh && python tests/golden/gen_dynamic_query_golden.py

The RawBackend answers ``snapshot_list[node][ticks:nodes:attrs]`` as a 4-D (ticks, nodes, attrs, max_slots) array, NaN for
missing slots / unknown ticks, values through float32 (maro/backends/raw/snapshotlist.cpp:244-318,
_raw_backend_.pyx:263-315) — unlike the static backend's flat zero-padded float64.  Output: cim_dynamic_queries.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))

TOPOLOGY, DURATIONS = "toy.5p_ssddd_l0.3", 60
QUERIES = {
    "q_ports_some": ("ports", [3, 10, 1000], [0, 2], ["empty", "full"]),
    "q_vessel_lists": ("vessels", [5, 59], [0, 1, 2], ["remaining_space", "future_stop_list", "past_stop_list"]),
    "q_matrices": ("matrices", [59], [], ["vessel_plans", "full_on_ports"]),
    "q_ports_all_ticks": ("ports", [], [], ["shortage"]),
    "q_single": ("vessels", 7, 1, "full"),
    "q_float": ("ports", [20, 21], [], ["transfer_cost", "acc_booking"]),
}


def key_of(spec):
    node, ticks, nodes, attrs = spec
    t = ticks if not isinstance(ticks, list) or ticks else None
    n = nodes if not isinstance(nodes, list) or nodes else None
    return node, slice(t, n, attrs)


def run_queries(env):
    out = {}
    for name, spec in QUERIES.items():
        node, sl = key_of(spec)
        out[name] = np.asarray(env.snapshot_list[node][sl], np.float64)
    return out


def drive(env):
    """null actions except a LOAD of half the scope at every third decision (so transfer_cost is non-zero)"""
    from maro.simulator.scenarios.cim.common import Action, ActionType

    metrics, ev, done = env.step(None)
    k = 0
    while not done:
        act = Action(ev.vessel_idx, ev.port_idx, ev.action_scope.load // 2, ActionType.LOAD) if k % 3 == 0 else None
        k += 1
        metrics, ev, done = env.step(act)


if __name__ == "__main__":
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    os.environ["DEFAULT_BACKEND_NAME"] = "dynamic"
    sys.path[:0] = [os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "oracle", "_ref", "_stubs")]
    from maro.simulator import Env

    env = Env("cim", TOPOLOGY, durations=DURATIONS)
    drive(env)
    out = run_queries(env)
    for k, v in out.items():
        print(k, v.shape, "nan:", int(np.isnan(v).sum()))
    np.savez_compressed(os.path.join(HERE, "cim_dynamic_queries.npz"), **out)
