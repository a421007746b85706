"""Golden CIM traces of the UNMODIFIED reference in DecisionMode.Joint (maro/simulator/core.py:354-366): every decision event
of a tick is returned in one list, the action scopes are evaluated before any of the tick's actions is applied, the answers
are applied in list order; a shorter answer list leaves the remaining decision events without an action.

    bash oracle/build_ref.sh && python tests/golden/gen_cim_joint_golden.py

Per env-step the trace records one row per decision [step, tick, port, vessel, scope.load, scope.discharge, early_discharge]
and the metrics; the answer tape is a function of (decision ordinal): hashed random quantity inside the scope, every 5th
decision `None`, and on steps with >= 3 decisions the last one is left unanswered.  Output: tests/golden/cim_joint_<case>.npz."""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))

CASES = {
    "toy4p_l00_160": dict(topology="toy.4p_ssdd_l0.0", durations=160, pseed=3),
    "toy5p_l03_140_res4": dict(topology="toy.5p_ssddd_l0.3", durations=140, pseed=9, snapshot_resolution=4, max_snapshots=12),
    "gt22p_l08_70": dict(topology="global_trade.22p_l0.8", durations=70, pseed=1),
}


def hash_u32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def answers(rows, pseed, ordinal):
    """rows: [[tick, port, vessel, load, discharge, early], ...] of one step -> list of (vessel, port, qty, type) or None;
    may be shorter than rows.  `ordinal` = number of decisions seen before this step."""
    out = []
    n = len(rows) - 1 if len(rows) >= 3 else len(rows)
    for k in range(n):
        o = ordinal + k
        if o % 5 == 4:
            out.append(None)
            continue
        d = rows[k]
        h = hash_u32(pseed * 0x9E3779B9 + o * 0x85EBCA6B + 0x51ED270B)
        to_dis = d[4] > 0 and (h & 1)
        # quantities stay small: scopes are evaluated before the step's earlier answers are applied, so two vessels at one
        # port must not exhaust it together (the reference asserts on over-scope actions)
        scope = (d[4] if to_dis else d[3]) // 4
        qty = (h >> 3) % (scope + 1) if scope > 0 else 0
        out.append((d[2], d[1], int(qty), 1 if to_dis else 0))
    return out


def run_case(name, spec):
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    sys.path[:0] = [os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "oracle", "_ref", "_stubs")]
    from maro.simulator import DecisionMode, Env
    from maro.simulator.scenarios.cim.common import Action, ActionType

    env = Env("cim", spec["topology"], durations=spec["durations"], snapshot_resolution=spec.get("snapshot_resolution", 1),
              max_snapshots=spec.get("max_snapshots"), decision_mode=DecisionMode.Joint)
    rows, mets, step, ordinal = [], [], 0, 0
    metrics, decs, done = env.step(None)
    while not done:
        cur = [[d.tick, d.port_idx, d.vessel_idx, d.action_scope.load, d.action_scope.discharge, d.early_discharge] for d in decs]
        for r in cur:
            rows.append([step] + r)
        mets.append([int(metrics["order_requirements"]), int(metrics["container_shortage"]), int(metrics["operation_number"])])
        acts = [None if a is None else Action(a[0], a[1], a[2], ActionType.DISCHARGE if a[3] else ActionType.LOAD)
                for a in answers(cur, spec["pseed"], ordinal)]
        ordinal += len(cur)
        step += 1
        metrics, decs, done = env.step(acts)
    sl = env.snapshot_list
    frames = sorted(sl.get_frame_index_list())
    P, V = len(sl["ports"]), len(sl["vessels"])
    out = {"rows": np.asarray(rows, np.int64).reshape(-1, 7), "step_metrics": np.asarray(mets, np.int64).reshape(-1, 3),
           "final_metrics": np.asarray([int(metrics[k]) for k in ("order_requirements", "container_shortage", "operation_number")], np.int64),
           "frames": np.asarray(frames, np.int32)}
    for a in ("empty", "full", "on_consignee", "on_shipper", "shortage", "acc_shortage", "transfer_cost"):
        out["ports/" + a] = sl["ports"][frames::a].reshape(len(frames), P)
    for a in ("empty", "full", "remaining_space", "early_discharge"):
        out["vessels/" + a] = sl["vessels"][frames::a].reshape(len(frames), V)
    out["matrices/vessel_plans"] = sl["matrices"][frames::"vessel_plans"].reshape(len(frames), -1)
    np.savez_compressed(os.path.join(HERE, f"cim_joint_{name}.npz"), **out)
    print(name, "steps", step, "decisions", len(rows), "final", out["final_metrics"].tolist(), flush=True)


if __name__ == "__main__":
    ctx = mp.get_context("spawn")
    for n in sys.argv[1:] or list(CASES):
        p = ctx.Process(target=run_case, args=(n, CASES[n]))
        p.start()
        p.join()
        if p.exitcode != 0:
            raise SystemExit(f"case {n} failed")
