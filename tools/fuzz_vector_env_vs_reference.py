"""Randomised check of ``maro_b200.vector_env.VectorEnv`` (host code over the emulator-backed batch) against the
reference's multi-process ``maro.vector_env.VectorEnv``: random sequences of broadcast / list / dict steps with the
hashed random agent, comparing the returned metrics / decision lists, ``is_done``, ``tick`` and ``frame_index`` after
every call, then a query through ``snapshot_list``.  Build-container tool (needs oracle/_ref).

    python tools/fuzz_vector_env_vs_reference.py [n_cases] [first_seed]
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def script_of(seed):
    rng = np.random.default_rng(seed)
    topo = str(rng.choice(["toy.4p_ssdd_l0.0", "toy.4p_ssdd_l0.5", "toy.5p_ssddd_l0.3"]))
    B = int(rng.integers(2, 4))
    ops = []
    for _ in range(int(rng.integers(25, 70))):
        r = rng.random()
        if r < 0.5:
            ops.append(("list",))
        elif r < 0.8:
            ops.append(("dict", sorted(int(x) for x in rng.choice(B, int(rng.integers(1, B + 1)), replace=False))))
        else:
            ops.append(("none",))
    return dict(topology=topo, durations=int(rng.integers(30, 90)), batch=B, ops=ops, pseed=int(rng.integers(0, 99)))


def drive(env, spec, Action, ActionType):
    from gen_cim_golden import policy_random

    log = []
    B = spec["batch"]
    metrics, decisions, done = env.step(None)
    cur = list(decisions)  # last decision per env
    steps = [0] * B

    def describe(ms, ds, dn):
        return [[None if m is None else [int(m["order_requirements"]), int(m["container_shortage"]), int(m["operation_number"])] for m in ms],
                [None if d is None else [int(d.tick), int(d.port_idx), int(d.vessel_idx), int(d.action_scope.load),
                                         int(d.action_scope.discharge), int(d.early_discharge)] for d in ds],
                bool(dn), [int(t) for t in env.tick], [int(f) for f in env.frame_index]]

    log.append(describe(metrics, decisions, done))

    def act_for(i):
        d = cur[i]
        if d is None:
            return None
        row = [d.tick, d.port_idx, d.vessel_idx, d.action_scope.load, d.action_scope.discharge, d.early_discharge]
        v, p, q, t = policy_random(row, spec["pseed"], i, steps[i])
        steps[i] += 1
        return Action(v, p, q, ActionType.DISCHARGE if t else ActionType.LOAD)

    for op in spec["ops"]:
        if done:
            break
        if op[0] == "list":
            idx = list(range(B))
            metrics, decisions, done = env.step([act_for(i) for i in idx])
        elif op[0] == "dict":
            idx = op[1]
            metrics, decisions, done = env.step({i: act_for(i) for i in idx})
        else:
            idx = list(range(B))
            for i in idx:
                steps[i] += cur[i] is not None
            metrics, decisions, done = env.step(None)
        for k, i in enumerate(idx):
            cur[i] = decisions[k]
        log.append(describe(metrics, decisions, done))
    q = env.snapshot_list["ports"][0:[0, 1]:["empty", "booking"]]
    log.append([np.asarray(x, np.float64).tolist() for x in q])
    return log


def ref_run(spec, q):
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(1, os.path.join(ROOT, "oracle", "_ref", "_stubs"))
    from maro.simulator.scenarios.cim.common import Action, ActionType
    from maro.vector_env import VectorEnv

    with VectorEnv(batch_num=spec["batch"], scenario="cim", topology=spec["topology"], durations=spec["durations"]) as env:
        q.put(json.dumps(drive(env, spec, Action, ActionType)))


def ours(spec):
    import maro_b200.vector_env.vector_env as venv_mod
    from emul_batch import EmulCimBatch
    from maro_b200.scenarios.cim.common import Action, ActionType

    venv_mod.CimBatch = EmulCimBatch
    with venv_mod.VectorEnv(batch_num=spec["batch"], scenario="cim", topology=spec["topology"], durations=spec["durations"]) as env:
        return json.dumps(drive(env, spec, Action, ActionType))


def main():
    n, first = (int(sys.argv[1]) if len(sys.argv) > 1 else 8), (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = mp.get_context("spawn")
    bad = 0
    for seed in range(first, first + n):
        spec = script_of(seed)
        q = ctx.Queue()
        p = ctx.Process(target=ref_run, args=(spec, q))
        p.start()
        ref = q.get()
        p.join()
        got = ours(spec)
        if ref != got:
            bad += 1
            a, b = json.loads(ref), json.loads(got)
            k = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), min(len(a), len(b)))
            print(seed, "MISMATCH at call", k, "op", (spec["ops"][k - 1] if 0 < k <= len(spec["ops"]) else None),
                  "\n  ref ", a[k] if k < len(a) else None, "\n  ours", b[k] if k < len(b) else None, flush=True)
        else:
            print(seed, "ok", spec["topology"], spec["batch"], len(json.loads(ref)), "calls", flush=True)
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
