"""Randomised check of the drop-in Python façade against the reference's own Env: ``maro_b200.simulator.Env`` (driven
through the emulator-backed batch of tests/emul_batch.py, i.e. the real host code + the kernel logic) vs
``maro.simulator.Env`` from oracle/_ref, property by property: tick, frame_index, snapshot_list frame indices and
lengths, get_ticks_frame_index_mapping, queries with explicit / empty tick and node lists, metrics, agent_idx_list,
summary node mapping.  Build-container tool (needs oracle/_ref).

    python tools/fuzz_facade_vs_reference.py [n_cases] [first_seed]
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOPOLOGIES = ["toy.4p_ssdd_l0.0", "toy.4p_ssdd_l0.6", "toy.5p_ssddd_l0.2", "toy.6p_sssbdd_l0.4", "global_trade.22p_l0.1"]


def spec_of(seed):
    rng = np.random.default_rng(seed)
    kind = ["cim", "cim", "citi_bike", "vm_scheduling"][int(rng.integers(0, 4))]
    if kind != "cim":
        return dict(scenario=kind, durations=int(rng.integers(60, 500) if kind == "citi_bike" else rng.integers(30, 150)),
                    snapshot_resolution=int(rng.choice([1, 2, 5, 10])),
                    max_snapshots=(None if rng.random() < 0.5 else int(rng.integers(2, 25))),
                    stop_after=(None if rng.random() < 0.6 else int(rng.integers(1, 12))))
    t = str(rng.choice(TOPOLOGIES))
    return dict(scenario="cim", topology=t, durations=int(rng.integers(20, 45 if t.startswith("global") else 130)),
                snapshot_resolution=int(rng.choice([1, 1, 2, 3, 7])),
                max_snapshots=(None if rng.random() < 0.5 else int(rng.integers(2, 25))),
                stop_after=(None if rng.random() < 0.6 else int(rng.integers(1, 12))))


def topology_dir(spec):
    """config folder of the non-CIM scenarios (the committed test fixtures)"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    if spec["scenario"] == "citi_bike":
        import gen_bike_golden

        return gen_bike_golden.data_config_dir("bike_toy")
    import tempfile

    import gen_vm_golden
    import yaml

    d = tempfile.mkdtemp()
    with open(os.path.join(d, "config.yml"), "w") as fp:
        yaml.safe_dump(gen_vm_golden.CASES["synth_160_bestfit"]["conf"], fp, sort_keys=False)
    return d


def first_action(spec, dec, mod):
    """answer every decision with the simplest valid action of the scenario (None for citi_bike / cim)"""
    if spec["scenario"] == "vm_scheduling" and dec is not None:
        return mod.AllocateAction(vm_id=dec.vm_id, pm_id=dec.valid_pms[0])
    return None


def observe(env, spec, mod=None):
    """run (optionally stop early) and collect every observable of the façade"""
    out = {}
    metrics, dec, done = env.step(None)
    n = 0
    while not done and (spec["stop_after"] is None or n < spec["stop_after"]):
        metrics, dec, done = env.step(first_action(spec, dec, mod))
        n += 1
    if spec["scenario"] != "cim":
        sl = env.snapshot_list
        frames = [int(f) for f in sl.get_frame_index_list()]
        out["tick"], out["frame_index"], out["done"], out["steps"] = int(env.tick), int(env.frame_index), bool(done), n
        out["frames"], out["len"] = frames, len(sl)
        out["mapping"] = sorted((int(k), int(v)) for k, v in env.get_ticks_frame_index_mapping().items())
        out["agents"] = None if env.agent_idx_list is None else [int(a) for a in env.agent_idx_list]
        if spec["scenario"] == "citi_bike":
            out["metrics"] = {k: int(v) for k, v in dict(metrics).items()}
            qs = [sl["stations"][frames[-2:]::["bikes", "shortage", "trip_requirement"]], sl["stations"][:1:"fulfillment"],
                  sl["matrices"][frames[-1]::"trips_adj"], sl["stations"][[10 ** 6]:0:"bikes"]]
            out["n_nodes"] = [len(sl["stations"])]
        else:
            out["metrics"] = {k: (round(float(v), 9) if isinstance(v, float) or hasattr(v, "dtype") else
                                  ([int(v.due_to_agent), int(v.due_to_resource)] if hasattr(v, "due_to_agent") else int(v)))
                              for k, v in dict(metrics).items()}
            qs = [sl["pms"][frames[-2:]::["cpu_cores_allocated", "cpu_utilization", "energy_consumption"]],
                  sl["racks"][:0:"empty_machine_num"], sl["regions"][frames[-1]::["total_machine_num", "empty_machine_num"]],
                  sl["pms"][[10 ** 6]:0:"cpu_cores_capacity"]]
            out["n_nodes"] = [len(sl[k]) for k in ("pms", "racks", "clusters", "data_centers", "zones", "regions")]
        out["queries"] = [np.asarray(q, np.float64).tolist() for q in qs]  # exact float64 values
        return out
    sl = env.snapshot_list
    frames = [int(f) for f in sl.get_frame_index_list()]
    out["tick"], out["frame_index"], out["done"], out["steps"] = int(env.tick), int(env.frame_index), bool(done), n
    out["frames"], out["len"] = frames, len(sl)
    out["mapping"] = sorted((int(k), int(v)) for k, v in env.get_ticks_frame_index_mapping().items())
    out["metrics"] = {k: int(v) for k, v in dict(metrics).items()}
    out["agents"] = [int(a) for a in env.agent_idx_list]
    out["n_nodes"] = [len(sl["ports"]), len(sl["vessels"])]
    q1 = sl["ports"][frames[-3:]:[0, 1]:["empty", "full", "acc_shortage"]]
    q2 = sl["ports"][:0:"booking"]                 # all frames, node 0
    q3 = sl["vessels"][frames[-1]::["remaining_space", "future_stop_list"]]  # all nodes
    q4 = sl["matrices"][[frames[0], 10 ** 6]::"vessel_plans"]  # unknown frame -> zeros
    out["queries"] = [np.asarray(q, np.float64).round(6).tolist() for q in (q1, q2, q3, q4)]
    out["node_mapping"] = env.summary["node_mapping"]
    return out


def ref_run(spec, q):
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(1, os.path.join(ROOT, "oracle", "_ref", "_stubs"))
    from maro.simulator import Env

    mod = None
    if spec["scenario"] == "citi_bike":
        np.random.seed(5)
    if spec["scenario"] == "vm_scheduling":
        import maro.simulator.scenarios.vm_scheduling as mod
    env = Env(spec["scenario"], spec.get("topology") or topology_dir(spec), durations=spec["durations"],
              snapshot_resolution=spec["snapshot_resolution"], max_snapshots=spec["max_snapshots"])
    q.put(json.dumps(observe(env, spec, mod), default=int, sort_keys=True))


def ours(spec):
    import maro_b200.scenarios.vm_scheduling as vm_mod
    import maro_b200.simulator.env as env_mod
    from emul_batch import EmulBikeBatch, EmulCimBatch, EmulVmBatch

    env_mod.CimBatch, env_mod.BikeBatch, env_mod.VmBatch = EmulCimBatch, EmulBikeBatch, EmulVmBatch
    env = env_mod.Env(spec["scenario"], spec.get("topology") or topology_dir(spec), durations=spec["durations"],
                      snapshot_resolution=spec["snapshot_resolution"], max_snapshots=spec["max_snapshots"],
                      options={"transfer_seed": 5})
    return json.dumps(observe(env, spec, vm_mod), default=int, sort_keys=True)


def main():
    n, first = (int(sys.argv[1]) if len(sys.argv) > 1 else 12), (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    ctx = mp.get_context("spawn")
    bad = 0
    for seed in range(first, first + n):
        spec = spec_of(seed)
        q = ctx.Queue()
        p = ctx.Process(target=ref_run, args=(spec, q))
        p.start()
        try:
            ref = q.get(timeout=180)
        except Exception:
            p.terminate()
            print(seed, "reference failed (skipped)", spec, flush=True)
            continue
        p.join()
        got = ours(spec)
        if ref != got:
            bad += 1
            a, b = json.loads(ref), json.loads(got)
            diff = [k for k in a if a[k] != b.get(k)]
            print(seed, "MISMATCH", spec, diff, {k: (a[k], b.get(k)) for k in diff[:2]}, flush=True)
        else:
            print(seed, "ok", spec, flush=True)
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
