"""Generate golden vm_scheduling traces from the UNMODIFIED reference (oracle/_ref).

    bash oracle/build_ref.sh && python tests/golden/gen_vm_golden.py

Datasets: the reference's own toy fixture (tests/data/vm_scheduling/azure.2019.toy -> tests/golden/vm_toy) and the
synthetic trace of tests/golden/vm_synth_gen.py under several configs (tight capacity with a buffer budget -> resource
postponements and failures; oversubscription -> overloads and kills).  Agents: first valid PM, best fit
(examples/vm_scheduling/rule_based_algorithm/best_fit.py, metric remaining_cpu_cores), and a mixed agent that also
postpones / sends no action.  Output: tests/golden/vm_<case>.npz.
"""
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))

BASE = dict(BUFFER_TIME_BUDGET=0, DELAY_DURATION=1, TICKS_PER_HOUR=12, KILL_ALL_VMS_IF_OVERLOAD=True,
            MAX_CPU_OVERSUBSCRIPTION_RATE=1.15, MAX_MEM_OVERSUBSCRIPTION_RATE=1, MAX_UTILIZATION_RATE=1,
            PRICE_PER_CPU_CORES_PER_HOUR=0.0698, PRICE_PER_MEMORY_PER_HOUR=0.0078, UNIT_ENERGY_PRICE_PER_KWH=0.07,
            POWER_USAGE_EFFICIENCY=1.7, PROCESSED_DATA_URL="")


def config(data, pms, pm_per_rack, racks, **over):
    """one region / zone / data centre / cluster type; `pms` = list of (cpu, memory, busy, idle) PM types"""
    files = {"vm_toy": ("vmtable_toy.bin", "vm_cpu_readings-file-1-of-toy.bin"),
             "vm_synth": ("vmtable_synth.bin", "vm_cpu_readings-file-1-of-synth.bin")}[data]
    conf = dict(BASE)
    conf.update(over)
    conf["VM_TABLE"] = os.path.join(HERE, data, files[0])
    conf["CPU_READINGS"] = os.path.join(HERE, data, files[1])
    conf["components"] = {
        "pm": [{"pm_type": i, "cpu": c, "memory": m, "power_curve": {"calibration_parameter": 1.4, "busy_power": b, "idle_power": d}}
               for i, (c, m, b, d) in enumerate(pms)],
        "rack": [{"type": "a", "pm": [{"pm_type": i, "pm_amount": pm_per_rack} for i in range(len(pms))]}],
        "cluster": [{"type": "C1", "rack": [{"rack_type": "a", "rack_amount": racks}]}],
    }
    conf["architecture"] = {"region": [{"name": "R", "zone": [{"name": "Z", "data_center": [
        {"name": "D", "cluster": [{"type": "C1", "cluster_amount": 1}]}]}]}]}
    return conf


def config_multi(data, **over):
    """two regions / three zones / four data centres, two cluster types, two rack types, two PM types (the shape of the
    reference's own test config tests/data/vm_scheduling/config.yml, scaled down)"""
    conf = config(data, [(32, 128, 185, 120), (16, 112, 100, 60)], 1, 1, **over)
    conf["components"]["rack"] = [{"type": "a", "pm": [{"pm_type": 0, "pm_amount": 2}, {"pm_type": 1, "pm_amount": 1}]},
                                  {"type": "b", "pm": [{"pm_type": 1, "pm_amount": 2}]}]
    conf["components"]["cluster"] = [{"type": "C1", "rack": [{"rack_type": "a", "rack_amount": 2}, {"rack_type": "b", "rack_amount": 1}]},
                                     {"type": "C2", "rack": [{"rack_type": "b", "rack_amount": 2}]}]
    conf["architecture"] = {"region": [
        {"name": "R1", "zone": [{"name": "Z1", "data_center": [{"name": "D1", "cluster": [{"type": "C1", "cluster_amount": 1}]},
                                                                {"name": "D2", "cluster": [{"type": "C2", "cluster_amount": 2}]}]},
                                {"name": "Z2", "data_center": [{"name": "D3", "cluster": [{"type": "C1", "cluster_amount": 1}]}]}]},
        {"name": "R2", "zone": [{"name": "Z3", "data_center": [{"name": "D4", "cluster": [{"type": "C2", "cluster_amount": 1},
                                                                                          {"type": "C1", "cluster_amount": 1}]}]}]}]}
    return conf


CASES = {
    "toy_5_first": dict(conf=config("vm_toy", [(32, 128, 185, 120)], 10, 10, MAX_CPU_OVERSUBSCRIPTION_RATE=1), durations=5, agent="first"),
    "synth_160_bestfit": dict(conf=config("vm_synth", [(32, 128, 185, 120), (16, 112, 100, 60)], 3, 2), durations=160, agent="best"),
    "synth_160_tight_budget": dict(conf=config("vm_synth", [(32, 64, 185, 120)], 2, 2, BUFFER_TIME_BUDGET=6, DELAY_DURATION=2),
                                   durations=160, agent="best", snapshot_resolution=4, max_snapshots=16),
    "synth_140_multi_region": dict(conf=config_multi("vm_synth", BUFFER_TIME_BUDGET=3, MAX_CPU_OVERSUBSCRIPTION_RATE=1.5,
                                                      MAX_UTILIZATION_RATE=1.2), durations=140, agent="mixed",
                                   snapshot_resolution=2, max_snapshots=30),
    # start_tick > 0: the VM table and the readings are entered mid-trace (second readings file: the reader switches at init)
    "synth_start90_60_bestfit": dict(conf=config("vm_synth", [(32, 128, 185, 120)], 3, 2, BUFFER_TIME_BUDGET=2), start_tick=90,
                                     durations=60, agent="best", snapshot_resolution=3),
    "synth_120_oversub_mixed": dict(conf=config("vm_synth", [(16, 96, 150, 90)], 4, 2, MAX_CPU_OVERSUBSCRIPTION_RATE=2.5,
                                                 MAX_UTILIZATION_RATE=3, BUFFER_TIME_BUDGET=4), durations=120, agent="mixed"),
}

PM_ATTRS = ("cluster_id", "cpu_cores_allocated", "cpu_cores_capacity", "cpu_utilization", "data_center_id", "energy_consumption",
            "id", "memory_allocated", "memory_capacity", "oversubscribable", "pm_type", "rack_id", "region_id", "zone_id")
UPPER = {"racks": ("cluster_id", "data_center_id", "empty_machine_num", "id", "region_id", "total_machine_num", "zone_id"),
         "clusters": ("data_center_id", "empty_machine_num", "id", "region_id", "total_machine_num", "zone_id"),
         "data_centers": ("empty_machine_num", "id", "region_id", "total_machine_num", "zone_id"),
         "zones": ("empty_machine_num", "id", "region_id", "total_machine_num"),
         "regions": ("empty_machine_num", "id", "total_machine_num")}
METRICS = ("total_vm_requests", "total_incomes", "energy_consumption_cost", "total_profit", "total_energy_consumption",
           "successful_allocation", "successful_completion", "failed_allocation", "failed_completion")


def agent_action(kind, dec, env, step, n_pm):
    """-> (action row [vm_id, kind, arg, 0] or None)"""
    if kind == "first":
        return [dec.vm_id, 0, dec.valid_pms[0], 0]
    info = env.snapshot_list["pms"][env.frame_index:dec.valid_pms:["cpu_cores_capacity", "cpu_cores_allocated"]].reshape(-1, 2)
    best = dec.valid_pms[int(np.argmin(info[:, 0] - info[:, 1]))]
    if kind == "best":
        return [dec.vm_id, 0, best, 0]
    r = (step * 2654435761 + dec.vm_id * 40503) % 17
    if r == 0:
        return None  # empty action list
    if r in (1, 2):
        return [dec.vm_id, 1, 1 + r % 2, 0]  # PostponeAction
    return [dec.vm_id, 0, dec.valid_pms[(step + dec.vm_id) % len(dec.valid_pms)], 0]


def run_case(name, spec, out_dir=None):
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(1, os.path.join(ROOT, "oracle", "_ref", "_stubs"))
    import yaml
    from maro.simulator import Env
    from maro.simulator.scenarios.vm_scheduling import AllocateAction, PostponeAction

    d = tempfile.mkdtemp()
    with open(os.path.join(d, "config.yml"), "w") as fp:
        yaml.safe_dump(spec["conf"], fp, sort_keys=False)
    env = Env("vm_scheduling", d, start_tick=spec.get("start_tick", 0), durations=spec["durations"],
              snapshot_resolution=spec.get("snapshot_resolution", 1),
              max_snapshots=spec.get("max_snapshots"))
    n_pm = len(env.snapshot_list["pms"])
    rows, valid, mets, acts = [], [], [], []
    metrics, dec, done = env.step(None)
    step = 0
    while not done:
        rows.append([env.tick, dec.vm_id, dec.frame_index, dec.vm_cpu_cores_requirement, dec.vm_memory_requirement, dec.vm_sub_id,
                     int(dec.vm_category), dec.remaining_buffer_time, len(dec.valid_pms)])
        v = np.full(n_pm, -1, np.int32)
        v[:len(dec.valid_pms)] = dec.valid_pms
        valid.append(v)
        mets.append([float(metrics[k]) for k in METRICS] + [metrics["total_latency"].due_to_agent, metrics["total_latency"].due_to_resource,
                                                              metrics["total_oversubscriptions"], metrics["total_overload_pms"],
                                                              metrics["total_overload_vms"]])
        a = agent_action(spec["agent"], dec, env, step, n_pm)
        acts.append(a if a is not None else [-1, -1, -1, -1])
        if a is None:
            action = None
        elif a[1] == 0:
            action = AllocateAction(vm_id=a[0], pm_id=a[2])
        else:
            action = PostponeAction(vm_id=a[0], postpone_step=a[2])
        step += 1
        metrics, dec, done = env.step(action)
    final = [float(metrics[k]) for k in METRICS] + [metrics["total_latency"].due_to_agent, metrics["total_latency"].due_to_resource,
                                                     metrics["total_oversubscriptions"], metrics["total_overload_pms"],
                                                     metrics["total_overload_vms"]]
    sl = env.snapshot_list
    frames = sorted(sl.get_frame_index_list())
    out = {"steps": np.asarray(rows, np.int64).reshape(-1, 9), "valid": np.asarray(valid, np.int32).reshape(-1, n_pm),
           "metrics": np.asarray(mets, np.float64).reshape(-1, 14), "actions": np.asarray(acts, np.int32).reshape(-1, 4),
           "final_metrics": np.asarray(final, np.float64), "final_tick": np.asarray(env.tick), "frames": np.asarray(frames, np.int32)}
    for a in PM_ATTRS:
        x = sl["pms"][frames::a].reshape(len(frames), n_pm)
        out["pms/" + a] = x.astype(np.float32 if a in ("cpu_utilization", "energy_consumption") else np.int32)
    for node, attrs in UPPER.items():
        n = len(sl[node])
        for a in attrs:
            out[f"{node}/{a}"] = sl[node][frames::a].reshape(len(frames), n).astype(np.int32)
    np.savez_compressed(os.path.join(out_dir or HERE, f"vm_{name}.npz"), **out)
    print(name, "steps", len(rows), "final", [round(x, 4) for x in final], "tick", env.tick, flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    ctx = mp.get_context("spawn")
    for n in names:
        p = ctx.Process(target=run_case, args=(n, CASES[n]))
        p.start()
        p.join()
        if p.exitcode != 0:
            raise SystemExit(f"case {n} failed")
