"""Generate golden CIM traces from the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh).

Run in the build container only (the reference cannot travel to the GPU box):

    bash oracle/build_ref.sh && python tests/golden/gen_cim_golden.py

One fresh process per case (the reference's SimRandom is process-global, SURVEY.md §8c trap ii).  Each case
drives ``maro.simulator.Env`` with a deterministic action tape (null, or the counter-hash random policy that
oracle/cim_oracle.c and the device policy kernel also implement) and records, per env-step, the decision
payload ints + metrics, and at the end every snapshot still held by the ring for every attribute.
Output: tests/golden/cim_<case>.npz (compressed; integers as int32, transfer_cost as float32).
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))

CASES = {
    # name: dict(topology, durations, policy (0 null / 1 random), seed (policy), replica, env kwargs)
    "toy4p_l00_100_null": dict(topology="toy.4p_ssdd_l0.0", durations=100, policy=0),
    "toy4p_l00_1120_null": dict(topology="toy.4p_ssdd_l0.0", durations=1120, policy=0, snapshots=False),
    "toy4p_l00_300_rand_r0": dict(topology="toy.4p_ssdd_l0.0", durations=300, policy=1, pseed=0, replica=0),
    "toy4p_l00_300_rand_r7": dict(topology="toy.4p_ssdd_l0.0", durations=300, policy=1, pseed=0, replica=7),
    "toy4p_l08_200_rand": dict(topology="toy.4p_ssdd_l0.8", durations=200, policy=1, pseed=3, replica=1),
    "toy4p_l08_200_seed7": dict(topology="toy.4p_ssdd_l0.8", durations=200, policy=1, pseed=3, replica=2, topo_seed=7),
    "toy5p_l03_150_res5_ring10": dict(topology="toy.5p_ssddd_l0.3", durations=150, policy=1, pseed=1, replica=0,
                                      snapshot_resolution=5, max_snapshots=10),
    "toy6p_l05_120_rand": dict(topology="toy.6p_sssbdd_l0.5", durations=120, policy=1, pseed=5, replica=3),
    "gt22p_l08_60_rand": dict(topology="global_trade.22p_l0.8", durations=60, policy=1, pseed=0, replica=0),
    "gt22p_l00_60_null": dict(topology="global_trade.22p_l0.0", durations=60, policy=0),
    # the 22-port noisy topology of the reference's own CIM tests (tests/cim/test_cim_scenario.py:281-324, 391-460)
    "case22p_200_null": dict(topology=os.path.join(ROOT, "tests", "golden", "_case_cfg"), durations=200, policy=0),
    # two episodes: Env.reset(keep_seed=False) draws a new topology seed from the route_init stream
    # (cim_data_container_helpers.py:56-66); the recorded trace is the SECOND episode
    "toy4p_l08_120_reset_newseed": dict(topology="toy.4p_ssdd_l0.8", durations=120, policy=1, pseed=4, replica=0,
                                        reset_new_seed=True),
    "toy4p_l03_start7_90": dict(topology="toy.4p_ssdd_l0.3", durations=90, policy=1, pseed=6, replica=1, start_tick=7,
                                snapshot_resolution=2),
    # BASELINE config #4 at full length: global_trade.22p_l0.8, 500 ticks, topology seed 4096 + 3 (replica 3 of a batch whose
    # replica r runs seed 4096 + r % 8), hashed random agent; every 20th snapshot + the last one are kept
    "gt22p_l08_500_rand_seed4099": dict(topology="global_trade.22p_l0.8", durations=500, policy=1, pseed=0, replica=3,
                                        topo_seed=4099, keep_frames=20),
    "toy4p_l00_start5": dict(topology="toy.4p_ssdd_l0.0", durations=60, policy=1, pseed=2, replica=0, start_tick=0,
                             snapshot_resolution=3),
}

PORT_ATTRS = ("acc_booking", "acc_fulfillment", "acc_shortage", "booking", "capacity", "empty", "fulfillment",
              "full", "on_consignee", "on_shipper", "shortage", "transfer_cost")
VESSEL_ATTRS = ("capacity", "early_discharge", "empty", "full", "is_parking", "last_loc_idx", "loc_port_idx",
                "next_loc_idx", "remaining_space", "route_idx", "past_stop_list", "past_stop_tick_list",
                "future_stop_list", "future_stop_tick_list")
MATRIX_ATTRS = ("full_on_ports", "full_on_vessels", "vessel_plans")


def hash_u32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def policy_random(dec, seed, replica, step):
    """Same arithmetic as cim_policy_random in oracle/cim_oracle.c."""
    h1 = hash_u32(seed ^ hash_u32((replica * 0x9E3779B9 + step * 0x85EBCA6B + 0x1234567) & 0xFFFFFFFF))
    h2 = hash_u32((h1 + 0x68BC21EB) & 0xFFFFFFFF)
    load, dis = dec[3], dec[4]
    to_discharge = dis > 0 and (h1 & 1)
    scope = dis if to_discharge else load
    qty = h2 % (scope + 1) if scope > 0 else 0
    return dec[2], dec[1], int(qty), 1 if to_discharge else 0


def policy_pair(d, spec, step):
    """policy 2: a LOAD followed by a DISCHARGE in one step (both inside the decision's scope: the load only adds empties
    to the vessel, so the discharge bound of the original scope still holds) — exercises action lists."""
    h = hash_u32(spec.get("pseed", 0) * 0x9E3779B9 + step * 0x85EBCA6B + 77)
    load = h % (d[3] + 1) if d[3] > 0 else 0
    dis = (h >> 11) % (d[4] + 1) if d[4] > 0 else 0
    return [(d[2], d[1], int(load), 0), (d[2], d[1], int(dis), 1)]


def run_case(name, spec, out_dir=None):
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(1, os.path.join(ROOT, "oracle", "_ref", "_stubs"))
    from maro.simulator import Env
    from maro.simulator.scenarios.cim.common import Action, ActionType

    topo = spec["topology"]
    if topo.endswith("_case_cfg"):
        # materialise the test config (kept as JSON in tests/golden) as a config.yml folder for the reference
        import json
        import tempfile

        import yaml

        d = tempfile.mkdtemp()
        with open(os.path.join(HERE, "cim_case_config.json")) as fp:
            conf = json.load(fp)
        with open(os.path.join(d, "config.yml"), "w") as fp:
            yaml.safe_dump(conf, fp, sort_keys=False)
        topo = d
    env = Env("cim", topo, start_tick=spec.get("start_tick", 0), durations=spec["durations"],
              snapshot_resolution=spec.get("snapshot_resolution", 1), max_snapshots=spec.get("max_snapshots"))
    if "topo_seed" in spec:
        env.set_seed(spec["topo_seed"])
        env.reset(keep_seed=True)
    if spec.get("reset_new_seed"):
        metrics, dec, done = env.step(None)
        while not done:
            metrics, dec, done = env.step(None)
        env.reset(keep_seed=False)
    rows = []
    step = 0
    metrics, dec, done = env.step(None)
    while not done:
        d = [dec.tick, dec.port_idx, dec.vessel_idx, dec.action_scope.load, dec.action_scope.discharge,
             dec.early_discharge]
        rows.append(d + [int(metrics["order_requirements"]), int(metrics["container_shortage"]),
                         int(metrics["operation_number"])])
        if spec["policy"] == 1:
            v, p, q, t = policy_random(d, spec.get("pseed", 0), spec.get("replica", 0), step)
            action = Action(v, p, q, ActionType.DISCHARGE if t else ActionType.LOAD)
        elif spec["policy"] == 2:
            action = [Action(v, p, q, ActionType.DISCHARGE if t else ActionType.LOAD) for v, p, q, t in policy_pair(d, spec, step)]
            if step % 7 == 3:
                action = None
        else:
            action = None
        step += 1
        metrics, dec, done = env.step(action)
    out = {
        "steps": np.asarray(rows, np.int64).reshape(-1, 9),
        "final_metrics": np.asarray([int(metrics["order_requirements"]), int(metrics["container_shortage"]),
                                     int(metrics["operation_number"])], np.int64),
        "final_tick": np.asarray(env.tick),
    }
    if spec.get("snapshots", True):
        sl = env.snapshot_list
        frames = sorted(sl.get_frame_index_list())
        if spec.get("keep_frames"):
            frames = [f for f in frames if f % spec["keep_frames"] == 0 or f == frames[-1]]
        out["frames"] = np.asarray(frames, np.int32)
        nf = len(frames)
        P, V = len(sl["ports"]), len(sl["vessels"])
        for a in PORT_ATTRS:
            x = sl["ports"][frames::a].reshape(nf, P)
            out["ports/" + a] = x.astype(np.float32 if a == "transfer_cost" else np.int32)
        for a in VESSEL_ATTRS:
            x = sl["vessels"][frames::a].reshape(nf, V, -1)
            out["vessels/" + a] = x.astype(np.int32)
        for a in MATRIX_ATTRS:
            x = sl["matrices"][frames::a].reshape(nf, -1)
            out["matrices/" + a] = x.astype(np.int32)
    np.savez_compressed(os.path.join(out_dir or HERE, f"cim_{name}.npz"), **out)
    print(name, "steps", len(rows), "final", out["final_metrics"].tolist(), "tick", env.tick, flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    ctx = mp.get_context("spawn")
    for n in names:
        p = ctx.Process(target=run_case, args=(n, CASES[n]))
        p.start()
        p.join()
        if p.exitcode != 0:
            raise SystemExit(f"case {n} failed")
