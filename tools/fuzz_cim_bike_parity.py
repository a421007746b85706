"""Randomised differential check of the CIM and citi_bike paths: unmodified reference (oracle/_ref, one process per
case) vs the C oracles vs the device code under the host emulator.  CIM: random built-in topology (all noise levels),
topology seed, durations, start tick 0, snapshot resolution / ring, null or hashed-random agent.  citi_bike: the frozen
toy trace with random transfer seeds, start ticks, durations, resolutions, rings, null / greedy agent.
Build-container tool (needs oracle/_ref).

    python tools/fuzz_cim_bike_parity.py [n_cases] [first_seed]
"""
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

CIM_TOPOLOGIES = [f"{base}_l0.{k}" for base in ("toy.4p_ssdd", "toy.5p_ssddd", "toy.6p_sssbdd", "global_trade.22p") for k in range(9)]


def cim_spec(seed):
    rng = np.random.default_rng(seed)
    topo = str(rng.choice(CIM_TOPOLOGIES))
    big = topo.startswith("global")
    spec = dict(topology=topo, durations=int(rng.integers(30, 70 if big else 260)), policy=int(rng.integers(0, 3)),
                pseed=int(rng.integers(0, 1000)), replica=int(rng.integers(0, 64)),
                snapshot_resolution=int(rng.choice([1, 1, 2, 5])))
    if rng.random() < 0.5:
        spec["max_snapshots"] = int(rng.integers(3, 30))
    if rng.random() < 0.4:
        spec["topo_seed"] = int(rng.integers(1, 100000))
    return spec


def bike_spec(seed):
    rng = np.random.default_rng(seed)
    start = int(rng.choice([0, 0, 300, 900]))
    spec = dict(data="bike_toy", start_tick=start, durations=int(rng.integers(200, 1200)), policy=int(rng.integers(0, 2)),
                snapshot_resolution=int(rng.choice([1, 3, 10, 20])), np_seed=int(rng.integers(0, 100000)))
    if rng.random() < 0.5:
        spec["max_snapshots"] = int(rng.integers(3, 40))
    return spec


def bike_filter_spec(seed):
    """the 26-station synthetic dataset with a random action-scope filter chain (filters that DROP neighbours), decision /
    snapshot resolutions that do not divide each other (stale trip-window cache entries), short rings"""
    rng = np.random.default_rng(seed + 77777)
    filters = ['      - type: "distance"\n        num: %d' % int(rng.integers(5, 26))] if rng.random() < 0.7 else []
    for _ in range(int(rng.integers(1, 3))):
        if rng.random() < 0.5:
            filters.append('      - type: "requirements"\n        num: %d' % int(rng.integers(2, 20)))
        else:
            filters.append('      - type: "trip_window"\n        windows: %d\n        num: %d' % (int(rng.integers(1, 9)), int(rng.integers(2, 20))))
    text = ("decision:\n  extra_cost_mode: %s\n  resolution: %d\n  effective_time_mean: %d\n  effective_time_std: %d\n"
            "  supply_water_mark_ratio: %.2f\n  demand_water_mark_ratio: %.2f\n  action_scope:\n    low: %.2f\n    high: %.2f\n    filters:\n%s\n"
            'time_zone: "America/New_York"\n') % (str(rng.choice(["source", "target", "target_neighbors"])), int(rng.choice([5, 12, 20, 30])),
                                                  int(rng.integers(3, 25)), int(rng.integers(1, 6)), float(rng.uniform(0.6, 0.9)),
                                                  float(rng.uniform(0.1, 0.3)), float(rng.uniform(0, 0.3)), float(rng.uniform(0.6, 1.0)),
                                                  "\n".join(filters))
    start = int(rng.choice([0, 0, 200, 600]))
    spec = dict(data="bike_synth26", start_tick=start, durations=int(rng.integers(150, 800)), policy=int(rng.integers(0, 2)),
                snapshot_resolution=int(rng.choice([1, 3, 7, 10, 20])), np_seed=int(rng.integers(0, 100000)), decision_text=text)
    if rng.random() < 0.5:
        spec["max_snapshots"] = int(rng.integers(2, 30))
    return spec


def main():
    import gen_bike_golden as gb
    import gen_cim_golden as gc
    from bike_helpers import assert_bike_snapshots_equal, bike_topology, drive_bike
    from emul import BikeEmulEnv, EmulEnv
    from helpers import assert_snapshots_equal, case_topology, drive
    from oracle.bike_oracle import BikeOracle
    from oracle.cim_oracle import CimOracle

    n, first = (int(sys.argv[1]) if len(sys.argv) > 1 else 16), (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    out = tempfile.mkdtemp()
    ctx = mp.get_context("spawn")
    p0 = ctx.Process(target=gb.prepare_reference_cases)
    p0.start(); p0.join()
    bad = 0
    for seed in range(first, first + n):
        for kind in (("bikef",) if os.environ.get("FUZZ_ONLY") == "bike_filters" else ("cim", "bike", "bikef")):
            spec = cim_spec(seed) if kind == "cim" else (bike_spec(seed) if kind == "bike" else bike_filter_spec(seed))
            name = f"fuzz{seed}{'f' if kind == 'bikef' else ''}"
            p = ctx.Process(target=(gc if kind == "cim" else gb).run_case, args=(name, spec, out))
            p.start(); p.join()
            if p.exitcode != 0:
                print(seed, kind, "reference failed (skipped)", spec)
                continue
            gold = np.load(os.path.join(out, f"{'cim' if kind == 'cim' else 'bike'}_{name}.npz"))
            res, ring = spec.get("snapshot_resolution", 1), spec.get("max_snapshots")
            try:
                if kind == "cim":
                    topo = case_topology(spec)
                    for e, step in ((CimOracle(topo, 0, res, ring), None), (EmulEnv(topo, 1, 0, res, ring), "step1")):
                        fn = (lambda a, e=e: e.step(a)) if step is None else (lambda a, e=e: e.step1(a))
                        rows, final, dec, st = drive(fn, spec)
                        assert np.array_equal(rows, gold["steps"]) and final.tolist() == gold["final_metrics"].tolist() and st == 1
                        assert_snapshots_equal(e.snapshot, gold, topo)
                else:
                    topo = bike_topology(spec)
                    for e, fn in ((lambda: BikeOracle(topo, res, ring), "step"), (lambda: BikeEmulEnv(topo, res, ring), "step1")):
                        env = e()
                        rows, scopes, final, st, dec = drive_bike(getattr(env, fn), spec, topo.n_stations)
                        assert np.array_equal(rows, gold["steps"]) and np.array_equal(scopes, gold["scopes"])
                        assert final.tolist() == gold["final_metrics"].tolist() and st == 1
                        assert_bike_snapshots_equal(env.snapshot, gold, topo.n_stations)
            except AssertionError as ex:
                bad += 1
                print(seed, kind, "MISMATCH", str(ex)[:300], spec)
                continue
            print(seed, kind, "ok", len(gold["steps"]), "steps", {k: v for k, v in spec.items() if k != "decision_text"}, flush=True)
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
