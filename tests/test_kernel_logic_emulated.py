"""CPU: the kernel's per-replica logic (maro_b200/csrc/cim_core.cuh compiled for the host, one lane) against the
golden reference traces and against the oracle.  Finds logic bugs before GPU time is spent; the GPU parity tests
(-m gpu) run the real kernels through the C ABI."""
import numpy as np
import pytest

from emul import EmulEnv
from helpers import CASES, assert_snapshots_equal, case_topology, drive, load_golden
from oracle.cim_oracle import CimOracle


def _lane_configs():
    out = []
    for name in sorted(CASES):
        out.append((name, 0))  # the width the library would pick for the topology
        if name in ("toy4p_l00_300_rand_r0", "toy4p_l08_200_rand", "toy6p_l05_120_rand", "gt22p_l08_60_rand"):
            out += [(name, 1), (name, 32)] if not name.startswith("gt22p") else [(name, 8)]
    return out


@pytest.mark.parametrize("name,lanes", _lane_configs())
def test_emulated_kernel_matches_reference_trace(name, lanes):
    """lanes = lanes per replica (0: library default); the cooperative phases must give the same result for any
    group width, including widths smaller than the number of events in a phase (chunk loops)."""
    spec = CASES[name]
    topo = case_topology(spec)
    gold = load_golden(name)
    e = EmulEnv(topo, 1, spec.get("start_tick", 0), spec.get("snapshot_resolution", 1), spec.get("max_snapshots"),
                lanes=lanes)
    rows, final, dec, st = drive(lambda a: e.step1(a), spec)
    assert rows.shape == gold["steps"].shape
    if not np.array_equal(rows, gold["steps"]):
        bad = np.argwhere(rows != gold["steps"])[0]
        raise AssertionError(f"step {bad[0]} col {bad[1]}: got {rows[bad[0]]} want {gold['steps'][bad[0]]}")
    assert final.tolist() == gold["final_metrics"].tolist()
    assert e.tick() == int(gold["final_tick"])
    assert st == 1
    assert e.step1(None)[0] == 2
    if "frames" in gold:
        assert_snapshots_equal(e.snapshot, gold, topo)


def test_emulated_counters_match_oracle():
    spec = CASES["toy4p_l00_300_rand_r0"]
    topo = case_topology(spec)
    e = EmulEnv(topo)
    o = CimOracle(topo)
    drive(lambda a: e.step1(a), spec)
    drive(lambda a: o.step(a), spec)
    assert e.counters().tolist() == o.counters().tolist()
    assert np.array_equal(e.frame(), o.frame())


# ------------------------------------------------------------------------------------------------ citi_bike
from bike_helpers import BIKE_CASES, assert_bike_snapshots_equal, bike_topology, drive_bike, load_bike_golden  # noqa: E402
from emul import BikeEmulEnv  # noqa: E402
from oracle.bike_oracle import BikeOracle  # noqa: E402


@pytest.mark.parametrize("name,lanes", [(n, 0) for n in sorted(BIKE_CASES)] + [("toy_600_greedy_res1", 1), ("toy_600_greedy_res1", 32)])
def test_emulated_bike_kernel_matches_reference_trace(name, lanes):
    spec = BIKE_CASES[name]
    topo = bike_topology(spec)
    gold = load_bike_golden(name)
    e = BikeEmulEnv(topo, spec["snapshot_resolution"], spec.get("max_snapshots"), lanes=lanes)
    rows, scopes, final, st, dec = drive_bike(lambda a: e.step1(a), spec, topo.n_stations)
    assert rows.shape == gold["steps"].shape
    if not np.array_equal(rows, gold["steps"]):
        bad = np.argwhere(rows != gold["steps"])[0]
        raise AssertionError(f"step {bad[0]} col {bad[1]}: got {rows[bad[0]]} want {gold['steps'][bad[0]]}")
    assert np.array_equal(scopes, gold["scopes"])
    assert final.tolist() == gold["final_metrics"].tolist()
    assert e.tick() == int(gold["final_tick"]) and st == 1
    assert e.step1(None)[0] == 2
    assert_bike_snapshots_equal(e.snapshot, gold, topo.n_stations)
    o = BikeOracle(topo, spec["snapshot_resolution"], spec.get("max_snapshots"))
    drive_bike(lambda a: o.step(a), spec, topo.n_stations)
    assert e.counters().tolist() == o.counters().tolist()


def test_emulated_bike_per_replica_transfer_seed():
    """maro_bike_set_transfer_seeds: a replica re-seeded on the device (numpy legacy seeding recurrence in bike_replica_reset)
    follows the reference trace recorded with that np.random seed, whatever the topology's own transfer_seed is."""
    from bike_helpers import bike_config
    from emul import lib as emul_lib
    from maro_b200.scenarios.citi_bike.data import build_bike_topology

    name = "toy_600_greedy_res1"
    spec, gold = BIKE_CASES[name], load_bike_golden(name)
    topo = build_bike_topology(bike_config(spec["data"]), 0, spec["durations"], transfer_seed=spec["np_seed"] + 12345)
    e = BikeEmulEnv(topo, spec["snapshot_resolution"], spec.get("max_snapshots"))
    emul_lib().bike_emul_reseed(e._h, 0, spec["np_seed"])
    rows, scopes, final, st, dec = drive_bike(lambda a: e.step1(a), spec, topo.n_stations)
    assert np.array_equal(rows, gold["steps"]) and final.tolist() == gold["final_metrics"].tolist()
    assert_bike_snapshots_equal(e.snapshot, gold, topo.n_stations)
