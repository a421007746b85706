"""CPU: the citi_bike C oracle (oracle/bike_oracle.c) against traces of the unmodified reference."""
import numpy as np
import pytest

from bike_helpers import BIKE_CASES, assert_bike_snapshots_equal, bike_topology, drive_bike, load_bike_golden
from oracle.bike_oracle import BikeOracle, policy_greedy
from bike_helpers import greedy_py


@pytest.mark.parametrize("name", sorted(BIKE_CASES))
def test_bike_oracle_matches_reference_trace(name):
    spec = BIKE_CASES[name]
    topo = bike_topology(spec)
    gold = load_bike_golden(name)
    o = BikeOracle(topo, spec["snapshot_resolution"], spec.get("max_snapshots"))
    rows, scopes, final, st, dec = drive_bike(lambda a: o.step(a), spec, topo.n_stations)
    assert rows.shape == gold["steps"].shape
    if not np.array_equal(rows, gold["steps"]):
        bad = np.argwhere(rows != gold["steps"])[0]
        raise AssertionError(f"step {bad[0]} col {bad[1]}: got {rows[bad[0]]} want {gold['steps'][bad[0]]}")
    assert np.array_equal(scopes, gold["scopes"])
    assert final.tolist() == gold["final_metrics"].tolist()
    assert o.tick == int(gold["final_tick"]) and st == 1
    assert o.step(None)[0] == 2
    assert_bike_snapshots_equal(o.snapshot, gold, topo.n_stations)


def test_greedy_policy_c_matches_python():
    rng = np.random.default_rng(1)
    for _ in range(200):
        S = int(rng.integers(2, 6))
        dec = np.zeros(8 + 2 * S, np.int32)
        dec[1] = rng.integers(0, S); dec[3] = rng.integers(0, 2); dec[4] = S
        for k in range(S):
            dec[8 + 2 * k] = k
            dec[9 + 2 * k] = rng.integers(0, 5)
        assert policy_greedy(dec).tolist() == greedy_py(dec)
