"""CPU: the C oracle (oracle/cim_oracle.c) against golden traces of the unmodified reference, and against the
known answers in the reference's own tests/docs.  This is what makes the oracle "pinned"."""
import numpy as np
import pytest

from helpers import CASES, assert_snapshots_equal, case_topology, drive, load_golden
from oracle.cim_oracle import CimOracle, policy_random
from helpers import policy_random_py


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_trace(name):
    spec = CASES[name]
    topo = case_topology(spec)
    gold = load_golden(name)
    o = CimOracle(topo, spec.get("start_tick", 0), spec.get("snapshot_resolution", 1), spec.get("max_snapshots"))
    rows, final, dec, st = drive(lambda a: o.step(a), spec)
    assert rows.shape == gold["steps"].shape
    if not np.array_equal(rows, gold["steps"]):
        bad = np.argwhere(rows != gold["steps"])[0]
        raise AssertionError(f"step {bad[0]} col {bad[1]}: got {rows[bad[0]]} want {gold['steps'][bad[0]]}")
    assert final.tolist() == gold["final_metrics"].tolist()
    assert o.tick == int(gold["final_tick"])
    assert st == 1
    # stepping a finished env -> (None, None, True)  (core.py:128-131)
    assert o.step(None)[0] == 2
    if "frames" in gold:
        assert_snapshots_equal(o.snapshot, gold, topo)


def test_docs_anchor_null_policy_1120():
    # docs/source/scenarios/container_inventory_management.rst:148-165
    from maro_b200.scenarios.cim.topology import build_topology

    o = CimOracle(build_topology("toy.4p_ssdd_l0.0", 1120))
    n, met = o.run_episode(0)
    assert met.tolist() == [2240000, 2190000, 0]
    assert n == 796


def test_reference_test_order_state_known_answers():
    """tests/cim/test_cim_scenario.py:281-324 — (booking, shortage, empty) of the 22 ports after the first step."""
    spec = CASES["case22p_200_null"]
    topo = case_topology(spec)
    o = CimOracle(topo)
    st, dec, met = o.step(None)
    # first decision: tick 5, vessel 35 (test_vessel_movement :206-214)
    assert (dec[0], dec[2]) == (5, 35)
    # action scope of the first decision (test_early_discharge :404-407)
    assert (dec[3], dec[4], dec[5]) == (1240, 0, 0)
    # (booking, shortage, empty) per port after the first step — tests/cim/test_cim_scenario.py:297-320
    truth = [[223, 0, 14726], [16, 0, 916], [18, 0, 917], [89, 0, 5516], [84, 0, 4613], [72, 0, 4603],
             [26, 0, 1374], [24, 0, 1378], [48, 0, 2756], [54, 0, 2760], [26, 0, 1379], [99, 0, 5534],
             [137, 0, 7340], [19, 0, 912], [13, 0, 925], [107, 0, 6429], [136, 0, 9164], [64, 0, 3680],
             [24, 0, 1377], [31, 0, 1840], [109, 0, 6454], [131, 0, 7351]]
    from helpers import named_frames
    fr = named_frames([o.frame()], topo)
    got = np.stack([fr["ports/booking"][0], fr["ports/shortage"][0], fr["ports/empty"][0]], 1)
    assert got.tolist() == truth
    st, dec, met = o.step(None)
    assert (dec[0], dec[2]) == (6, 27)


def test_reference_test_early_discharge_known_answers():
    """tests/cim/test_cim_scenario.py:391-435 — (full, empty, early_discharge) of vessel 35 at its next decisions."""
    from helpers import named_frames
    spec = CASES["case22p_200_null"]
    topo = case_topology(spec)
    o = CimOracle(topo)
    st, dec, met = o.step(None)
    v, p = int(dec[2]), int(dec[1])
    st, dec, met = o.step(np.asarray([[v, p, 1201, 0], [v, p, 1, 1]], np.int32))
    history = []
    while st == 0:
        st, dec, met = o.step(None)
        if st == 0 and dec[2] == 35:
            fr = named_frames([o.frame()], topo)
            history.append((int(fr["vessels/full"][0, 35, 0]), int(fr["vessels/empty"][0, 35, 0]),
                            int(fr["vessels/early_discharge"][0, 35, 0])))
    assert history == [(465, 838, 362), (756, 547, 291), (1261, 42, 505), (1303, 0, 42), (1303, 0, 0),
                       (1303, 0, 0), (803, 0, 0)]


def test_policy_hash_matches_python():
    rng = np.random.default_rng(0)
    for _ in range(200):
        dec = rng.integers(0, 5000, 8).astype(np.int32)
        seed, rep, step = (int(x) for x in rng.integers(0, 2**31, 3))
        a = policy_random(dec, seed, rep, step)
        b = policy_random_py([int(x) for x in dec[:6]], seed, rep, step)
        assert a.tolist() == list(b)


def test_py_sum_restatement_matches_cpython_sum():
    """list_sum_normalize calls builtin sum(); since CPython 3.12 that is Neumaier-compensated (bltinmodule.c).  The
    restatement used by the oracle / host tables / device code must agree bit for bit with this interpreter."""
    import ctypes as C
    import random
    import sys

    from oracle import cim_oracle

    assert sys.version_info >= (3, 12), "golden traces were recorded on python >= 3.12 (compensated sum)"
    L = cim_oracle.lib()
    L.cim_oracle_py_sum.restype = C.c_double
    L.cim_oracle_py_sum.argtypes = [C.c_void_p, C.c_int]
    rnd = random.Random(1)
    for trial in range(2000):
        n = rnd.randint(1, 40)
        xs = [rnd.uniform(0, 1) * 10 ** rnd.randint(-6, 3) for _ in range(n)]
        a = np.asarray(xs, np.float64)
        assert L.cim_oracle_py_sum(a.ctypes.data, n) == sum(xs), (trial, xs)
