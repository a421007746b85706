"""CPU: the vm_scheduling device code (maro_b200/csrc/vm_core.cuh) under the thread-per-lane emulator, against the
reference traces and the C oracle.  Integers and the energy metrics are bit-exact; total_incomes / total_profit use
a maintained live-price sum and are compared to 1e-9."""
import numpy as np
import pytest

from emul import VmEmulEnv
from oracle.vm_oracle import VmOracle
from vm_helpers import VM_CASES, assert_metrics_close, assert_vm_snapshots_equal, drive_vm, load_vm_golden, vm_topology

EXACT_COLS = [0, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]



def _row(dec, n_valid):
    """header + valid PM ids of a decision row; word 11 (offset of this build's remaining-cores extension) is not part of the
    reference's DecisionEvent and absent from the oracle's rows"""
    return list(dec[:11]) + list(dec[12:12 + n_valid])


def _abi_metrics(row):
    from maro_b200 import _abi

    return _abi.vm_metrics_dict(row)
  # every metric except total_incomes (1) and total_profit (3)


@pytest.mark.parametrize("lanes", [32, 8, 1])
@pytest.mark.parametrize("name", sorted(VM_CASES))
def test_vm_device_logic_matches_reference_trace(name, lanes):
    spec = VM_CASES[name]
    topo = vm_topology(spec)
    gold = load_vm_golden(name)
    res, ms = spec.get("snapshot_resolution", 1), spec.get("max_snapshots")
    e = VmEmulEnv(topo, res, ms, lanes=lanes)
    rows, valid, mets, final, st, dec = drive_vm(lambda a: e.step(a), gold, topo.n_pm)
    assert rows.shape == gold["steps"].shape
    if not np.array_equal(rows, gold["steps"]):
        bad = np.argwhere(rows != gold["steps"])[0]
        raise AssertionError(f"step {bad[0]} col {bad[1]}: got {rows[bad[0]]} want {gold['steps'][bad[0]]}")
    assert np.array_equal(valid, gold["valid"])
    assert_metrics_close(mets, gold["metrics"], "per-step")
    assert_metrics_close(final, gold["final_metrics"], "final")
    assert e.tick() == int(gold["final_tick"]) and st == 1
    assert e.step(None)[0] == 2
    assert_vm_snapshots_equal(e.snapshot, gold, topo)
    # against the oracle: same counters, bit-identical frames and energy metrics
    o = VmOracle(topo, res, ms)
    _, _, omets, ofinal, _, _ = drive_vm(lambda a: o.step(a), gold, topo.n_pm)
    assert np.array_equal(mets[:, EXACT_COLS], omets[:, EXACT_COLS])
    assert np.array_equal(final[EXACT_COLS], ofinal[EXACT_COLS])
    o.step(None)
    assert np.array_equal(e.counters(), o.counters())
    assert np.array_equal(e.frame(), o.frame())
    for f in gold["frames"].tolist():
        assert np.array_equal(e.snapshot(int(f)), o.snapshot(int(f))), f


def test_vm_device_logic_reset_and_bad_action():
    spec = VM_CASES["synth_160_bestfit"]
    topo = vm_topology(spec)
    gold = load_vm_golden("synth_160_bestfit")
    e = VmEmulEnv(topo)
    first = drive_vm(lambda a: e.step(a), gold, topo.n_pm)
    e.reset()
    again = drive_vm(lambda a: e.step(a), gold, topo.n_pm)
    for a, b in zip(first[:4], again[:4]):
        assert np.array_equal(a, b)
    # an action naming a VM that is not the pending decision's -> BAD_ACTION (reference: "The VM id ... is invalid.")
    e.reset()
    st, dec, _ = e.step(None)
    assert st == 0
    st, dec, _ = e.step([[dec[1] + 12345, 0, dec[12], 0]])
    assert st == -1
    assert e.step(None)[0] == 2
    # allocating to a PM id outside the cluster is rejected too
    e.reset()
    st, dec, _ = e.step(None)
    st, _, _ = e.step([[dec[1], 0, topo.n_pm, 0]])
    assert st == -1


@pytest.mark.parametrize("lanes", [32, 8])
def test_vm_device_logic_tiny_and_denormal_utilisations(lanes):
    """30 % of the readings replaced by tiny / denormal float32 values: decisions, frames and energy metrics stay
    bit-identical to the oracle (the float64 utilisation sums run in the reference's list order)."""
    spec = VM_CASES["synth_120_oversub_mixed"]
    topo = vm_topology(spec)
    rng = np.random.default_rng(5)
    v = topo.util_val.copy()
    pick = rng.random(len(v)) < 0.3
    tiny = np.asarray([1e-12, 3e-20, 1e-40, 7.5e-6, 2e-9], np.float32).astype(np.float64)
    v[pick] = tiny[rng.integers(0, len(tiny), int(pick.sum()))]
    topo.util_val = v
    e, o = VmEmulEnv(topo, lanes=lanes), VmOracle(topo)
    (st, dec, met), (ost, odec, omet) = e.step(None), o.step(None)
    n = 0
    while ost == 0:
        assert st == 0 and _row(dec, odec[10]) == _row(odec, odec[10]), n
        assert np.array_equal(met[EXACT_COLS], omet[EXACT_COLS]), n
        a = o.best_fit(odec)
        (st, dec, met), (ost, odec, omet) = e.step(a.reshape(1, 4)), o.step(a.reshape(1, 4))
        n += 1
    assert st == 1 and n > 50
    assert np.array_equal(met[EXACT_COLS], omet[EXACT_COLS])
    assert np.array_equal(e.frame(), o.frame()) and np.array_equal(e.counters(), o.counters())


def test_vm_trace_generator_loads_and_kernel_matches_oracle(tmp_path):
    """tools/vm_trace_gen.py (the synthetic azure-scale trace of BASELINE config #5) at a small size: the .bin files load
    through the host loader, and a best-fit episode of the device logic equals the oracle's."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    import vm_trace_gen

    from maro_b200.scenarios.vm_scheduling.data import build_vm_topology

    vm_path, cpu_path = vm_trace_gen.generate(str(tmp_path), n_vm=400, ticks=300, seed=3, mean_concurrent_cores=250.0)
    topo = build_vm_topology(vm_trace_gen.azure_like_config(vm_path, cpu_path, n_pm=10), 0, 300)
    assert topo.n_vm == 400 and topo.n_pm == 10 and topo.error is None
    e, o = VmEmulEnv(topo, 1, 8), VmOracle(topo, 1, 8)
    (st, dec, met), (ost, odec, omet) = e.step(None), o.step(None)
    n = 0
    while ost == 0:
        assert st == 0 and _row(dec, odec[10]) == _row(odec, odec[10]), n
        a = o.best_fit(odec)
        (st, dec, met), (ost, odec, omet) = e.step(a.reshape(1, 4)), o.step(a.reshape(1, 4))
        n += 1
    assert st == 1 and n >= 300
    assert np.array_equal(met[EXACT_COLS], omet[EXACT_COLS])
    assert np.array_equal(e.frame(), o.frame()) and np.array_equal(e.counters(), o.counters())
    m = _abi_metrics(met)
    assert m["total_vm_requests"] == 400 and m["successful_allocation"] + m["failed_allocation"] == 400


def test_vm_decision_row_extension_holds_remaining_cores():
    """MARO_VM_DEC_EXT_OFFSET: the decision row carries capacity - allocated (cores) of every valid PM in the decision's frame,
    what the reference's rule-based agents fetch with a snapshot query per decision (rule_based_algorithm/best_fit.py:38-44)"""
    from emul_batch import EmulVmBatch
    from maro_b200 import _abi
    from vm_helpers import VM_CASES, vm_topology

    topo = vm_topology(VM_CASES["synth_160_bestfit"])
    b = EmulVmBatch(topo, 1)
    lay, _ = _abi.vm_frame_layout(topo)
    dec, met = b.step(None)
    for k in range(40):
        d = dec[0]
        if d[6] != 0:
            break
        n, ext = int(d[10]), int(d[11])
        f = b.read_frame(0)
        rem = f[lay["pms"]["cpu_cores_capacity"][0]:][:topo.n_pm] - f[lay["pms"]["cpu_cores_allocated"][0]:][:topo.n_pm]
        ids = d[12:12 + n]
        assert ext == 12 + topo.n_pm and d[ext:ext + n].tolist() == rem[ids].tolist(), k
        a = np.zeros((1, 1, 4), np.int32)
        a[0, 0] = [d[1], 0, ids[int(np.argmin(d[ext:ext + n]))], 0]
        dec, met = b.step(a)
    assert k > 20
