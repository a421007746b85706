"""CPU: the vm_scheduling C oracle (oracle/vm_oracle.c) against traces of the unmodified reference."""
import numpy as np
import pytest

from oracle.vm_oracle import VmOracle
from vm_helpers import VM_CASES, assert_metrics_close, assert_vm_snapshots_equal, drive_vm, load_vm_golden, vm_topology


@pytest.mark.parametrize("name", sorted(VM_CASES))
def test_vm_oracle_matches_reference_trace(name):
    spec = VM_CASES[name]
    topo = vm_topology(spec)
    gold = load_vm_golden(name)
    o = VmOracle(topo, spec.get("snapshot_resolution", 1), spec.get("max_snapshots"))
    rows, valid, mets, final, st, dec = drive_vm(lambda a: o.step(a), gold, topo.n_pm)
    assert rows.shape == gold["steps"].shape
    if not np.array_equal(rows, gold["steps"]):
        bad = np.argwhere(rows != gold["steps"])[0]
        raise AssertionError(f"step {bad[0]} col {bad[1]}: got {rows[bad[0]]} want {gold['steps'][bad[0]]}")
    assert np.array_equal(valid, gold["valid"])
    assert_metrics_close(mets, gold["metrics"], "per-step")
    assert_metrics_close(final, gold["final_metrics"], "final")
    assert o.tick == int(gold["final_tick"]) and st == 1
    assert o.step(None)[0] == 2
    assert_vm_snapshots_equal(o.snapshot, gold, topo)


def test_vm_reference_known_answers():
    """tests/vm_scheduling/test_vm_scheduling_scenario.py:112-136 (hierarchy counts of the test config) and :155-171
    (price model of azure.2019.toy with the first-valid-PM agent)."""
    import yaml

    from maro_b200.scenarios.vm_scheduling.data import build_vm_topology

    conf = yaml.safe_load(CONFIG_1130)
    conf["VM_TABLE"] = VM_CASES["toy_5_first"]["conf"]["VM_TABLE"]
    conf["CPU_READINGS"] = VM_CASES["toy_5_first"]["conf"]["CPU_READINGS"]
    t = build_vm_topology(conf, 0, 3)
    assert (t.n_region, t.n_zone, t.n_dc, t.n_cluster, t.n_rack, t.n_pm) == (2, 2, 3, 8, 75, 1130)
    gold = load_vm_golden("toy_5_first")
    fm = gold["final_metrics"]
    assert abs(fm[1] - 0.185) < 0.01 and abs(fm[2] - 0.595) < 0.01 and abs(fm[3] + 0.410) < 0.01


CONFIG_1130 = """
BUFFER_TIME_BUDGET: 0
DELAY_DURATION: 1
TICKS_PER_HOUR: 12
KILL_ALL_VMS_IF_OVERLOAD: True
MAX_CPU_OVERSUBSCRIPTION_RATE: 1.15
MAX_MEM_OVERSUBSCRIPTION_RATE: 1
MAX_UTILIZATION_RATE: 1
PRICE_PER_CPU_CORES_PER_HOUR: 0.0698
PRICE_PER_MEMORY_PER_HOUR: 0.0078
UNIT_ENERGY_PRICE_PER_KWH: 0.07
POWER_USAGE_EFFICIENCY: 1.7
components:
  pm:
    - {pm_type: 0, cpu: 32, memory: 128, power_curve: {calibration_parameter: 1.4, busy_power: 10, idle_power: 1}}
    - {pm_type: 1, cpu: 16, memory: 112, power_curve: {calibration_parameter: 1.4, busy_power: 10, idle_power: 1}}
  rack:
    - {type: 'a', pm: [{pm_type: 0, pm_amount: 10}, {pm_type: 1, pm_amount: 10}]}
    - {type: 'b', pm: [{pm_type: 1, pm_amount: 10}]}
  cluster:
    - {type: 'JP1', rack: [{rack_type: 'a', rack_amount: 5}, {rack_type: 'b', rack_amount: 5}]}
    - {type: 'FN1', rack: [{rack_type: 'a', rack_amount: 3}, {rack_type: 'b', rack_amount: 2}]}
architecture:
  region:
    - name: 'APAC'
      zone:
        - name: 'asia-northeast1'
          data_center:
            - {name: 'Japan', cluster: [{type: 'JP1', cluster_amount: 2}]}
            - {name: 'Korea', cluster: [{type: 'JP1', cluster_amount: 5}]}
    - name: 'EU'
      zone:
        - name: 'eu-north1'
          data_center:
            - {name: 'Finland', cluster: [{type: 'FN1', cluster_amount: 1}]}
"""
