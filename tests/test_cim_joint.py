"""CPU: DecisionMode.Joint (maro/simulator/core.py:354-366) — the C oracle and the device logic under the emulator against
traces of the unmodified reference run in Joint mode (tests/golden/gen_cim_joint_golden.py)."""
import importlib.util
import os

import numpy as np
import pytest

from maro_b200.scenarios.cim.topology import build_topology
from oracle.cim_oracle import CimOracle

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("gen_joint", os.path.join(HERE, "golden", "gen_cim_joint_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)


def drive_joint(step_fn, spec):
    """step_fn(answers or None) -> (status, rows [n][8], metrics[3]); returns (rows [n][7], step metrics, final metrics)"""
    rows, mets, step, ordinal = [], [], 0, 0
    st, decs, met = step_fn(None)
    while st == 0:
        cur = [[int(x) for x in d[:6]] for d in decs]
        for r in cur:
            rows.append([step] + r)
        mets.append([int(x) for x in met])
        ans = gen.answers(cur, spec["pseed"], ordinal)
        ordinal += len(cur)
        step += 1
        st, decs, met = step_fn(ans)
    return np.asarray(rows, np.int64).reshape(-1, 7), np.asarray(mets, np.int64).reshape(-1, 3), np.asarray(met, np.int64), st


def check_against_gold(name, rows, mets, final, snapshot):
    gold = np.load(os.path.join(HERE, "golden", f"cim_joint_{name}.npz"))
    assert rows.shape == gold["rows"].shape
    if not np.array_equal(rows, gold["rows"]):
        bad = np.argwhere(rows != gold["rows"])[0]
        raise AssertionError(f"decision {bad[0]}: got {rows[bad[0]]} want {gold['rows'][bad[0]]}")
    assert np.array_equal(mets, gold["step_metrics"]) and final.tolist() == gold["final_metrics"].tolist()
    return gold


@pytest.mark.parametrize("name", sorted(gen.CASES))
def test_oracle_joint_mode_matches_reference_trace(name):
    from helpers import named_frames

    spec = gen.CASES[name]
    topo = build_topology(spec["topology"], spec["durations"])
    o = CimOracle(topo, 0, spec.get("snapshot_resolution", 1), spec.get("max_snapshots"))
    rows, mets, final, st = drive_joint(o.step_joint, spec)
    assert st == 1
    gold = check_against_gold(name, rows, mets, final, o.snapshot)
    fr = named_frames([o.snapshot(int(f)) for f in gold["frames"]], topo)
    for key in ("ports/empty", "ports/full", "ports/on_consignee", "ports/on_shipper", "ports/shortage", "ports/acc_shortage",
                "ports/transfer_cost", "matrices/vessel_plans"):
        assert np.array_equal(np.asarray(fr[key], np.float64), gold[key]), key
    for key in ("vessels/empty", "vessels/full", "vessels/remaining_space", "vessels/early_discharge"):
        assert np.array_equal(np.asarray(fr[key][:, :, 0], np.float64), gold[key]), key


@pytest.mark.parametrize("name,lanes", [(n, 0) for n in sorted(gen.CASES)] + [("toy4p_l00_160", 16), ("toy5p_l03_140_res4", 32)])
def test_emulated_kernel_joint_mode_matches_reference_trace(name, lanes):
    """(lane k of a replica's group carries answer k: a step takes at most `lanes per replica` answers, >= 8)"""
    from emul import EmulEnv

    spec = gen.CASES[name]
    topo = build_topology(spec["topology"], spec["durations"])
    e = EmulEnv(topo, 1, 0, spec.get("snapshot_resolution", 1), spec.get("max_snapshots"), lanes=lanes, decision_mode=1)
    rows, mets, final, st = drive_joint(e.step_joint, spec)
    assert st == 1
    check_against_gold(name, rows, mets, final, e.snapshot)
    o = CimOracle(topo, 0, spec.get("snapshot_resolution", 1), spec.get("max_snapshots"))
    drive_joint(o.step_joint, spec)
    assert np.array_equal(e.frame(), o.frame())
    assert e.counters().tolist() == o.counters().tolist()
    for f in (0, 5, spec["durations"] // spec.get("snapshot_resolution", 1) - 1):
        a, b = e.snapshot(f), o.snapshot(f)
        assert (a is None and b is None) or np.array_equal(a, b), f


def env_joint_case(name="toy5p_l03_140_res4"):
    """maro_b200.simulator.Env(decision_mode=DecisionMode.Joint): the façade returns the list of decision events and takes the
    list of answers, like the reference; replayed against the reference trace"""
    from maro_b200.scenarios.cim.common import Action, ActionType
    from maro_b200.simulator import DecisionMode, Env

    spec = gen.CASES[name]
    env = Env("cim", spec["topology"], durations=spec["durations"], snapshot_resolution=spec.get("snapshot_resolution", 1),
              max_snapshots=spec.get("max_snapshots"), decision_mode=DecisionMode.Joint)
    gold = np.load(os.path.join(HERE, "golden", f"cim_joint_{name}.npz"))
    rows, mets, step, ordinal = [], [], 0, 0
    metrics, decs, done = env.step(None)
    while not done:
        assert isinstance(decs, list) and len(decs) >= 1
        cur = [[d.tick, d.port_idx, d.vessel_idx, d.action_scope.load, d.action_scope.discharge, d.early_discharge] for d in decs]
        rows += [[step] + r for r in cur]
        mets.append([metrics["order_requirements"], metrics["container_shortage"], metrics["operation_number"]])
        acts = [None if a is None else Action(a[0], a[1], a[2], ActionType.DISCHARGE if a[3] else ActionType.LOAD)
                for a in gen.answers(cur, spec["pseed"], ordinal)]
        ordinal += len(cur)
        step += 1
        metrics, decs, done = env.step(acts)
    assert np.array_equal(np.asarray(rows, np.int64), gold["rows"]) and np.array_equal(np.asarray(mets, np.int64), gold["step_metrics"])
    assert [metrics["order_requirements"], metrics["container_shortage"], metrics["operation_number"]] == gold["final_metrics"].tolist()
    frames = gold["frames"].tolist()
    got = env.snapshot_list["ports"][frames::["empty", "full"]].reshape(len(frames), -1, 2)
    assert np.array_equal(got[:, :, 0], gold["ports/empty"]) and np.array_equal(got[:, :, 1], gold["ports/full"])
    assert env.step(None) == (None, None, True)
    env.close()


def test_env_facade_joint_mode_emulated(monkeypatch):
    import maro_b200.simulator.env as env_mod
    from emul_batch import EmulCimBatch

    monkeypatch.setattr(env_mod, "CimBatch", EmulCimBatch)
    env_joint_case()


def vector_env_joint_case(name="toy4p_l00_160"):
    """VectorEnv(decision_mode=Joint): every env returns the list of its tick's decision events and takes a list of answers"""
    from maro_b200.scenarios.cim.common import Action, ActionType
    from maro_b200.simulator import DecisionMode
    from maro_b200.vector_env import VectorEnv

    spec = gen.CASES[name]
    gold = np.load(os.path.join(HERE, "golden", f"cim_joint_{name}.npz"))
    B = 3
    with VectorEnv(batch_num=B, scenario="cim", topology=spec["topology"], durations=spec["durations"],
                   snapshot_resolution=spec.get("snapshot_resolution", 1), max_snapshots=spec.get("max_snapshots"),
                   decision_mode=DecisionMode.Joint) as env:
        rows, step, ordinal = [[] for _ in range(B)], 0, 0
        metrics, decs, done = env.step(None)
        while not done:
            acts = []
            for i in range(B):
                cur = [[d.tick, d.port_idx, d.vessel_idx, d.action_scope.load, d.action_scope.discharge, d.early_discharge] for d in decs[i]]
                rows[i] += [[step] + r for r in cur]
                acts.append([None if a is None else Action(a[0], a[1], a[2], ActionType.DISCHARGE if a[3] else ActionType.LOAD)
                             for a in gen.answers(cur, spec["pseed"], ordinal)])
            ordinal += len(decs[0])
            step += 1
            metrics, decs, done = env.step(acts)
        for i in range(B):
            assert np.array_equal(np.asarray(rows[i], np.int64), gold["rows"])
            assert [metrics[i][k] for k in ("order_requirements", "container_shortage", "operation_number")] == gold["final_metrics"].tolist()


def test_vector_env_joint_mode_emulated(monkeypatch):
    import maro_b200.vector_env.vector_env as venv_mod
    from emul_batch import EmulCimBatch

    monkeypatch.setattr(venv_mod, "CimBatch", EmulCimBatch)
    vector_env_joint_case()
