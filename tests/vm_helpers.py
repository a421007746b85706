"""Shared helpers for the vm_scheduling tests."""
import importlib.util
import os

import numpy as np

from maro_b200 import _abi
from maro_b200.scenarios.vm_scheduling.data import build_vm_topology

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("gen_vm_golden", os.path.join(GOLDEN, "gen_vm_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
VM_CASES = gen.CASES


def vm_topology(spec):
    st = spec.get("start_tick", 0)
    return build_vm_topology(dict(spec["conf"]), st, st + spec["durations"])


def load_vm_golden(name):
    return np.load(os.path.join(GOLDEN, f"vm_{name}.npz"))


def metrics_vector(met_row):
    """int64[16] metrics row -> the 14 numbers of the golden files (floats decoded)."""
    d = _abi.vm_metrics_dict(met_row)
    return np.asarray([d[k] for k in gen.METRICS] + [d["latency_due_to_agent"], d["latency_due_to_resource"],
                                                      d["total_oversubscriptions"], d["total_overload_pms"],
                                                      d["total_overload_vms"]], np.float64)


def drive_vm(step_fn, gold, n_pm):
    """Replays the recorded action tape.  step_fn(actions or None) -> (status, dec row, metrics row)."""
    rows, valid, mets = [], [], []
    st, dec, met = step_fn(None)
    k = 0
    while st == 0:
        rows.append([dec[0], dec[1], dec[2], dec[3], dec[4], dec[5], dec[8], dec[9], dec[10]])
        v = np.full(n_pm, -1, np.int32)
        v[:dec[10]] = dec[12:12 + dec[10]]
        valid.append(v)
        mets.append(metrics_vector(met))
        a = gold["actions"][k]
        k += 1
        st, dec, met = step_fn(None if a[1] < 0 else a.reshape(1, 4))
    return (np.asarray(rows, np.int64).reshape(-1, 9), np.asarray(valid, np.int32).reshape(-1, n_pm),
            np.asarray(mets, np.float64).reshape(-1, 14), metrics_vector(met), st, dec)


def assert_metrics_close(got, want, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    if (err > 1e-9).any():
        bad = np.argwhere(err > 1e-9)[0]
        raise AssertionError(f"{what} metrics differ at {bad.tolist()}: got {got[tuple(bad)]!r} want {want[tuple(bad)]!r}")


def vm_named_frames(words_by_frame, topo):
    lay, fw = _abi.vm_frame_layout(topo)
    w = np.asarray(words_by_frame, np.int32)
    out = {}
    for node, attrs in lay.items():
        for a, (off, n, _) in attrs.items():
            x = w[:, off:off + n]
            out[f"{node}/{a}"] = x.view(np.float32) if a in _abi.VM_FLOAT_ATTRS else x
    return out


def assert_vm_snapshots_equal(get_snapshot, gold, topo):
    frames = gold["frames"].tolist()
    rows = []
    for f in frames:
        s = get_snapshot(int(f))
        assert s is not None, f"frame {f} missing"
        rows.append(s)
    named = vm_named_frames(rows, topo)
    for key, val in named.items():
        g = gold[key]
        if val.dtype == np.float32:
            ok = np.allclose(val, g, rtol=1e-6, atol=1e-7)
        else:
            ok = np.array_equal(val, g)
        if not ok:
            bad = np.argwhere(val != g)[0]
            raise AssertionError(f"{key} differs first at {bad.tolist()} (frame {frames[bad[0]]}): got {val[tuple(bad)]} want {g[tuple(bad)]}")
