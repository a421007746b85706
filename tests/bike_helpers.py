"""Shared helpers for the citi_bike tests."""
import importlib.util
import os

import numpy as np
import yaml

from maro_b200 import _abi
from maro_b200.scenarios.citi_bike.data import build_bike_topology

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("gen_bike_golden", os.path.join(GOLDEN, "gen_bike_golden.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)
BIKE_CASES = gen.CASES


def bike_config(data_name, decision_text=None):
    src = os.path.join(GOLDEN, data_name)
    if data_name == "bike_toy":
        conf = yaml.safe_load(gen.TOY_DECISION)
        files = dict(trip_data="trips.bin", weather_data="KNYC_daily.bin", stations_init_data="station_meta.csv",
                     distance_adj_data="distance_adj.csv")
    else:
        if decision_text:
            conf = yaml.safe_load(decision_text)
        else:
            with open(os.path.join(src, "decision.yml")) as fp:
                conf = yaml.safe_load(fp)
        files = dict(trip_data="trips.bin", weather_data="weathers.bin", stations_init_data="stations.csv",
                     distance_adj_data="distance_adj.csv")
    for k, v in files.items():
        conf[k] = os.path.join(src, v)
    return conf


def bike_topology(spec):
    st = spec.get("start_tick", 0)
    return build_bike_topology(bike_config(spec["data"], spec.get("decision_text")), st, st + spec["durations"], transfer_seed=spec["np_seed"])


def load_bike_golden(name):
    return np.load(os.path.join(GOLDEN, f"bike_{name}.npz"))


def greedy_py(dec):
    """examples/citi_bike/greedy/launcher.py:35-65 with top-1, on a decision row."""
    station, n = int(dec[1]), int(dec[4])
    best = None
    for k in range(n):
        idx, v = int(dec[8 + 2 * k]), int(dec[9 + 2 * k])
        if idx == station:
            continue
        if best is None or (v, idx) > best:
            best = (v, idx)
    if best is None:
        return [-1, -1, 0, 0]
    v, idx = best
    return [station, idx, v, 0] if dec[3] == 0 else [idx, station, v, 0]


def drive_bike(step_fn, spec, S):
    """step_fn(actions or None) -> (status, dec row, metrics).  Returns rows[n][8], scopes[n][S], final metrics, status."""
    rows, scopes = [], []
    st, dec, met = step_fn(None)
    while st == 0:
        rows.append([dec[0], dec[1], dec[2], dec[3], dec[4]] + list(met))
        sv = np.full(S, -1, np.int64)
        for k in range(int(dec[4])):
            sv[int(dec[8 + 2 * k])] = int(dec[9 + 2 * k])
        scopes.append(sv)
        act = np.asarray([greedy_py(dec)], np.int32) if spec["policy"] == 1 else None
        st, dec, met = step_fn(act)
    return (np.asarray(rows, np.int64).reshape(-1, 8), np.asarray(scopes, np.int64).reshape(-1, S),
            np.asarray(met, np.int64), st, dec)


def bike_named_frames(words_by_frame, S):
    lay, fw = _abi.bike_frame_layout(S)
    w = np.asarray(words_by_frame, np.int32)
    out = {}
    for a, (off, n, slots) in lay["stations"].items():
        out["stations/" + a] = w[:, off:off + S]
    off, _, slots = lay["matrices"]["trips_adj"]
    out["matrices/trips_adj"] = w[:, off:off + slots]
    return out


def assert_bike_snapshots_equal(get_snapshot, gold, S):
    frames = gold["frames"].tolist()
    rows = []
    for f in frames:
        s = get_snapshot(int(f))
        assert s is not None, f"frame {f} missing"
        rows.append(s)
    named = bike_named_frames(rows, S)
    for key, val in named.items():
        g = gold[key]
        if not np.array_equal(val, g):
            bad = np.argwhere(val != g)[0]
            raise AssertionError(f"{key} differs first at {bad.tolist()} (frame {frames[bad[0]]}): got {val[tuple(bad)]} want {g[tuple(bad)]}")
