"""Host-side citi_bike loader: config + MARO ``.bin`` traces + station / distance CSVs -> flat static tables.

Init-time work.  Restates from scratch what the reference does in
``maro/data_lib/binary_reader.py`` (header ``"<4s b I Q I QQ QQ qq"`` + YAML meta + packed little-endian items,
``ItemTickPicker.items`` :71-113), ``citi_bike/business_engine.py:218-396`` (``_init``, ``_init_adj_matrix``,
``_update_station_extra_features``), ``stations_info.py``, ``adj_loader.py``, ``weather_table.py`` and the neighbour
ordering of ``decision_strategy.py:385-397``.
"""
from __future__ import annotations

import csv
import datetime
import os
import struct
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_HEADER = struct.Struct("<4s b I Q I QQ QQ qq")
_DTYPES = {"i": "<i4", "i4": "<i4", "i2": "<i2", "i8": "<i8", "f": "<f4", "d": "<f8"}


def read_bin(path: str):
    """Parse a MARO binary trace -> (structured ndarray of items, starttime, endtime)."""
    import yaml

    with open(os.path.expanduser(path), "rb") as fp:
        buf = fp.read()
    name, ftype, ver, count, isize, moff, msize, doff, dsize, st, et = _HEADER.unpack_from(buf)
    if name != b"MARO":
        raise ValueError(f"{path}: not a MARO binary file")

    class _Loader(yaml.SafeLoader):
        pass

    _Loader.add_multi_constructor("!", lambda loader, suffix, node: loader.construct_mapping(node))
    meta = yaml.load(buf[moff:moff + msize].decode(), Loader=_Loader)
    fields = []
    for a in meta["attributes"]:
        slot = int(a.get("slot") or 1)
        dt = _DTYPES[a["dtype"]]
        fields.append((a["name"], dt) if slot == 1 else (a["name"], dt, (slot,)))
    dtype = np.dtype(fields)
    assert dtype.itemsize == isize, (dtype.itemsize, isize)
    items = np.frombuffer(buf, dtype, count, doff)
    return items, int(st), int(et)


def _tz(name: str):
    try:
        from dateutil.tz import gettz

        return gettz(name)
    except Exception:  # pragma: no cover
        from zoneinfo import ZoneInfo

        return ZoneInfo(name)


def _is_us_holiday(date) -> bool:
    try:
        import holidays  # optional; the reference uses holidays.US()

        return date in _is_us_holiday.__dict__.setdefault("cal", holidays.US())
    except Exception:
        return False


@dataclass
class BikeTopology:
    config: dict
    n_stations: int
    start_tick: int
    max_tick: int
    station_bikes: np.ndarray
    station_capacity: np.ndarray
    station_id: np.ndarray
    nbr_offset: np.ndarray  # [S+1] neighbours sorted by distance (distance != 0), decision_strategy.py:385-397
    nbr_idx: np.ndarray
    trip_offset: np.ndarray  # [max_tick+1] into trip_* (trips of tick t = [offset[t], offset[t+1]))
    trip_src: np.ndarray
    trip_dst: np.ndarray
    trip_dur: np.ndarray
    day_of_tick: np.ndarray  # [max_tick] -> row of day_feat
    day_feat: np.ndarray  # [n_days][4] weekday, holiday, weather, temperature
    resolution: int
    time_mean: float
    time_std: float
    supply_ratio: float
    demand_ratio: float
    scope_low: float
    scope_high: float
    extra_cost_mode: int  # 0 source, 1 target, 2 target_neighbors
    transfer_seed: int
    #: decision.action_scope.filters as (type, num, windows): 0 distance / 1 requirements / 2 trip_window
    #: (decision_strategy.py:15-163, 345-362)
    filters: tuple = ()

    @property
    def max_delay(self) -> int:
        d = int(self.trip_dur.max()) if len(self.trip_dur) else 1
        return max(d, int(self.time_mean + 8 * self.time_std) + 1, 2)


def build_bike_topology(config: dict, start_tick: int, max_tick: int, transfer_seed: int = 0) -> BikeTopology:
    dec = config["decision"]
    # ---- stations (stations_info.py)
    rows = []
    with open(os.path.expanduser(config["stations_init_data"])) as fp:
        for row in csv.DictReader(fp):
            rows.append((int(row["station_index"]), int(row["init"]), int(row["capacity"]), int(float(row["station_id"]))))
    S = len(rows)
    bikes, cap, sid = np.zeros(S, np.int32), np.zeros(S, np.int32), np.zeros(S, np.int32)
    for idx, b, c, i in rows:
        bikes[idx], cap[idx], sid[idx] = b, c, i
    # ---- distance adjacency -> neighbours sorted by distance (stable: ties keep index order)
    adj = []
    with open(os.path.expanduser(config["distance_adj_data"])) as fp:
        for k, row in enumerate(csv.reader(fp)):
            if k == 0:
                continue
            adj.append([float(c) for c in row])
    adj = np.asarray(adj, np.float64).reshape(S, S)
    nbr_offset, nbr_idx = [0], []
    for s in range(S):
        nb = sorted([(i, d) for i, d in enumerate(adj[s]) if d != 0.0], key=lambda kv: kv[1])
        nbr_idx += [i for i, _ in nb]
        nbr_offset.append(len(nbr_idx))
    # the action-scope filter chain, in configuration order (decision_strategy.py:345-362)
    kinds = {"distance": 0, "requirements": 1, "trip_window": 2}
    filters = []
    for f in dec["action_scope"].get("filters", []) or []:
        if f["type"] not in kinds:
            raise KeyError(f"unknown action-scope filter type {f['type']!r}")
        filters.append((kinds[f["type"]], int(f["num"]), int(f.get("windows", 0) or 0)))
    # ---- trips -> per-tick lists (ItemTickPicker.items, binary_reader.py:80-113)
    items, st, et = read_bin(config["trip_data"])
    ts = items["timestamp"].astype(np.int64)
    end_time = st + max_tick * 60
    ok = (ts >= st + start_tick * 60) & (ts <= end_time)
    ts, dur = ts[ok], items["durations"][ok].astype(np.int32)
    src, dst = items["src_station"][ok].astype(np.int32), items["dest_station"][ok].astype(np.int32)
    tick = ((ts - st) // 60).astype(np.int64)
    # items behind the picker's cursor are dropped (unsorted input); items at/after max_tick are never visited
    cur = np.maximum.accumulate(tick) if len(tick) else tick
    keep = (tick >= cur) & (tick < max_tick)
    tick, dur, src, dst = tick[keep], dur[keep], src[keep], dst[keep]
    trip_offset = np.zeros(max_tick + 1, np.int32)
    np.add.at(trip_offset, tick + 1, 1)
    trip_offset = np.cumsum(trip_offset).astype(np.int32)
    # ---- per-day features (_update_station_extra_features, business_engine.py:370-396; weather_table.py)
    tz = _tz(config["time_zone"])
    start_dt = datetime.datetime.fromtimestamp(st, datetime.timezone.utc).astimezone(tz)
    witems, _, _ = read_bin(config["weather_data"])
    lut = {}
    for it in witems:
        d = datetime.datetime.fromtimestamp(int(it["timestamp"]), datetime.timezone.utc).astimezone(tz).date()
        lut[d] = (int(it["weather"]), float(it["temp"]))
    day_of_tick = np.zeros(max_tick, np.int32)
    feats, last = [], None
    for t in range(max_tick):
        d = (start_dt + datetime.timedelta(minutes=t)).date()  # wall-clock arithmetic like relativedelta(minutes=)
        if d != last:
            w, temp = lut.get(d, (0, 0))
            feats.append([d.weekday(), 1 if _is_us_holiday(d) else 0, int(w), int(temp)])
            last = d
        day_of_tick[t] = len(feats) - 1
    mode = {"source": 0, "target": 1}.get(str(dec["extra_cost_mode"]), 2)
    return BikeTopology(
        config=config, n_stations=S, start_tick=start_tick, max_tick=max_tick, station_bikes=bikes,
        station_capacity=cap, station_id=sid, nbr_offset=np.asarray(nbr_offset, np.int32),
        nbr_idx=np.asarray(nbr_idx, np.int32), trip_offset=trip_offset, trip_src=src, trip_dst=dst, trip_dur=dur,
        day_of_tick=day_of_tick, day_feat=np.asarray(feats, np.int32).reshape(-1, 4), resolution=int(dec["resolution"]),
        time_mean=float(dec["effective_time_mean"]), time_std=float(dec["effective_time_std"]),
        supply_ratio=float(dec["supply_water_mark_ratio"]), demand_ratio=float(dec["demand_water_mark_ratio"]),
        scope_low=float(dec["action_scope"]["low"]), scope_high=float(dec["action_scope"]["high"]),
        extra_cost_mode=mode, transfer_seed=int(transfer_seed) & 0xFFFFFFFF, filters=tuple(filters))


def load_bike_config(topology: str) -> dict:
    """``topology`` = folder with ``config.yml`` (the reference resolves built-in names to such a folder too)."""
    import yaml

    path = os.path.join(topology, "config.yml") if os.path.isdir(topology) else topology
    with open(path) as fp:
        conf = yaml.safe_load(fp)
    base = os.path.dirname(os.path.abspath(path))
    for k in ("trip_data", "weather_data", "stations_init_data", "distance_adj_data"):
        v = os.path.expanduser(str(conf[k]))
        conf[k] = v if os.path.isabs(v) else os.path.join(base, v)
    return conf
