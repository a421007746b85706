"""Host-side vm_scheduling loader: config + VM table / CPU readings traces -> flat static tables.

Init-time work.  Restates from scratch ``VmSchedulingBusinessEngine._load_configs / _init_structure``
(maro/simulator/scenarios/vm_scheduling/business_engine.py:131-440), the VM request stream (``:449-493``), the
``CpuReader`` with its file switching (``cpu_reader.py:9-77``) and the per-VM utilisation series semantics of
``VirtualMachine.add_utilization / get_utilization`` (``virtual_machine.py:73-90``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from ..citi_bike.data import read_bin


def _find_item(key, dictionary):
    for k, v in dictionary.items():
        if k == key:
            yield v
        elif isinstance(v, list):
            for item in v:
                yield from _find_item(key, item)
        elif isinstance(v, dict):
            yield from _find_item(key, v)


class _TickPicker:
    """ItemTickPicker (binary_reader.py:71-113) over an already filtered item array, time unit = seconds.  Items are
    consumed front to back: earlier timestamps are dropped, equal ones picked, the scan stops at the first later one."""

    def __init__(self, items, starttime):
        self.items, self.pos, self.starttime = items, 0, starttime
        self.ts = items["timestamp"].astype(np.int64)
        self.sorted = bool(len(self.ts) < 2 or (np.diff(self.ts) >= 0).all())

    def pick(self, tick):
        """-> the picked items (a structured array slice when the timestamps are sorted)."""
        t0 = self.starttime + tick
        if self.sorted:
            lo = max(self.pos, int(np.searchsorted(self.ts, t0, "left")))
            hi = max(lo, int(np.searchsorted(self.ts, t0, "right")))
            self.pos = hi
            return self.items[lo:hi]
        out = []
        while self.pos < len(self.items):
            ts = int(self.ts[self.pos])
            if ts >= t0:
                if ts - t0 < 1:
                    out.append(self.pos)
                    self.pos += 1
                else:
                    break
            else:
                self.pos += 1  # unsorted leftovers are dropped
        return self.items[out]


class _CpuReader:
    """CpuReader (cpu_reader.py:9-77)."""

    def __init__(self, data_path: str, start_tick: int):
        self.path = data_path
        self._open(first=True)
        while start_tick > self.et:
            self._switch()

    def _open(self, first: bool):
        items, self.st, self.et = read_bin(self.path)
        ts = items["timestamp"].astype(np.int64)
        lo, hi = (self.st + self.st, self.st + self.et) if first else (self.st, self.et)
        self.picker = _TickPicker(items[(ts >= lo) & (ts <= hi)], self.st)

    @staticmethod
    def _next_name(path):
        parts = path.split("-")
        parts[2] = str(int(parts[2]) + 1)
        return "-".join(parts)

    def _switch(self):
        self.path = self._next_name(self.path)
        self._open(first=False)

    def _pick(self, tick):
        got = self.picker.pick(tick - self.st)
        end_time = int(got["timestamp"][-1]) if len(got) else 0
        return got["vm_id"].astype(np.int64), got["cpu_utilization"].astype(np.float64), end_time

    def items_arrays(self, tick):
        """CpuReader.items(tick) as (vm ids, utilisations) in arrival order (a later row of the same id wins)."""
        ids, vals, end_time = self._pick(tick)
        if end_time == 8638:
            return ids, vals
        while end_time == self.et:
            nxt = os.path.expanduser(self._next_name(self.path))
            if not os.path.exists(nxt):
                break
            self._switch()
            if self.st == end_time:
                i2, v2, _ = self._pick(tick)
                ids, vals = np.concatenate([ids, i2]), np.concatenate([vals, v2])
        return ids, vals

    def items(self, tick) -> Dict[int, float]:
        ids, vals = self.items_arrays(tick)
        return {int(i): float(v) for i, v in zip(ids, vals)}


@dataclass
class VmTopology:
    config: dict
    start_tick: int
    max_tick: int
    # hierarchy
    n_pm: int
    n_rack: int
    n_cluster: int
    n_dc: int
    n_zone: int
    n_region: int
    pm_attr: np.ndarray  # [n_pm][8] cpu, memory, pm_type, region, zone, dc, cluster, rack
    pm_idle_energy: np.ndarray  # [n_pm] float64
    pmtype_power: np.ndarray  # [n_types][3] calibration, busy, idle (float64)
    rack_range: np.ndarray  # [n_rack][2] pm lo, hi
    rack_ids: np.ndarray  # [n_rack][4] region, zone, dc, cluster
    cluster_range: np.ndarray  # [n_cluster][2] rack lo, hi
    cluster_ids: np.ndarray  # [n_cluster][3] region, zone, dc
    dc_range: np.ndarray  # [n_dc][2] cluster lo, hi
    dc_ids: np.ndarray  # [n_dc][2] region, zone
    zone_range: np.ndarray  # [n_zone][2] dc lo, hi
    zone_ids: np.ndarray  # [n_zone] region
    region_range: np.ndarray  # [n_region][2] zone lo, hi
    # VM table (index = position in the trace)
    vm_attr: np.ndarray  # [n_vm][8] vm_id, sub_id, deploy_id, created tick, lifetime, category, cores, memory
    vm_price: np.ndarray  # [n_vm] float64 unit price per tick
    req_offset: np.ndarray  # [max_tick+1] requests of tick t = vm indices [offset[t], offset[t+1])
    vm_sorted_ids: np.ndarray  # vm ids sorted
    vm_sorted_idx: np.ndarray  # matching vm indices
    # utilisation series per VM from its request tick: forward-filled readings + "a reading exists at this tick"
    util_offset: np.ndarray  # [n_vm+1]
    util_val: np.ndarray  # float64
    util_has: np.ndarray  # uint8 as int32
    # scalars
    delay_duration: int
    buffer_budget: int
    ticks_per_hour: float
    max_cpu_over: float
    max_mem_over: float
    max_util_rate: float
    unit_energy_price: float
    pue: float
    kill_all: int
    error: str = None

    @property
    def n_vm(self) -> int:
        return len(self.vm_attr)


def load_vm_config(topology: str) -> dict:
    import yaml

    path = os.path.join(topology, "config.yml") if os.path.isdir(topology) else topology
    with open(path) as fp:
        conf = yaml.safe_load(fp)
    base = os.path.dirname(os.path.abspath(path))
    for k in ("VM_TABLE", "CPU_READINGS"):
        v = os.path.expanduser(str(conf[k]))
        conf[k] = v if os.path.isabs(v) else (v if os.path.exists(v) else os.path.join(base, v))
    return conf


def build_vm_topology(conf: dict, start_tick: int, max_tick: int) -> VmTopology:
    comp, arch = conf["components"], conf["architecture"]
    tph = conf["TICKS_PER_HOUR"]
    cluster_cfg = {c["type"]: {r["rack_type"]: r["rack_amount"] for r in c["rack"]} for c in comp["cluster"]}
    rack_cfg = {r["type"]: {p["pm_type"]: p["pm_amount"] for p in r["pm"]} for r in comp["rack"]}
    pm_cfg = {i: p for i, p in enumerate(comp["pm"])}

    def energy(pm_type, util):  # _cpu_utilization_to_energy_consumption (:671-688)
        pc = pm_cfg[pm_type]["power_curve"]
        u = min(1, util / 100)
        per_hour = pc["idle_power"] + (pc["busy_power"] - pc["idle_power"]) * (2 * u - pow(u, pc["calibration_parameter"]))
        return (per_hour / tph) / 1000

    pms, racks, rack_ids, clusters, cluster_ids, dcs, dc_ids, zones, zone_ids, regions = ([] for _ in range(10))
    # _init_regions / zones / data_centers / clusters / racks / pms (:300-440): depth-first, ids in visiting order
    for region_list in _find_item("region", arch):
        for region_dict in region_list:
            rid = len(regions)
            z_lo = len(zones)
            for zone_dict in region_dict["zone"]:
                zid = len(zones)
                d_lo = len(dcs)
                for dc_dict in zone_dict["data_center"]:
                    did = len(dcs)
                    c_lo = len(clusters)
                    for cl in dc_dict["cluster"]:
                        for _ in range(cl["cluster_amount"]):
                            cid = len(clusters)
                            r_lo = len(racks)
                            for rack_type, rack_amount in cluster_cfg[cl["type"]].items():
                                for _ in range(rack_amount):
                                    kid = len(racks)
                                    p_lo = len(pms)
                                    for pm_type, pm_amount in rack_cfg[rack_type].items():
                                        for _ in range(pm_amount):
                                            pms.append([pm_cfg[pm_type]["cpu"], pm_cfg[pm_type]["memory"], pm_type,
                                                        rid, zid, did, cid, kid])
                                    racks.append([p_lo, len(pms)])
                                    rack_ids.append([rid, zid, did, cid])
                            clusters.append([r_lo, len(racks)])
                            cluster_ids.append([rid, zid, did])
                    dcs.append([c_lo, len(clusters)])
                    dc_ids.append([rid, zid])
                zones.append([d_lo, len(dcs)])
                zone_ids.append(rid)
            regions.append([z_lo, len(zones)])
    pm_attr = np.asarray(pms, np.int32).reshape(-1, 8)
    pm_idle = np.asarray([energy(int(t), 0) for t in pm_attr[:, 2]], np.float64)
    pmtype_power = np.asarray([[p["power_curve"]["calibration_parameter"], p["power_curve"]["busy_power"],
                                p["power_curve"]["idle_power"]] for p in comp["pm"]], np.float64)

    # ---- VM request stream: items_tick_picker(start_tick, max_tick, "s") (:89-90)
    items, st, et = read_bin(conf["VM_TABLE"])
    ts = items["timestamp"].astype(np.int64)
    ok = (ts >= st + start_tick) & (ts <= st + max_tick)
    items = items[ok]
    tick = (items["timestamp"].astype(np.int64) - st)
    cur = np.maximum.accumulate(tick) if len(tick) else tick
    keep = (tick >= cur) & (tick < max_tick)
    items, tick = items[keep], tick[keep]
    n_vm = len(items)
    vm_attr = np.zeros((n_vm, 8), np.int32)
    for k, name in enumerate(("vm_id", "sub_id", "deploy_id", None, "vm_lifetime", "vm_category", "vm_cpu_cores", "vm_memory")):
        vm_attr[:, k] = tick if name is None else items[name]
    price = (conf["PRICE_PER_CPU_CORES_PER_HOUR"] * vm_attr[:, 6].astype(np.float64)
             + conf["PRICE_PER_MEMORY_PER_HOUR"] * vm_attr[:, 7].astype(np.float64)) / tph
    req_offset = np.zeros(max_tick + 1, np.int32)
    np.add.at(req_offset, tick + 1, 1)
    req_offset = np.cumsum(req_offset).astype(np.int32)
    order = np.argsort(vm_attr[:, 0], kind="stable")

    # ---- utilisation series: replay CpuReader.items(tick) for every tick, forward-fill per VM from its request tick.
    # A VM needs its series only while it can still be pending or live: at most buffer budget + lifetime ticks after
    # its request (a zero lifetime never matches a later deletion tick, so such a VM lives to the end).
    reader = _CpuReader(conf["CPU_READINGS"], start_tick)
    req, life = vm_attr[:, 3].astype(np.int64), vm_attr[:, 4].astype(np.int64)
    budget = int(conf["BUFFER_TIME_BUDGET"]) + int(conf["DELAY_DURATION"]) + 2
    last = np.where(life <= 0, max_tick - 1, np.minimum(max_tick - 1, req + life + budget))
    util_offset = np.zeros(n_vm + 1, np.int64)
    np.cumsum(np.maximum(last - req + 1, 0), out=util_offset[1:])
    util_val = np.zeros(int(util_offset[-1]), np.float64)
    util_has = np.zeros(int(util_offset[-1]), np.int32)
    sorted_ids, sorted_idx = vm_attr[order, 0].astype(np.int64), order
    unique_ids = bool(n_vm < 2 or (np.diff(sorted_ids) > 0).all())
    cur = np.zeros(n_vm, np.float64)
    hasnow = np.zeros(n_vm, np.int32)
    error = None
    for t in range(start_tick, max_tick):
        rid, rval = reader.items_arrays(t)
        hi = int(req_offset[t + 1])
        if len(rid) and n_vm:
            if unique_ids:
                pos = np.minimum(np.searchsorted(sorted_ids, rid), n_vm - 1)
                ok = sorted_ids[pos] == rid
                vi, vv = sorted_idx[pos[ok]], rval[ok]
            else:  # several table rows share a vm id: each of them sees the reading
                lo_, hi_ = np.searchsorted(sorted_ids, rid, "left"), np.searchsorted(sorted_ids, rid, "right")
                rep = hi_ - lo_
                vi = sorted_idx[np.concatenate([np.arange(a_, b_) for a_, b_ in zip(lo_, hi_)])] if rep.sum() else rid[:0]
                vv = np.repeat(rval, rep)
            if len(vi) != len(np.unique(vi)):  # a later row of the same id wins
                _, first_rev = np.unique(vi[::-1], return_index=True)
                keep_ = np.sort(len(vi) - 1 - first_rev)
                vi, vv = vi[keep_], vv[keep_]
            cur[vi] = vv
            hasnow[vi] = 1
        else:
            vi = np.zeros(0, np.int64)
        new = np.arange(int(req_offset[t]), hi)
        missing = new[hasnow[new] == 0]
        if len(missing):
            # the reference raises this from BusinessEngine.step (:476-477); surfaced when the env is created
            error = error or f"The VM id: '{int(vm_attr[missing[0], 0])}' does not exist at this tick."
            cur[missing] = 0.0
            hasnow[missing] = 1
        act = np.nonzero(last[:hi] >= t)[0]
        if len(act):
            p = util_offset[act] + (t - req[act])
            util_val[p] = cur[act]
            util_has[p] = hasnow[act]
        hasnow[vi] = 0
        hasnow[missing] = 0
    return VmTopology(
        config=conf, start_tick=start_tick, max_tick=max_tick, n_pm=len(pm_attr), n_rack=len(racks),
        n_cluster=len(clusters), n_dc=len(dcs), n_zone=len(zones), n_region=len(regions), pm_attr=pm_attr,
        pm_idle_energy=pm_idle, pmtype_power=pmtype_power, rack_range=np.asarray(racks, np.int32).reshape(-1, 2),
        rack_ids=np.asarray(rack_ids, np.int32).reshape(-1, 4), cluster_range=np.asarray(clusters, np.int32).reshape(-1, 2),
        cluster_ids=np.asarray(cluster_ids, np.int32).reshape(-1, 3), dc_range=np.asarray(dcs, np.int32).reshape(-1, 2),
        dc_ids=np.asarray(dc_ids, np.int32).reshape(-1, 2), zone_range=np.asarray(zones, np.int32).reshape(-1, 2),
        zone_ids=np.asarray(zone_ids, np.int32), region_range=np.asarray(regions, np.int32).reshape(-1, 2),
        vm_attr=vm_attr, vm_price=price, req_offset=req_offset, vm_sorted_ids=vm_attr[order, 0].copy(),
        vm_sorted_idx=order.astype(np.int32), util_offset=util_offset.astype(np.int32), util_val=util_val,
        util_has=util_has, delay_duration=int(conf["DELAY_DURATION"]), buffer_budget=int(conf["BUFFER_TIME_BUDGET"]),
        ticks_per_hour=float(tph), max_cpu_over=float(conf["MAX_CPU_OVERSUBSCRIPTION_RATE"]),
        max_mem_over=float(conf["MAX_MEM_OVERSUBSCRIPTION_RATE"]), max_util_rate=float(conf["MAX_UTILIZATION_RATE"]),
        unit_energy_price=float(conf["UNIT_ENERGY_PRICE_PER_KWH"]), pue=float(conf["POWER_USAGE_EFFICIENCY"]),
        kill_all=1 if conf["KILL_ALL_VMS_IF_OVERLOAD"] else 0, error=error)
