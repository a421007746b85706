"""ctypes mirror of ``include/maro_b200.h`` (struct layouts + enums).  Pure declarations, no library loading."""
from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 2

DEC_TICK, DEC_PORT, DEC_VESSEL, DEC_SCOPE_LOAD, DEC_SCOPE_DISCHARGE, DEC_EARLY_DISCHARGE, DEC_STATUS, DEC_STEP = range(8)
DECISION_WORDS = 8
ACTION_WORDS = 4
METRIC_WORDS = 3
STATUS_DECISION, STATUS_DONE, STATUS_FINISHED, STATUS_INACTIVE = 0, 1, 2, 3
STATUS_BAD_ACTION, STATUS_QUEUE_OVERFLOW = -1, -2
ACTION_LOAD, ACTION_DISCHARGE = 0, 1
NODE_PORTS, NODE_VESSELS, NODE_MATRICES = 0, 1, 2

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class MaroCimTopology(C.Structure):
    _fields_ = [
        ("n_ports", C.c_int32), ("n_vessels", C.c_int32), ("n_routes", C.c_int32),
        ("past_stop_number", C.c_int32), ("future_stop_number", C.c_int32),
        ("max_tick", C.c_int32), ("order_mode", C.c_int32), ("total_containers", C.c_int32),
        ("container_volume", C.c_double),
        ("port_capacity", _i32p), ("port_init_empty", _i32p),
        ("full_return_base", _f64p), ("full_return_noise", _f64p),
        ("empty_return_base", _f64p), ("empty_return_noise", _f64p),
        ("source_base", _f64p), ("source_noise", _f64p),
        ("target_offset", _i32p), ("target_port", _i32p), ("target_base", _f64p), ("target_noise", _f64p),
        ("vessel_capacity", _i32p), ("vessel_init_empty", _i32p), ("vessel_route", _i32p),
        ("vessel_period", _i32p), ("vessel_route_start", _i32p), ("vessel_leg_offset", _i32p),
        ("vessel_leg", _i32p), ("stop_offset", _i32p), ("stop_arrival", _i32p), ("stop_leave", _i32p),
        ("stop_port", _i32p), ("route_offset", _i32p), ("route_port", _i32p), ("order_proportion", _i32p),
        ("order_number_seed", C.c_uint32), ("buffer_time_seed", C.c_uint32),
    ]


class MaroCimConfig(C.Structure):
    _fields_ = [
        ("n_replicas", C.c_int32), ("start_tick", C.c_int32), ("snapshot_resolution", C.c_int32),
        ("max_snapshots", C.c_int32), ("device", C.c_int32), ("queue_capacity", C.c_int32),
        ("max_actions", C.c_int32), ("replica_topology", _i32p), ("decision_mode", C.c_int32),
    ]


_I32_FIELDS = ("port_capacity", "port_init_empty", "target_offset", "target_port", "vessel_capacity",
               "vessel_init_empty", "vessel_route", "vessel_period", "vessel_route_start", "vessel_leg_offset",
               "vessel_leg", "stop_offset", "stop_arrival", "stop_leave", "stop_port", "route_offset", "route_port",
               "order_proportion")
_F64_FIELDS = ("full_return_base", "full_return_noise", "empty_return_base", "empty_return_noise", "source_base",
               "source_noise", "target_base", "target_noise")


def topology_struct(topo):
    """Build a ``MaroCimTopology`` from a ``CimTopology``; returns (struct, keepalive list of arrays)."""
    s = MaroCimTopology()
    keep = []
    for name in ("n_ports", "n_vessels", "n_routes", "past_stop_number", "future_stop_number", "max_tick",
                 "order_mode", "total_containers"):
        setattr(s, name, int(getattr(topo, name)))
    s.container_volume = float(topo.container_volume)
    for name in _I32_FIELDS:
        a = np.ascontiguousarray(getattr(topo, name), dtype=np.int32)
        if a.size == 0:
            a = np.zeros(1, np.int32)
        keep.append(a)
        setattr(s, name, a.ctypes.data_as(_i32p))
    for name in _F64_FIELDS:
        a = np.ascontiguousarray(getattr(topo, name), dtype=np.float64)
        if a.size == 0:
            a = np.zeros(1, np.float64)
        keep.append(a)
        setattr(s, name, a.ctypes.data_as(_f64p))
    s.order_number_seed = int(topo.stream_seeds["order_number"]) & 0xFFFFFFFF
    s.buffer_time_seed = int(topo.stream_seeds["buffer_time"]) & 0xFFFFFFFF
    return s, keep


# canonical frame layout helpers (DESIGN.md "Frame layout") -------------------------------------------------
PORT_ATTRS = ("acc_booking", "acc_fulfillment", "acc_shortage", "booking", "capacity", "empty", "fulfillment",
              "full", "on_consignee", "on_shipper", "shortage", "transfer_cost")
VESSEL_SCALAR_ATTRS = ("capacity", "early_discharge", "empty", "full", "is_parking", "last_loc_idx",
                       "loc_port_idx", "next_loc_idx", "remaining_space", "route_idx")
VESSEL_LIST_ATTRS = ("past_stop_list", "past_stop_tick_list", "future_stop_list", "future_stop_tick_list")
MATRIX_ATTRS = ("full_on_ports", "full_on_vessels", "vessel_plans")


def frame_layout(P: int, V: int, past_n: int, future_n: int):
    """name -> (word offset, n_nodes, slots) for every attribute of every node type."""
    lay = {"ports": {}, "vessels": {}, "matrices": {}}
    off = 0
    for a in PORT_ATTRS:
        lay["ports"][a] = (off, P, 1)
        off += P
    for a in VESSEL_SCALAR_ATTRS:
        lay["vessels"][a] = (off, V, 1)
        off += V
    for a, n in zip(VESSEL_LIST_ATTRS, (past_n, past_n, future_n, future_n)):
        lay["vessels"][a] = (off, V, n)
        off += V * n
    for a, n in zip(MATRIX_ATTRS, (P * P, V * P, V * P)):
        lay["matrices"][a] = (off, 1, n)
        off += n
    return lay, off


# ---------------------------------------------------------------------------------------------------- citi_bike
class MaroBikeTopology(C.Structure):
    _fields_ = [
        ("n_stations", C.c_int32), ("n_days", C.c_int32), ("max_tick", C.c_int32), ("resolution", C.c_int32),
        ("extra_cost_mode", C.c_int32), ("transfer_seed", C.c_uint32),
        ("time_mean", C.c_double), ("time_std", C.c_double), ("supply_ratio", C.c_double),
        ("demand_ratio", C.c_double), ("scope_low", C.c_double), ("scope_high", C.c_double),
        ("station_bikes", _i32p), ("station_capacity", _i32p), ("station_id", _i32p), ("nbr_offset", _i32p),
        ("nbr_idx", _i32p), ("trip_offset", _i32p), ("trip_src", _i32p), ("trip_dst", _i32p), ("trip_dur", _i32p),
        ("day_of_tick", _i32p), ("day_feat", _i32p),
        ("n_filters", C.c_int32), ("filter_type", C.c_int32 * 4), ("filter_num", C.c_int32 * 4), ("filter_windows", C.c_int32 * 4),
    ]


BIKE_DEC_HEAD = 8
BIKE_STATION_ATTRS = ("bikes", "capacity", "extra_cost", "failed_return", "fulfillment", "holiday", "id", "min_bikes",
                      "shortage", "temperature", "transfer_cost", "trip_requirement", "weather", "weekday")


def bike_topology_struct(topo):
    s = MaroBikeTopology()
    keep = []
    s.n_stations = topo.n_stations
    s.n_days = len(topo.day_feat)
    s.max_tick = topo.max_tick
    s.resolution = topo.resolution
    s.extra_cost_mode = topo.extra_cost_mode
    s.transfer_seed = topo.transfer_seed
    for name in ("time_mean", "time_std", "supply_ratio", "demand_ratio", "scope_low", "scope_high"):
        setattr(s, name, float(getattr(topo, name)))
    filters = list(getattr(topo, "filters", ()) or ())
    if len(filters) > 4:
        raise ValueError("at most 4 action-scope filters")
    s.n_filters = len(filters)
    for k, (ftype, num, windows) in enumerate(filters):
        s.filter_type[k], s.filter_num[k], s.filter_windows[k] = int(ftype), int(num), int(windows)
    for field, attr in (("station_bikes", "station_bikes"), ("station_capacity", "station_capacity"),
                        ("station_id", "station_id"), ("nbr_offset", "nbr_offset"), ("nbr_idx", "nbr_idx"),
                        ("trip_offset", "trip_offset"), ("trip_src", "trip_src"), ("trip_dst", "trip_dst"),
                        ("trip_dur", "trip_dur"), ("day_of_tick", "day_of_tick"), ("day_feat", "day_feat")):
        a = np.ascontiguousarray(getattr(topo, attr), dtype=np.int32).reshape(-1)
        if a.size == 0:
            a = np.zeros(1, np.int32)
        keep.append(a)
        setattr(s, field, a.ctypes.data_as(_i32p))
    return s, keep


def bike_frame_layout(S: int):
    lay = {"stations": {}, "matrices": {}}
    off = 0
    for a in BIKE_STATION_ATTRS:
        lay["stations"][a] = (off, S, 1)
        off += S
    lay["matrices"]["trips_adj"] = (off, 1, S * S)
    return lay, off + S * S


# ---------------------------------------------------------------------------------------------------- vm_scheduling
_VM_I32 = ("pm_attr", "rack_range", "rack_ids", "cluster_range", "cluster_ids", "dc_range", "dc_ids", "zone_range", "zone_ids",
           "region_range", "vm_attr", "req_offset", "vm_sorted_ids", "vm_sorted_idx", "util_offset", "util_has")
_VM_F64 = ("pm_idle_energy", "pmtype_power", "vm_price", "util_val")


class MaroVmTopology(C.Structure):
    _fields_ = [
        ("n_pm", C.c_int32), ("n_rack", C.c_int32), ("n_cluster", C.c_int32), ("n_dc", C.c_int32), ("n_zone", C.c_int32),
        ("n_region", C.c_int32), ("n_pm_types", C.c_int32), ("n_vm", C.c_int32),
        ("max_tick", C.c_int32), ("delay_duration", C.c_int32), ("buffer_budget", C.c_int32), ("kill_all", C.c_int32),
        ("ticks_per_hour", C.c_double), ("max_cpu_over", C.c_double), ("max_mem_over", C.c_double),
        ("max_util_rate", C.c_double), ("unit_energy_price", C.c_double), ("pue", C.c_double),
        ("pm_attr", _i32p), ("pm_idle_energy", _f64p), ("pmtype_power", _f64p), ("rack_range", _i32p), ("rack_ids", _i32p),
        ("cluster_range", _i32p), ("cluster_ids", _i32p), ("dc_range", _i32p), ("dc_ids", _i32p), ("zone_range", _i32p),
        ("zone_ids", _i32p), ("region_range", _i32p), ("vm_attr", _i32p), ("vm_price", _f64p), ("req_offset", _i32p),
        ("vm_sorted_ids", _i32p), ("vm_sorted_idx", _i32p), ("util_offset", _i32p), ("util_val", _f64p), ("util_has", _i32p),
    ]


VM_DEC_TICK, VM_DEC_VM_ID, VM_DEC_FRAME_INDEX, VM_DEC_CPU, VM_DEC_MEMORY, VM_DEC_SUB_ID = 0, 1, 2, 3, 4, 5
VM_DEC_STATUS, VM_DEC_STEP, VM_DEC_CATEGORY, VM_DEC_BUFFER_TIME, VM_DEC_N_VALID = 6, 7, 8, 9, 10
VM_DEC_HEAD = 12
VM_ACTION_ALLOCATE, VM_ACTION_POSTPONE = 0, 1
VM_METRIC_WORDS = 16
VM_METRIC_NAMES = ("total_vm_requests", "total_incomes", "energy_consumption_cost", "total_profit",
                   "total_energy_consumption", "successful_allocation", "successful_completion", "failed_allocation",
                   "failed_completion", "latency_due_to_agent", "latency_due_to_resource", "total_oversubscriptions",
                   "total_overload_pms", "total_overload_vms")
VM_METRIC_FLOAT = (1, 2, 3, 4)
VM_NODE_ATTRS = {
    "pms": ("cluster_id", "cpu_cores_allocated", "cpu_cores_capacity", "cpu_utilization", "data_center_id",
            "energy_consumption", "id", "memory_allocated", "memory_capacity", "oversubscribable", "pm_type", "rack_id",
            "region_id", "zone_id"),
    "racks": ("cluster_id", "data_center_id", "empty_machine_num", "id", "region_id", "total_machine_num", "zone_id"),
    "clusters": ("data_center_id", "empty_machine_num", "id", "region_id", "total_machine_num", "zone_id"),
    "data_centers": ("empty_machine_num", "id", "region_id", "total_machine_num", "zone_id"),
    "zones": ("empty_machine_num", "id", "region_id", "total_machine_num"),
    "regions": ("empty_machine_num", "id", "total_machine_num"),
}
VM_FLOAT_ATTRS = ("cpu_utilization", "energy_consumption")


def vm_topology_struct(topo):
    s = MaroVmTopology()
    keep = []
    for name in ("n_pm", "n_rack", "n_cluster", "n_dc", "n_zone", "n_region", "max_tick", "delay_duration", "buffer_budget",
                 "kill_all"):
        setattr(s, name, int(getattr(topo, name)))
    s.n_pm_types = len(topo.pmtype_power)
    s.n_vm = topo.n_vm
    for name in ("ticks_per_hour", "max_cpu_over", "max_mem_over", "max_util_rate", "unit_energy_price", "pue"):
        setattr(s, name, float(getattr(topo, name)))
    for name in _VM_I32:
        a = np.ascontiguousarray(getattr(topo, name), dtype=np.int32).reshape(-1)
        if a.size == 0:
            a = np.zeros(1, np.int32)
        keep.append(a)
        setattr(s, name, a.ctypes.data_as(_i32p))
    for name in _VM_F64:
        a = np.ascontiguousarray(getattr(topo, name), dtype=np.float64).reshape(-1)
        if a.size == 0:
            a = np.zeros(1, np.float64)
        keep.append(a)
        setattr(s, name, a.ctypes.data_as(_f64p))
    return s, keep


def vm_frame_layout(topo):
    counts = {"pms": topo.n_pm, "racks": topo.n_rack, "clusters": topo.n_cluster, "data_centers": topo.n_dc,
              "zones": topo.n_zone, "regions": topo.n_region}
    lay, off = {}, 0
    for node, attrs in VM_NODE_ATTRS.items():
        lay[node] = {}
        for a in attrs:
            lay[node][a] = (off, counts[node], 1)
            off += counts[node]
    return lay, off


def vm_metrics_dict(row):
    out = {}
    r = np.asarray(row, np.int64)
    for i, name in enumerate(VM_METRIC_NAMES):
        out[name] = float(r[i:i + 1].view(np.float64)[0]) if i in VM_METRIC_FLOAT else int(r[i])
    return out
