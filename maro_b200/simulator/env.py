"""``Env`` — drop-in for ``maro.simulator.Env`` (maro/simulator/core.py:20-259, abs_core.py:25-176) for the CIM
scenario, backed by one replica (or a view on one replica of a shared batch) of the CUDA core."""
from __future__ import annotations

import os
from enum import IntEnum
from math import ceil, floor
from typing import List, Optional

import numpy as np

from .. import _abi
from ..batch import BikeBatch, CimBatch, VmBatch
from ..scenarios.cim.common import Action, ActionScope, ActionType, DecisionEvent, encode_action
from ..scenarios.cim.topology import CimTopology, build_topology, load_config, next_topology_seed


class DecisionMode(IntEnum):
    Sequential = 0
    Joint = 1


class DocableDict(dict):
    """dict with a docstring, like maro/simulator/scenarios/helpers.py:DocableDict."""

    def __init__(self, doc: str, origin: dict):
        super().__init__(origin)
        self.__doc__ = doc


_METRICS_DOC = """CIM metrics: order_requirements (int), container_shortage (int), operation_number (int)."""


class _CimMetrics(DocableDict):
    __doc__ = _METRICS_DOC

    def __init__(self, a, b, c):  # (built per env per step by VectorEnv: no intermediate dict, no per-instance __doc__)
        dict.__init__(self, order_requirements=a, container_shortage=b, operation_number=c)


def make_metrics(row) -> DocableDict:
    return _CimMetrics(int(row[0]), int(row[1]), int(row[2]))


def parse_query_key(key: slice):
    """``[ticks : nodes : attrs]`` -> (ticks, nodes, attrs) lists; empty = "all", attrs None = no query (frame.pyx:754-801)"""
    ticks = [] if key.start is None else (list(key.start) if isinstance(key.start, (tuple, list, np.ndarray)) else [key.start])
    nodes = [] if key.stop is None else (list(key.stop) if isinstance(key.stop, (tuple, list, np.ndarray)) else [key.stop])
    if key.step is None:
        return ticks, nodes, None
    attrs = list(key.step) if isinstance(key.step, (tuple, list)) else [key.step]
    return ticks, nodes, attrs


class SnapshotNode:
    """``env.snapshot_list["ports"][ticks:nodes:attrs]`` (frame.pyx:734-801) -> 1-D float64 (np_backend.pyx:520-549)."""

    def __init__(self, owner: "SnapshotList", node: str, n_nodes: int):
        self._owner, self._node, self._n = owner, node, n_nodes

    def __len__(self):
        return self._n

    def __getitem__(self, key: slice):
        ticks, nodes, attrs = parse_query_key(key)
        if attrs is None:
            return None
        return self._owner._query(self._node, ticks, nodes, attrs)


class SnapshotList:
    """Read-only façade over the device snapshot ring of one replica (frame.pyx:804-846)."""

    def __init__(self, batch, replica: int):
        self._batch, self._replica = batch, replica
        self._nodes = {name: SnapshotNode(self, name, count) for name, count in batch.node_counts().items()}

    def __getitem__(self, name: str):
        return self._nodes.get(name)

    def __len__(self):
        return len(self.get_frame_index_list())

    def get_frame_index_list(self) -> List[int]:
        return self._batch.snapshot_frames(self._replica).tolist()

    def dump(self, folder: str):
        """``SnapshotList.dump`` (frame.pyx:839-846).  Static layout: what NumpyBackend.dump writes per node type
        (np_backend.pyx:391-401) — ``<node>.npy``, one structured array ``[1 + snapshots][nodes]`` whose row 0 is the live frame
        and rows 1.. the snapshots in ring order, in the node's declared dtypes, and ``<node>.meta`` (attribute names / slot
        counts).  Dynamic layout: ``snapshots_<node>.csv`` like the RawBackend (raw/snapshotlist.cpp:420-487)."""
        if not os.path.exists(folder):
            return
        from .dump import dump_snapshots

        dump_snapshots(self._batch, self._replica, folder)

    def _query(self, node, ticks, nodes, attrs):
        if len(ticks) == 0:
            ticks = self.get_frame_index_list()
        if len(nodes) == 0:
            nodes = list(range(len(self._nodes[node])))
        try:
            ids = [self._batch.attr_id(node, a) for a in attrs]
        except KeyError:
            raise KeyError(f"invalid attribute for {node}: {attrs}")
        dynamic = getattr(self._batch, "query_layout", "static") == "dynamic"
        if len(ticks) == 0:
            return None if dynamic else np.zeros(0, np.float64)  # (_raw_backend_.pyx:301-304: empty result -> None)
        flat = self._batch.query(node, ticks, nodes, ids, [self._replica])[0]
        if dynamic:  # RawBackend: 4-D (ticks, nodes, attrs, max_slots), NaN padded (_raw_backend_.pyx:306-315)
            return flat.reshape(self._batch.query_shape(node, ids, len(ticks), len(nodes)))
        return flat


class _NodeView:
    def __init__(self, values: dict, index: int):
        self.__dict__["_v"], self.__dict__["index"] = values, index

    def __getattr__(self, name):
        v = self._v[name]
        return v.item() if v.shape == () else v


class FrameView:
    """Read-only copy of the live frame (``env.current_frame``): ``.ports[i].empty``, ``.vessels[i].full`` …"""

    def __init__(self, words: np.ndarray, topo: CimTopology):
        lay, _ = _abi.frame_layout(topo.n_ports, topo.n_vessels, topo.past_stop_number, topo.future_stop_number)
        self.ports, self.vessels = [], []
        for node, target, n in (("ports", self.ports, topo.n_ports), ("vessels", self.vessels, topo.n_vessels)):
            for i in range(n):
                vals = {}
                for a, (off, _, slots) in lay[node].items():
                    w = words[off + i * slots: off + (i + 1) * slots]
                    if a == "transfer_cost":
                        w = w.view(np.float32)
                    vals[a] = w[0] if slots == 1 else w.copy()
                target.append(_NodeView(vals, i))
        self.matrix = [{a: words[off: off + slots].copy() for a, (off, _, slots) in lay["matrices"].items()}]


class GenericFrameView:
    """``env.current_frame`` for the scenarios whose nodes only have single-slot attributes plus matrices: attribute
    access by node list, e.g. ``.stations[i].bikes`` / ``.pms[i].cpu_cores_allocated``, ``.matrices[0].trips_adj``."""

    def __init__(self, words: np.ndarray, layout: dict, float_attrs=()):
        for node, attrs in layout.items():
            n = next(iter(attrs.values()))[1]
            views = []
            for i in range(n):
                vals = {}
                for a, (off, _, slots) in attrs.items():
                    w = words[off + i * slots: off + (i + 1) * slots]
                    if a in float_attrs:
                        w = w.view(np.float32)
                    vals[a] = w[0] if slots == 1 else w.copy()
                views.append(_NodeView(vals, i))
            setattr(self, node, views)


class Env:
    """Same constructor and members as ``maro.simulator.Env`` for the cim, citi_bike and vm_scheduling scenarios."""

    def __init__(self, scenario: str = None, topology: str = None, start_tick: int = 0, durations: int = 100,
                 snapshot_resolution: int = 1, max_snapshots: int = None, decision_mode=DecisionMode.Sequential,
                 business_engine_cls: type = None, disable_finished_events: bool = False,
                 record_finished_events: bool = False, record_file_path: str = None, options: Optional[dict] = None,
                 device: int = 0):
        if scenario not in ("cim", "citi_bike", "vm_scheduling"):
            raise NotImplementedError(f"scenario {scenario!r}: 'cim', 'citi_bike' and 'vm_scheduling' run on the CUDA core")
        if business_engine_cls is not None:
            raise NotImplementedError("custom business engines run on the reference Env, not on the CUDA core")
        self._joint = int(decision_mode) == int(DecisionMode.Joint)
        if self._joint and scenario != "cim":
            raise NotImplementedError("DecisionMode.Joint runs on the CUDA core for the cim scenario")
        self._scenario, self._topology = scenario, topology
        self._start_tick, self._durations = start_tick, durations
        self._snapshot_resolution, self._max_snapshots = snapshot_resolution, max_snapshots
        self._device = device
        self._name = f"{scenario}:{topology}"
        self._pending_seed: Optional[int] = None
        if scenario == "citi_bike":
            from ..scenarios.citi_bike.data import build_bike_topology, load_bike_config

            self._config = load_bike_config(topology)
            self._topo = build_bike_topology(self._config, start_tick, start_tick + durations,
                                             transfer_seed=int((options or {}).get("transfer_seed", 0)))
            self._batch = BikeBatch(self._topo, 1, snapshot_resolution, max_snapshots, device=device, max_actions=8)
        elif scenario == "vm_scheduling":
            from ..scenarios.vm_scheduling.data import build_vm_topology, load_vm_config

            self._config = load_vm_config(topology)
            self._topo = build_vm_topology(self._config, start_tick, start_tick + durations)
            self._batch = VmBatch(self._topo, 1, snapshot_resolution, max_snapshots, device=device, max_actions=8)
        else:
            self._config = load_config(topology)
            self._topo = build_topology(self._config, start_tick + durations)
            self._batch = CimBatch(self._topo, 1, start_tick, snapshot_resolution, max_snapshots, device=device,
                                   max_actions=8, decision_mode=int(self._joint))
        # the reference picks its backend per process from DEFAULT_BACKEND_NAME (maro/backends/frame.pyx:496-504); the two
        # answer snapshot queries in different layouts — same switch here, per Env
        self._backend_name = str((options or {}).get("backend_name") or os.environ.get("DEFAULT_BACKEND_NAME", "static"))
        if self._backend_name == "dynamic":
            self._batch.set_query_layout("dynamic")
        self._snapshots = SnapshotList(self._batch, 0)
        self._tick = start_tick
        self._last_metrics = make_metrics((0, 0, 0))
        self._act = np.zeros((1, self._batch.max_actions, 4), np.int32)
        self._nact = np.zeros(1, np.int32)

    def _step_joint(self, action):
        """DecisionMode.Joint (core.py:354-366): ``action`` = None or a list with one entry (Action or None) per decision event
        the previous step returned, in the same order (a shorter list leaves the rest unanswered); returns the LIST of this
        tick's decision events."""
        answers = [] if action is None else (action if isinstance(action, list) else [action])
        if len(answers) > self._act.shape[1]:
            raise ValueError("more answers than decision events")
        for k, a in enumerate(answers):
            if a is None:
                self._act[0, k] = (0, 0, 0, 2)
            elif isinstance(a, list):
                raise NotImplementedError("Joint mode on the CUDA core takes one Action (or None) per decision event")
            else:
                encode_action(a, self._act[0, k])
        self._nact[0] = len(answers)
        dec, met = self._batch.step(self._act, self._nact)
        rows = dec[0].reshape(-1, _abi.DECISION_WORDS)
        status = int(rows[0, _abi.DEC_STATUS])
        if status == _abi.STATUS_BAD_ACTION:
            raise AssertionError("invalid action: quantity exceeds the action scope (business_engine.py:731,736)")
        if status == _abi.STATUS_QUEUE_OVERFLOW:
            raise RuntimeError("event queue overflow: recreate the Env with a larger queue_capacity")
        if status == _abi.STATUS_FINISHED:
            return None, None, True
        self._tick = int(rows[0, _abi.DEC_TICK])
        self._last_metrics = make_metrics(met[0])
        if status == _abi.STATUS_DONE:
            return self._last_metrics, None, True
        events = []
        for d in rows:
            if int(d[_abi.DEC_STATUS]) != _abi.STATUS_DECISION:
                break
            events.append(DecisionEvent(int(d[0]), int(d[1]), int(d[2]), self._snapshots, ActionScope(int(d[3]), int(d[4])), int(d[5])))
        return self._last_metrics, events, False

    # ---- stepping (core.py:92-133)
    def step(self, action=None):
        if self._joint:
            return self._step_joint(action)
        if action is None:
            actions = []
        elif not isinstance(action, list):
            actions = [action]
        else:
            actions = action
        if len(actions) > self._act.shape[1]:
            raise ValueError("too many actions for one decision event")
        for i, a in enumerate(actions):
            if self._scenario == "citi_bike":
                from ..scenarios.citi_bike.common import encode_bike_action

                encode_bike_action(a, self._act[0, i])
            elif self._scenario == "vm_scheduling":
                from ..scenarios.vm_scheduling.common import encode_vm_action

                encode_vm_action(a, self._act[0, i])
            else:
                encode_action(a, self._act[0, i])
        self._nact[0] = len(actions)
        dec, met = self._batch.step(self._act, self._nact)
        d = dec[0]
        status = int(d[_abi.DEC_STATUS])
        if status == _abi.STATUS_BAD_ACTION:
            if self._scenario == "vm_scheduling":
                raise Exception("The VM id or PM id sent by agent is invalid. (vm_scheduling/business_engine.py:842)")
            raise AssertionError("invalid action: quantity exceeds the action scope (business_engine.py:731,736)")
        if status == _abi.STATUS_QUEUE_OVERFLOW:
            raise RuntimeError("event queue overflow: recreate the Env with a larger queue_capacity")
        if status == _abi.STATUS_FINISHED:
            return None, None, True
        self._tick = int(d[_abi.DEC_TICK])
        if self._scenario == "citi_bike":
            from ..scenarios.citi_bike.common import decode_bike_decision

            self._last_metrics = DocableDict("citi_bike metrics", {"trip_requirements": int(met[0][0]),
                                                                   "bike_shortage": int(met[0][1]),
                                                                   "operation_number": int(met[0][2])})
            if status == _abi.STATUS_DONE:
                return self._last_metrics, None, True
            return self._last_metrics, decode_bike_decision(d, self._snapshots), False
        if self._scenario == "vm_scheduling":
            from ..scenarios.vm_scheduling.common import decode_vm_decision, decode_vm_metrics

            self._last_metrics = DocableDict("vm_scheduling metrics", decode_vm_metrics(met[0]))
            if status == _abi.STATUS_DONE:
                return self._last_metrics, None, True
            return self._last_metrics, decode_vm_decision(d), False
        self._last_metrics = make_metrics(met[0])
        if status == _abi.STATUS_DONE:
            return self._last_metrics, None, True
        event = DecisionEvent(int(d[0]), int(d[1]), int(d[2]), self._snapshots, ActionScope(int(d[3]), int(d[4])), int(d[5]))
        return self._last_metrics, event, False

    def dump(self, path: Optional[str] = None, with_snapshots: bool = True) -> None:
        """``Env.dump`` — "Dump environment for restore", a no-op in the reference (core.py:135-141).  With a ``path`` the
        complete device state of the environment (frame, event queue, snapshot ring, RNG streams, topology tables) is written
        to that file; ``restore(path)`` on an Env built with the same arguments continues the episode bit for bit.  Without
        a path it stays the reference's no-op."""
        if path is not None:
            self._batch.save(path, with_snapshots)

    def restore(self, path: str) -> None:
        """Load a checkpoint written by ``dump(path)``; the next ``step`` continues from where the dump was taken (it expects
        the action for the decision that was pending then, exactly like the ``step`` that would have followed)."""
        self._batch.load(path)
        self._tick = int(self._batch.ticks()[0])

    def reset(self, keep_seed: bool = False) -> None:
        """core.py:143-170 + cim_data_container_helpers.py:56-66: ``keep_seed=False`` draws a new topology seed from
        the route_init stream; a seed set with ``set_seed`` takes effect here."""
        if self._scenario in ("citi_bike", "vm_scheduling"):  # their set_seed is a no-op (citi_bike/business_engine.py:198-199)
            self._batch.reset()
            self._tick = self._start_tick
            return
        seed = self._pending_seed
        if not keep_seed:
            seed = next_topology_seed(self._topo)
        if seed is not None:
            self._topo = build_topology(self._config, self._start_tick + self._durations, seed=seed)
            try:
                self._batch.set_topology(0, self._topo)
            except RuntimeError:
                # the new instance needs a longer event horizon / larger pool than the handle was sized for
                # (maro_cim_set_topology refuses instead of aliasing buckets): a fresh handle, like the reference's
                # reset rebuilds its data container (cim_data_container_helpers.py:56-70)
                self._batch.close()
                self._batch = CimBatch(self._topo, 1, self._start_tick, self._snapshot_resolution, self._max_snapshots,
                                       device=self._device, max_actions=8, decision_mode=int(self._joint))
                if self._backend_name == "dynamic":
                    self._batch.set_query_layout("dynamic")
                self._snapshots = SnapshotList(self._batch, 0)
            self._pending_seed = None
        self._batch.reset()
        self._tick = self._start_tick
        self._last_metrics = make_metrics((0, 0, 0))

    def set_seed(self, seed: int) -> None:
        assert seed is not None and isinstance(seed, int)
        self._pending_seed = seed

    # ---- properties (core.py:172-259)
    @property
    def configs(self) -> dict:
        return self._config

    @property
    def summary(self) -> dict:
        if self._scenario == "citi_bike":  # citi_bike/business_engine.py:149-163
            lay, _ = _abi.bike_frame_layout(self._topo.n_stations)
            from ..scenarios.citi_bike.common import DecisionEvent as BikeDecisionEvent

            return {"node_mapping": {i: int(sid) for i, sid in enumerate(self._topo.station_id)},
                    "node_detail": {n: {"number": next(iter(a.values()))[1], "attributes": {k: {"slots": v[2]} for k, v in a.items()}}
                                    for n, a in lay.items()},
                    "event_payload": {"RequireBike": ["timestamp", "durations", "src_station", "dest_station"],
                                      "ReturnBike": ["from_station_idx", "to_station_idx", "number"],
                                      "RebalanceBike": BikeDecisionEvent.summary_key,
                                      "DeliverBike": ["from_station_idx", "to_station_idx", "number"]}}
        if self._scenario == "vm_scheduling":  # vm_scheduling/business_engine.py:527-545
            lay, _ = _abi.vm_frame_layout(self._topo)
            from ..scenarios.vm_scheduling.common import DecisionEvent as VmDecisionEvent

            return {"node_mapping": {},
                    "node_detail": {n: {"number": next(iter(a.values()))[1], "attributes": {k: {"slots": v[2]} for k, v in a.items()}}
                                    for n, a in lay.items()},
                    "event_payload": {"REQUEST": ["vm_info", "remaining_buffer_time"], "PENDING_DECISION": VmDecisionEvent.summary_key}}
        t = self._topo
        lay, _ = _abi.frame_layout(t.n_ports, t.n_vessels, t.past_stop_number, t.future_stop_number)
        detail = {}
        for node, n in (("ports", t.n_ports), ("vessels", t.n_vessels), ("matrices", 1)):
            detail[node] = {"number": n, "attributes": {a: {"slots": slots} for a, (_, _, slots) in lay[node].items()}}
        return {
            "node_mapping": {"ports": {n: i for i, n in enumerate(t.port_names)},
                             "vessels": {n: i for i, n in enumerate(t.vessel_names)}},
            "node_detail": detail,
            "event_payload": {
                "ORDER": ["tick", "src_port_idx", "dest_port_idx", "quantity"],
                "RETURN_FULL": ["src_port_idx", "dest_port_idx", "quantity"],
                "VESSEL_ARRIVAL": ["port_idx", "vessel_idx"], "LOAD_FULL": ["port_idx", "vessel_idx"],
                "DISCHARGE_FULL": ["vessel_idx", "port_idx", "from_port_idx", "quantity"],
                "PENDING_DECISION": DecisionEvent.summary_key, "LOAD_EMPTY": Action.summary_key,
                "DISCHARGE_EMPTY": Action.summary_key, "VESSEL_DEPARTURE": ["port_idx", "vessel_idx"],
                "RETURN_EMPTY": ["port_idx", "quantity"]},
        }

    @property
    def name(self) -> str:
        return self._name

    @property
    def current_frame(self):
        if self._scenario == "citi_bike":
            return GenericFrameView(self._batch.read_frame(0), _abi.bike_frame_layout(self._topo.n_stations)[0])
        if self._scenario == "vm_scheduling":
            return GenericFrameView(self._batch.read_frame(0), _abi.vm_frame_layout(self._topo)[0], _abi.VM_FLOAT_ATTRS)
        return FrameView(self._batch.read_frame(0), self._topo)

    @property
    def tick(self) -> int:
        return self._tick

    @property
    def frame_index(self) -> int:
        return floor((self._tick - self._start_tick) / self._snapshot_resolution)

    @property
    def snapshot_list(self) -> SnapshotList:
        return self._snapshots

    @property
    def agent_idx_list(self) -> Optional[List[int]]:
        if self._scenario == "vm_scheduling":
            return None  # the reference's get_agent_idx_list is a docstring only (vm_scheduling/business_engine.py:534-535)
        return list(range(self._topo.n_stations if self._scenario == "citi_bike" else self._topo.n_ports))

    @property
    def metrics(self) -> dict:
        return self._last_metrics

    def get_finished_events(self) -> list:
        """core.py:241-249.  Event objects are never materialised on the device (the per-tick execution order
        is reconstructed from its sources), so there is no finished-event list to hand out; an empty list would be
        indistinguishable from "no events", hence the loud failure.  The executed-event COUNT is available
        (``env.batch.counters()[:, 2]``)."""
        raise NotImplementedError("finished-event objects are not materialised by the CUDA core "
                                  "(event counts: Env.batch.counters()); run the reference Env for event-level inspection")

    def get_pending_events(self, tick) -> list:
        """core.py:251-259 — see ``get_finished_events``."""
        raise NotImplementedError("pending-event objects are not materialised by the CUDA core")

    @property
    def batch(self):
        """the columnar batch (one replica) behind this Env"""
        return self._batch

    def get_ticks_frame_index_mapping(self) -> dict:
        mapping = {}
        res = self._snapshot_resolution
        for f in self._snapshots.get_frame_index_list():
            lo = self._start_tick + f * res
            for t in range(lo, min(lo + res, self._start_tick + self._durations)):
                mapping[t] = f
        return mapping

    def close(self):
        self._batch.close()
