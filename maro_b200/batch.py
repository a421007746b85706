"""``CimBatch`` — thin columnar wrapper over the C ABI: B CIM replicas resident on one GPU.

This is the layer both drop-in surfaces (``maro_b200.simulator.Env`` and ``maro_b200.vector_env.VectorEnv``)
sit on.  All arrays are numpy (host) unless a method says ``_device``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _abi, _native
from .scenarios.cim.topology import CimTopology

def _pinned_views(lib_fn, handle, B, A, dec_words, met_words=3):
    """numpy views over the library's pinned staging buffers (no copies)."""
    ptrs = [C.c_void_p() for _ in range(5)]
    _native.check(lib_fn(handle, *[C.byref(p) for p in ptrs]))

    def view(p, nbytes, dtype, shape):
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    return (view(ptrs[0], B * A * 16, np.int32, (B, A, 4)), view(ptrs[1], B * 4, np.int32, (B,)),
            view(ptrs[2], B, np.uint8, (B,)), view(ptrs[3], B * dec_words * 4, np.int32, (B, dec_words)),
            view(ptrs[4], B * met_words * 8, np.int64, (B, met_words)))


_NODE_TYPE = {"ports": _abi.NODE_PORTS, "vessels": _abi.NODE_VESSELS, "matrices": _abi.NODE_MATRICES}


class CimBatch:
    def __init__(self, topologies, n_replicas: int, start_tick: int = 0, snapshot_resolution: int = 1,
                 max_snapshots: Optional[int] = None, device: int = 0, max_actions: int = 1,
                 replica_topology: Optional[Sequence[int]] = None, queue_capacity: int = 0, decision_mode: int = 0):
        if isinstance(topologies, CimTopology):
            topologies = [topologies]
        self.topologies = list(topologies)
        self.n_replicas = int(n_replicas)
        self.decision_mode = int(decision_mode)  # 0 Sequential, 1 Joint (core.py:354-366)
        # Joint: one answer row per decision event of a tick (at most one per vessel), V decision rows per replica
        self.max_actions = max(int(max_actions), self.topologies[0].n_vessels) if self.decision_mode == 1 else int(max_actions)
        self.dec_words = _abi.DECISION_WORDS * (self.topologies[0].n_vessels if self.decision_mode == 1 else 1)
        self.start_tick = int(start_tick)
        self.snapshot_resolution = int(snapshot_resolution)
        self.device = int(device)
        L = _native.lib()
        self._keep = []
        arr = (_abi.MaroCimTopology * len(self.topologies))()
        for i, t in enumerate(self.topologies):
            s, keep = _abi.topology_struct(t)
            arr[i] = s
            self._keep.append(keep)
        cfg = _abi.MaroCimConfig()
        cfg.n_replicas = self.n_replicas
        cfg.start_tick = self.start_tick
        cfg.snapshot_resolution = self.snapshot_resolution
        cfg.max_snapshots = int(max_snapshots) if max_snapshots else 0
        self._max_snapshots = cfg.max_snapshots
        cfg.device = self.device
        cfg.queue_capacity = int(queue_capacity)
        cfg.max_actions = self.max_actions
        cfg.decision_mode = self.decision_mode
        if replica_topology is not None:
            rt = np.ascontiguousarray(replica_topology, np.int32)
            assert rt.shape == (self.n_replicas,)
            self._keep.append(rt)
            cfg.replica_topology = rt.ctypes.data_as(C.POINTER(C.c_int32))
        h = C.c_void_p()
        _native.check(L.maro_cim_create(arr, len(self.topologies), C.byref(cfg), C.byref(h)))
        self._h = h
        self.frame_words = L.maro_cim_frame_words(self._h)
        # reusable host output buffers
        self.decisions = np.zeros((self.n_replicas, self.dec_words), np.int32)
        self.metrics = np.zeros((self.n_replicas, _abi.METRIC_WORDS), np.int64)

    # -- lifecycle ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            _native.lib().maro_cim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        """Run on an external CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``; 0 = legacy default
        stream); ``None`` returns to the library's own stream."""
        if cuda_stream_ptr is None:
            _native.check(_native.lib().maro_cim_set_stream(self._h, None, 0))
        else:
            _native.check(_native.lib().maro_cim_set_stream(self._h, C.c_void_p(cuda_stream_ptr), 1))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        _native.check(_native.lib().maro_cim_reset(self._h, None if m is None else m.ctypes.data))

    def set_topology(self, index: int, topo: CimTopology):
        s, keep = _abi.topology_struct(topo)
        _native.check(_native.lib().maro_cim_set_topology(self._h, index, C.byref(s)))
        self.topologies[index] = topo

    # -- stepping ----------------------------------------------------------------------------------
    def step(self, actions=None, n_actions=None, active=None):
        """One Env.step for every (active) replica.  actions: int32 [B][max_actions][4] or None."""
        a = n = m = None
        if actions is not None:
            a = np.ascontiguousarray(actions, np.int32)
            assert a.size == self.n_replicas * self.max_actions * 4, a.shape
        if n_actions is not None:
            n = np.ascontiguousarray(n_actions, np.int32)
        if active is not None:
            m = np.ascontiguousarray(active, np.uint8)
        _native.check(_native.lib().maro_cim_step(
            self._h, None if m is None else m.ctypes.data, None if a is None else a.ctypes.data,
            None if n is None else n.ctypes.data, self.decisions.ctypes.data, self.metrics.ctypes.data))
        return self.decisions, self.metrics

    def pinned(self):
        """(actions, n_actions, active, decisions, metrics) numpy views over the library's pinned staging buffers; fill
        the inputs in place, call ``step_pinned`` and read the outputs in place — no host-side copies at all."""
        if getattr(self, "_pinned", None) is None:
            self._pinned = _pinned_views(_native.lib().maro_cim_pinned_buffers, self._h, self.n_replicas,
                                         self.max_actions, self.dec_words)
        return self._pinned

    def step_pinned(self, use_actions: bool = True, use_n_actions: bool = False, use_active: bool = False):
        _native.check(_native.lib().maro_cim_step_pinned(self._h, int(use_actions), int(use_n_actions), int(use_active)))

    def pinned_granularity(self) -> int:
        """replicas per independently steppable block of ``submit_pinned`` / ``wait_pinned`` (0: not available)"""
        return _native.lib().maro_cim_pinned_granularity(self._h)

    def submit_pinned(self, first: int, count: int, use_actions: bool = True, use_n_actions: bool = False, use_active: bool = False):
        """asynchronous half of ``step_pinned`` for the replicas [first, first + count)"""
        _native.check(_native.lib().maro_cim_submit_pinned(self._h, first, count, int(use_actions), int(use_n_actions), int(use_active)))

    def wait_pinned(self, first: int, count: int):
        _native.check(_native.lib().maro_cim_wait_pinned(self._h, first, count))

    def step_device(self, d_decisions: int, d_metrics: int, d_actions: int = 0, d_n_actions: int = 0, d_active: int = 0):
        """Asynchronous step on device pointers (ints, e.g. ``tensor.data_ptr()``)."""
        _native.check(_native.lib().maro_cim_step_device(self._h, d_active or None, d_actions or None,
                                                         d_n_actions or None, d_decisions, d_metrics))

    def rollout_device(self, d_decisions: int, d_metrics: int, n_steps: int, policy: int = 1, seed: int = 0,
                       replica_base: int = 0, d_trace: int = 0):
        """Resident mode: ``n_steps`` fused env-steps per replica in one launch, agent on the device (policy 0 = None
        actions, 1 = hashed hello-world agent).  ``d_decisions`` is in/out; ``d_trace`` ([n_steps][B][8] int32) optional."""
        _native.check(_native.lib().maro_cim_rollout_device(self._h, policy, seed, replica_base, n_steps, d_decisions,
                                                            d_metrics, d_trace or None))

    def random_policy_device(self, d_decisions: int, d_actions: int, seed: int, replica_base: int = 0):
        _native.check(_native.lib().maro_cim_random_policy_device(self._h, d_decisions, d_actions, seed, replica_base))

    # -- RL shaping on the device snapshot ring (examples/cim/rl/env_sampler.py:15-36, 66-80) ---------
    def rl_state_dim(self, look_back: int, n_port_attrs: int, n_vessel_attrs: int) -> int:
        return _native.lib().maro_cim_rl_state_dim(self._h, look_back, n_port_attrs, n_vessel_attrs)

    def rl_state_device(self, d_decisions: int, look_back: int, port_attrs, vessel_attrs, d_out: int, f32: bool = False):
        """d_out: [n_replicas][rl_state_dim] float64 (or float32 with ``f32``) device buffer"""
        pa = np.ascontiguousarray([a if isinstance(a, (int, np.integer)) else self.attr_id("ports", a) for a in port_attrs], np.int32)
        va = np.ascontiguousarray([a if isinstance(a, (int, np.integer)) else self.attr_id("vessels", a) for a in vessel_attrs], np.int32)
        fn = _native.lib().maro_cim_rl_state_f32_device if f32 else _native.lib().maro_cim_rl_state_device
        _native.check(fn(self._h, d_decisions, look_back, pa.ctypes.data, len(pa), va.ctypes.data, len(va), d_out))

    def rl_action_device(self, d_decisions: int, d_model_actions: int, d_action_space: int, n_action_space: int,
                         finite_vessel_space: bool, has_early_discharge: bool, d_actions: int):
        _native.check(_native.lib().maro_cim_rl_action_device(self._h, d_decisions, d_model_actions, d_action_space,
                                                              n_action_space, int(finite_vessel_space),
                                                              int(has_early_discharge), d_actions))

    def rl_action_ex_device(self, d_decisions: int, d_model_actions: int, model_actions_are_i64: bool, d_record: int, d_metrics_in: int,
                            d_metrics_final: int, d_action_space: int, n_action_space: int, finite_vessel_space: bool,
                            has_early_discharge: bool, d_actions: int):
        """maro_cim_rl_action_ex_device: the translation + the sampler's bookkeeping (int64 policy output, int32 record, running
        maximum of the metrics) in one launch; 0 for a pointer that is not wanted"""
        _native.check(_native.lib().maro_cim_rl_action_ex_device(self._h, d_decisions, d_model_actions, int(model_actions_are_i64),
                                                                 d_record or None, d_metrics_in or None, d_metrics_final or None,
                                                                 d_action_space, n_action_space, int(finite_vessel_space),
                                                                 int(has_early_discharge), d_actions))

    def rl_reward_device(self, d_ticks: int, d_ports: int, d_decay: int, time_window: int, fulfillment_factor: float,
                         shortage_factor: float, d_out: int):
        _native.check(_native.lib().maro_cim_rl_reward_device(self._h, d_ticks, d_ports, d_decay, time_window,
                                                              float(fulfillment_factor), float(shortage_factor), d_out))

    def rl_reward_batch_device(self, d_ticks: int, d_ports: int, n_rows: int, d_decay: int, time_window: int,
                               fulfillment_factor: float, shortage_factor: float, d_out: int):
        """rewards of [n_rows][B] (tick, port) pairs in one launch"""
        _native.check(_native.lib().maro_cim_rl_reward_batch_device(self._h, d_ticks, d_ports, n_rows, d_decay, time_window,
                                                                    float(fulfillment_factor), float(shortage_factor), d_out))

    # -- inspection --------------------------------------------------------------------------------
    def node_counts(self) -> dict:
        """node type name -> number of nodes (what ``len(env.snapshot_list[name])`` reports)"""
        t = self.topologies[0]
        return {"ports": t.n_ports, "vessels": t.n_vessels, "matrices": 1}

    def attr_id(self, node: str, name: str) -> int:
        i = _native.lib().maro_cim_attr_id(self._h, _NODE_TYPE[node], name.encode())
        if i < 0:
            raise KeyError(f"{node}.{name}")
        return i

    def attr_slots(self, node: str, attr_id: int) -> int:
        return _native.lib().maro_cim_attr_slots(self._h, _NODE_TYPE[node], attr_id)

    def ring_rows(self) -> int:
        """rows of the snapshot ring: ``max_snapshots`` or the number of frames of the episode (np_backend.pyx:481-518)"""
        total = -(-(self.topologies[0].max_tick - self.start_tick) // self.snapshot_resolution)
        return min(self._max_snapshots, total) if self._max_snapshots else total

    def save(self, path: str, with_snapshots: bool = True):
        """Device-state checkpoint of the whole batch (replica blocks, event queues, snapshot ring, RNG streams, topology
        tables) into one file — what ``Env.dump`` promises and the reference leaves unimplemented (core.py:135-141)."""
        _native.check(_native.lib().maro_cim_save(self._h, os.fsencode(path), int(with_snapshots)))

    def load(self, path: str):
        """Restore a checkpoint written by ``save`` on a batch of the same topologies / configuration."""
        _native.check(_native.lib().maro_cim_load(self._h, os.fsencode(path)))

    query_layout = "static"

    def set_query_layout(self, layout: str):
        """"static" (NumpyBackend conventions, the reference's default) or "dynamic" (RawBackend conventions: every
        attribute padded to the widest one's slots, NaN for missing slots / unknown frames, values through float32) —
        the reference's DEFAULT_BACKEND_NAME choice (maro/backends/frame.pyx:496-504), per handle."""
        assert layout in ("static", "dynamic")
        _native.check(_native.lib().maro_cim_set_query_layout(self._h, 1 if layout == "dynamic" else 0))
        self.query_layout = layout

    def _per_replica(self, node, at, n_frames, n_nodes) -> int:
        slots = [self.attr_slots(node, int(a)) for a in at]
        per_node = len(slots) * max(slots) if self.query_layout == "dynamic" else sum(slots)
        return per_node * n_frames * n_nodes

    def query_shape(self, node, attrs, n_frames, n_nodes):
        """shape of one replica's result under the dynamic layout: (frames, nodes, attrs, max_slots)"""
        slots = [self.attr_slots(node, a if isinstance(a, (int, np.integer)) else self.attr_id(node, a)) for a in attrs]
        return (n_frames, n_nodes, len(slots), max(slots))

    def query(self, node: str, frame_indices, nodes, attrs, replicas=None) -> np.ndarray:
        """float64 [n_replicas_queried, per_replica] in tick -> node -> attr -> slot order (np_backend.pyx:520-549)."""
        reps = np.arange(self.n_replicas, dtype=np.int32) if replicas is None else np.ascontiguousarray(replicas, np.int32)
        fr = np.ascontiguousarray(frame_indices, np.int32)
        nd = np.ascontiguousarray(nodes, np.int32)
        at = np.ascontiguousarray([a if isinstance(a, (int, np.integer)) else self.attr_id(node, a) for a in attrs], np.int32)
        per = self._per_replica(node, at, len(fr), len(nd))
        out = np.zeros((len(reps), per), np.float64)
        pr = C.c_int64()
        _native.check(_native.lib().maro_cim_query(self._h, reps.ctypes.data, len(reps), _NODE_TYPE[node], fr.ctypes.data,
                                                   len(fr), nd.ctypes.data, len(nd), at.ctypes.data, len(at),
                                                   out.ctypes.data, C.byref(pr)))
        assert pr.value == per
        return out

    def query_device(self, node: str, frame_indices, nodes, attrs, replicas=None):
        """Same gather as ``query`` but the result stays in HBM: returns a torch float64 CUDA tensor
        [n_replicas_queried, per_replica] (state / reward shaping without a host round trip)."""
        import torch

        reps = np.arange(self.n_replicas, dtype=np.int32) if replicas is None else np.ascontiguousarray(replicas, np.int32)
        fr = np.ascontiguousarray(frame_indices, np.int32)
        nd = np.ascontiguousarray(nodes, np.int32)
        at = np.ascontiguousarray([a if isinstance(a, (int, np.integer)) else self.attr_id(node, a) for a in attrs], np.int32)
        per = self._per_replica(node, at, len(fr), len(nd))
        out = torch.empty((len(reps), per), dtype=torch.float64, device=f"cuda:{self.device}")
        pr = C.c_int64()
        _native.check(_native.lib().maro_cim_query_device(self._h, reps.ctypes.data, len(reps), _NODE_TYPE[node],
                                                          fr.ctypes.data, len(fr), nd.ctypes.data, len(nd), at.ctypes.data,
                                                          len(at), out.data_ptr(), C.byref(pr)))
        assert pr.value == per
        return out

    def read_frame(self, replica: int = 0) -> np.ndarray:
        out = np.zeros(self.frame_words, np.int32)
        _native.check(_native.lib().maro_cim_read_frame(self._h, replica, out.ctypes.data, out.size))
        return out

    def ticks(self) -> np.ndarray:
        out = np.zeros(self.n_replicas, np.int32)
        _native.check(_native.lib().maro_cim_ticks(self._h, out.ctypes.data))
        return out

    def counters(self) -> np.ndarray:
        """int64 [B][4] cumulative {env_steps, ticks, events, snapshots}."""
        out = np.zeros((self.n_replicas, 4), np.int64)
        _native.check(_native.lib().maro_cim_counters(self._h, out.ctypes.data))
        return out

    def snapshot_frames(self, replica: int = 0) -> np.ndarray:
        cap = 1 << 16
        out = np.zeros(cap, np.int32)
        n = C.c_int32()
        _native.check(_native.lib().maro_cim_snapshot_frames(self._h, replica, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def snapshot_row(self, frame_index: int, replica: int = 0):
        """Raw frame words of one snapshot (None if the ring does not hold it) — assembled from queries."""
        if frame_index not in set(self.snapshot_frames(replica).tolist()):
            return None
        lay, fw = _abi.frame_layout(self.topologies[0].n_ports, self.topologies[0].n_vessels,
                                    self.topologies[0].past_stop_number, self.topologies[0].future_stop_number)
        row = np.zeros(fw, np.int32)
        for node, attrs in lay.items():
            n_nodes = next(iter(attrs.values()))[1]
            names = list(attrs)
            vals = self.query(node, [frame_index], np.arange(n_nodes), names, [replica])[0]
            pos = 0
            per_node = sum(attrs[a][2] for a in names)
            for nd in range(n_nodes):
                base = nd * per_node
                p = 0
                for a in names:
                    off, _, slots = attrs[a]
                    chunk = vals[base + p: base + p + slots]
                    if node == "ports" and a == "transfer_cost":
                        row[off + nd * slots: off + (nd + 1) * slots] = chunk.astype(np.float32).view(np.int32)
                    else:
                        row[off + nd * slots: off + (nd + 1) * slots] = chunk.astype(np.int64).astype(np.int32)
                    p += slots
            pos += 1
        return row


class BikeBatch:
    """Columnar wrapper over the citi_bike entry points of the C ABI (same shape as ``CimBatch``)."""

    _NODE = {"stations": 0, "matrices": 1}
    _PREFIX = "maro_bike"
    _MET_WORDS = 3

    def _f(self, name):
        return getattr(_native.lib(), f"{self._PREFIX}_{name}")

    def _topology_struct(self, topology):
        return _abi.bike_topology_struct(topology)

    def node_counts(self) -> dict:
        return {"stations": self.topology.n_stations, "matrices": 1}

    def ring_rows(self) -> int:
        total = -(-(int(self.topology.max_tick) - int(self.topology.start_tick)) // self.snapshot_resolution)
        return min(self._max_snapshots, total) if self._max_snapshots else total

    def save(self, path: str, with_snapshots: bool = True):
        """see ``CimBatch.save``"""
        _native.check(self._f("save")(self._h, os.fsencode(path), int(with_snapshots)))

    def load(self, path: str):
        _native.check(self._f("load")(self._h, os.fsencode(path)))

    query_layout = "static"
    _per_replica = CimBatch._per_replica
    query_shape = CimBatch.query_shape

    def set_query_layout(self, layout: str):
        """see ``CimBatch.set_query_layout``"""
        assert layout in ("static", "dynamic")
        _native.check(self._f("set_query_layout")(self._h, 1 if layout == "dynamic" else 0))
        self.query_layout = layout

    def set_transfer_seeds(self, seeds):
        """Per-replica ``np.random.seed`` of the transfer_time stream (every env of the reference's VectorEnv is its own
        process with its own numpy RandomState); uint32 [n_replicas] or None (= the topology's transfer_seed for all).
        Takes effect at each replica's next ``reset``."""
        if seeds is None:
            _native.check(self._f("set_transfer_seeds")(self._h, None))
            return
        sd = np.ascontiguousarray(seeds, np.uint32)
        assert sd.shape == (self.n_replicas,)
        _native.check(self._f("set_transfer_seeds")(self._h, sd.ctypes.data))

    def __init__(self, topology, n_replicas: int, snapshot_resolution: int = 1, max_snapshots: Optional[int] = None,
                 device: int = 0, max_actions: int = 1, queue_capacity: int = 0):
        self.topology = topology
        self.n_replicas, self.max_actions = int(n_replicas), int(max_actions)
        self._struct, self._keep = self._topology_struct(topology)
        cfg = _abi.MaroCimConfig()
        cfg.n_replicas = self.n_replicas
        cfg.start_tick = int(topology.start_tick)
        cfg.snapshot_resolution = int(snapshot_resolution)
        cfg.max_snapshots = int(max_snapshots) if max_snapshots else 0
        self.snapshot_resolution, self._max_snapshots = cfg.snapshot_resolution, cfg.max_snapshots
        cfg.device = int(device)
        cfg.queue_capacity = int(queue_capacity)
        cfg.max_actions = self.max_actions
        h = C.c_void_p()
        _native.check(self._f("create")(C.byref(self._struct), C.byref(cfg), C.byref(h)))
        self._h = h
        self.frame_words = self._f("frame_words")(self._h)
        self.dec_words = self._f("decision_words")(self._h)
        self.decisions = np.zeros((self.n_replicas, self.dec_words), np.int32)
        self.metrics = np.zeros((self.n_replicas, self._MET_WORDS), np.int64)

    def close(self):
        if getattr(self, "_h", None):
            self._f("destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        if cuda_stream_ptr is None:
            _native.check(self._f("set_stream")(self._h, None, 0))
        else:
            _native.check(self._f("set_stream")(self._h, C.c_void_p(cuda_stream_ptr), 1))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        _native.check(self._f("reset")(self._h, None if m is None else m.ctypes.data))

    def step(self, actions=None, n_actions=None, active=None):
        a = n = m = None
        if actions is not None:
            a = np.ascontiguousarray(actions, np.int32)
            assert a.size == self.n_replicas * self.max_actions * 4, a.shape
        if n_actions is not None:
            n = np.ascontiguousarray(n_actions, np.int32)
        if active is not None:
            m = np.ascontiguousarray(active, np.uint8)
        _native.check(self._f("step")(
            self._h, None if m is None else m.ctypes.data, None if a is None else a.ctypes.data,
            None if n is None else n.ctypes.data, self.decisions.ctypes.data, self.metrics.ctypes.data))
        return self.decisions, self.metrics

    def pinned(self):
        if getattr(self, "_pinned", None) is None:
            self._pinned = _pinned_views(self._f("pinned_buffers"), self._h, self.n_replicas,
                                         self.max_actions, self.dec_words, self._MET_WORDS)
        return self._pinned

    def step_pinned(self, use_actions: bool = True, use_n_actions: bool = False, use_active: bool = False):
        _native.check(self._f("step_pinned")(self._h, int(use_actions), int(use_n_actions), int(use_active)))

    def step_device(self, d_decisions: int, d_metrics: int, d_actions: int = 0, d_n_actions: int = 0, d_active: int = 0):
        _native.check(self._f("step_device")(self._h, d_active or None, d_actions or None,
                                                          d_n_actions or None, d_decisions, d_metrics))

    def greedy_policy_device(self, d_decisions: int, d_actions: int):
        _native.check(self._f("greedy_policy_device")(self._h, d_decisions, d_actions))

    def rollout_device(self, d_decisions: int, d_metrics: int, n_steps: int):
        """``n_steps`` fused env-steps per replica in one launch, greedy top-1 agent on the device; ``d_decisions`` in/out"""
        _native.check(self._f("rollout_device")(self._h, n_steps, d_decisions, d_metrics))

    def attr_id(self, node: str, name: str) -> int:
        i = self._f("attr_id")(self._h, self._NODE[node], name.encode())
        if i < 0:
            raise KeyError(f"{node}.{name}")
        return i

    def attr_slots(self, node: str, attr_id: int) -> int:
        return self._f("attr_slots")(self._h, self._NODE[node], attr_id)

    def query(self, node: str, frame_indices, nodes, attrs, replicas=None) -> np.ndarray:
        reps = np.arange(self.n_replicas, dtype=np.int32) if replicas is None else np.ascontiguousarray(replicas, np.int32)
        fr = np.ascontiguousarray(frame_indices, np.int32)
        nd = np.ascontiguousarray(nodes, np.int32)
        at = np.ascontiguousarray([a if isinstance(a, (int, np.integer)) else self.attr_id(node, a) for a in attrs], np.int32)
        per = self._per_replica(node, at, len(fr), len(nd))
        out = np.zeros((len(reps), per), np.float64)
        pr = C.c_int64()
        _native.check(self._f("query")(self._h, reps.ctypes.data, len(reps), self._NODE[node], fr.ctypes.data,
                                                    len(fr), nd.ctypes.data, len(nd), at.ctypes.data, len(at),
                                                    out.ctypes.data, C.byref(pr)))
        return out

    def read_frame(self, replica: int = 0) -> np.ndarray:
        out = np.zeros(self.frame_words, np.int32)
        _native.check(self._f("read_frame")(self._h, replica, out.ctypes.data, out.size))
        return out

    def ticks(self) -> np.ndarray:
        out = np.zeros(self.n_replicas, np.int32)
        _native.check(self._f("ticks")(self._h, out.ctypes.data))
        return out

    def counters(self) -> np.ndarray:
        out = np.zeros((self.n_replicas, 4), np.int64)
        _native.check(self._f("counters")(self._h, out.ctypes.data))
        return out

    def snapshot_frames(self, replica: int = 0) -> np.ndarray:
        cap = 1 << 16
        out = np.zeros(cap, np.int32)
        n = C.c_int32()
        _native.check(self._f("snapshot_frames")(self._h, replica, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def snapshot_row(self, frame_index: int, replica: int = 0):
        if frame_index not in set(self.snapshot_frames(replica).tolist()):
            return None
        S = self.topology.n_stations
        lay, fw = _abi.bike_frame_layout(S)
        row = np.zeros(fw, np.int32)
        names = list(lay["stations"])
        vals = self.query("stations", [frame_index], np.arange(S), names, [replica])[0].reshape(S, len(names))
        for k, a in enumerate(names):
            off = lay["stations"][a][0]
            row[off:off + S] = vals[:, k].astype(np.int64)
        off, _, slots = lay["matrices"]["trips_adj"]
        row[off:off + slots] = self.query("matrices", [frame_index], [0], ["trips_adj"], [replica])[0].astype(np.int64)
        return row


class VmBatch(BikeBatch):
    """Columnar wrapper over the vm_scheduling entry points of the C ABI.  Decision rows: ``_abi.VM_DEC_*`` header +
    valid PM ids; metrics rows: 16 x int64 (``_abi.vm_metrics_dict`` decodes the float64 slots)."""

    _NODE = {"pms": 0, "racks": 1, "clusters": 2, "data_centers": 3, "zones": 4, "regions": 5}
    _PREFIX = "maro_vm"
    _MET_WORDS = 16

    def _topology_struct(self, topology):
        if getattr(topology, "error", None):
            raise Exception(topology.error)
        return _abi.vm_topology_struct(topology)

    def node_counts(self) -> dict:
        t = self.topology
        return {"pms": t.n_pm, "racks": t.n_rack, "clusters": t.n_cluster, "data_centers": t.n_dc, "zones": t.n_zone,
                "regions": t.n_region}

    # The reference's static backend keeps cpu_utilization / energy_consumption as float64, in the live frame and in the
    # snapshots alike (np_backend.pyx:143-148, 293: the dtype comes from the decoded type NAME "float").  The device ring
    # stores them as float32 words; both are exact functions of the integer k = 100 * cpu_utilization, so the query
    # result is lifted back to the float64 values the reference returns: k / 100 and the energy model evaluated at k.
    def _energy_f64(self, pm_type: int, k: int) -> float:
        cache = self.__dict__.setdefault("_energy_cache", {})
        key = (pm_type, min(k, 10000))
        if key not in cache:
            calibration, busy, idle = (float(x) for x in self.topology.pmtype_power[pm_type])
            u = min(1, (key[1] / 100.0) / 100)
            cache[key] = ((idle + (busy - idle) * (2 * u - pow(u, calibration))) / self.topology.ticks_per_hour) / 1000
        return cache[key]

    def query(self, node: str, frame_indices, nodes, attrs, replicas=None) -> np.ndarray:
        out = super().query(node, frame_indices, nodes, attrs, replicas)
        if node != "pms" or self.query_layout == "dynamic":  # (the RawBackend holds these attributes as float32 itself)
            return out
        names = [a if isinstance(a, str) else _abi.VM_NODE_ATTRS["pms"][int(a)] for a in attrs]
        if not any(n in _abi.VM_FLOAT_ATTRS for n in names):
            return out
        nodes = np.ascontiguousarray(nodes, np.int32)
        view = out.reshape(out.shape[0], len(frame_indices), len(nodes), len(names))
        if "cpu_utilization" in names:
            k = np.rint(view[..., names.index("cpu_utilization")] * 100.0)
        else:
            k = np.rint(super().query(node, frame_indices, nodes, ["cpu_utilization"], replicas).reshape(view.shape[:3]) * 100.0)
        if "cpu_utilization" in names:
            view[..., names.index("cpu_utilization")] = k / 100.0
        if "energy_consumption" in names:
            e = view[..., names.index("energy_consumption")]
            types = self.topology.pm_attr[nodes, 2]
            held = e != 0.0  # frames the ring no longer holds read as zeros and stay zeros
            for idx in np.argwhere(held):
                e[tuple(idx)] = self._energy_f64(int(types[idx[2]]), int(k[tuple(idx)]))
        return out

    def greedy_policy_device(self, d_decisions: int, d_actions: int):
        raise AttributeError("vm_scheduling has best_fit_policy_device")

    def best_fit_policy_device(self, d_decisions: int, d_actions: int):
        _native.check(self._f("best_fit_policy_device")(self._h, d_decisions, d_actions))

    def rollout_device(self, d_decisions: int, d_metrics: int, n_steps: int):
        """``n_steps`` env-steps fused into one launch with the best-fit agent as a device callback (maro_vm_rollout_device);
        d_decisions [B][dec_words] int32 / d_metrics [B][16] int64 carry the last row across launches."""
        _native.check(_native.lib().maro_vm_rollout_device(self._h, int(n_steps), d_decisions, d_metrics))

    def snapshot_row(self, frame_index: int, replica: int = 0):
        if frame_index not in set(self.snapshot_frames(replica).tolist()):
            return None
        lay, fw = _abi.vm_frame_layout(self.topology)
        row = np.zeros(fw, np.int32)
        for node, attrs in lay.items():
            names = list(attrs)
            n = attrs[names[0]][1]
            vals = self.query(node, [frame_index], np.arange(n), names, [replica])[0].reshape(n, len(names))
            for k, a in enumerate(names):
                off = attrs[a][0]
                if a in _abi.VM_FLOAT_ATTRS:
                    row[off:off + n] = vals[:, k].astype(np.float32).view(np.int32)
                else:
                    row[off:off + n] = vals[:, k].astype(np.int64)
        return row
