"""TEST INFRASTRUCTURE — a drop-in for ``maro_b200.batch.CimBatch`` backed by the host emulation of the device code
(tests/emul.py), one single-replica emulator per replica.  Lets the CPU suite drive the Python façades
(maro_b200.simulator.Env, maro_b200.vector_env.VectorEnv, SnapshotList queries) exactly like the GPU suite does, with the
kernel logic instead of the kernel.  Never imported by the package."""
import numpy as np

from emul import EmulEnv
from maro_b200 import _abi
from maro_b200.scenarios.cim.topology import CimTopology

_NODES = ("ports", "vessels", "matrices")


class EmulCimBatch:
    def __init__(self, topologies, n_replicas, start_tick=0, snapshot_resolution=1, max_snapshots=None, device=0, max_actions=1,
                 replica_topology=None, queue_capacity=0, decision_mode=0):
        if isinstance(topologies, CimTopology):
            topologies = [topologies]
        self.topologies = list(topologies)
        self.decision_mode = int(decision_mode)
        V = self.topologies[0].n_vessels
        self.n_replicas = int(n_replicas)
        self.max_actions = max(int(max_actions), V) if self.decision_mode == 1 else int(max_actions)
        self.dec_words = 8 * (V if self.decision_mode == 1 else 1)
        self.start_tick, self.snapshot_resolution = int(start_tick), int(snapshot_resolution)
        self._cfg = (start_tick, snapshot_resolution, max_snapshots)
        self._rt = [0] * self.n_replicas if replica_topology is None else [int(x) for x in replica_topology]
        self._envs = [self._make(i) for i in range(self.n_replicas)]
        t = self.topologies[0]
        self._lay, self.frame_words = _abi.frame_layout(t.n_ports, t.n_vessels, t.past_stop_number, t.future_stop_number)
        self._attrs = {n: list(self._lay[n]) for n in _NODES}
        self.decisions = np.zeros((self.n_replicas, self.dec_words), np.int32)
        self.metrics = np.zeros((self.n_replicas, 3), np.int64)

    def _make(self, i):
        st, res, ms = self._cfg
        return EmulEnv(self.topologies[self._rt[i]], 1, st, res, ms, max_actions=self.max_actions, decision_mode=self.decision_mode)

    def node_counts(self):
        t = self.topologies[0]
        return {"ports": t.n_ports, "vessels": t.n_vessels, "matrices": 1}

    def ring_rows(self):
        st, res, ms = self._cfg
        total = -(-(self.topologies[0].max_tick - st) // res)
        return min(int(ms), total) if ms else total

    def snapshot_row(self, frame_index, replica=0):
        return self._envs[replica].snapshot(int(frame_index))

    def close(self):
        self._envs = []

    def set_stream(self, ptr):
        pass

    def reset(self, mask=None):
        for i in range(self.n_replicas):
            if mask is None or mask[i]:
                self._envs[i] = self._make(i)  # Env.reset: fresh replica (the step / tick counters are not inspected here)

    def set_topology(self, index, topo):
        self.topologies[index] = topo

    def step(self, actions=None, n_actions=None, active=None):
        for i, e in enumerate(self._envs):
            if active is not None and not active[i]:
                self.decisions[i] = 0
                self.decisions[i, 6] = _abi.STATUS_INACTIVE
                continue
            if actions is None:
                d, m = e.step(None)
            else:
                a = np.asarray(actions, np.int32).reshape(self.n_replicas, self.max_actions, 4)[i:i + 1]
                n = np.asarray([1 if n_actions is None else int(np.asarray(n_actions).reshape(-1)[i])], np.int32)
                d, m = e.step(a, n)
            self.decisions[i], self.metrics[i] = d[0], m[0]
        return self.decisions, self.metrics

    # -- inspection (static-backend query semantics, np_backend.pyx:520-549) -------------------------------------
    def attr_id(self, node, name):
        if name not in self._lay[node]:
            raise KeyError(f"{node}.{name}")
        return self._attrs[node].index(name)

    def attr_slots(self, node, attr_id):
        return self._lay[node][self._attrs[node][attr_id]][2]

    def snapshot_frames(self, replica=0):
        e = self._envs[replica]
        st, res, ms = self._cfg
        total = -(-(self.topologies[self._rt[replica]].max_tick - st) // res)
        return np.asarray([f for f in range(total) if e.snapshot(f) is not None], np.int32)

    query_layout = "static"

    def set_query_layout(self, layout):
        self.query_layout = layout

    def query_shape(self, node, attrs, n_frames, n_nodes):
        slots = [self._lay[node][a if isinstance(a, str) else self._attrs[node][int(a)]][2] for a in attrs]
        return (n_frames, n_nodes, len(slots), max(slots))

    def query(self, node, frame_indices, nodes, attrs, replicas=None):
        """static layout: zero-padded flat float64 (np_backend.pyx:520-549); dynamic: every attribute padded to the widest
        one's slots with NaN, unknown frames NaN, values through float32 (raw/snapshotlist.cpp:244-318) — an independent
        Python statement of what cim_query_kernel does on the device"""
        reps = range(self.n_replicas) if replicas is None else replicas
        dyn = self.query_layout == "dynamic"
        names = [a if isinstance(a, str) else self._attrs[node][int(a)] for a in attrs]
        width = max(self._lay[node][n][2] for n in names) if names else 0
        out = []
        for r in reps:
            vals = []
            for f in frame_indices:
                row = self._envs[r].snapshot(int(f)) if f >= 0 else None
                for nd in nodes:
                    for name in names:
                        off, _, slots = self._lay[node][name]
                        if row is None:
                            v = np.full(slots, np.nan) if dyn else np.zeros(slots)
                        else:
                            w = row[off + nd * slots: off + (nd + 1) * slots]
                            v = (w.view(np.float32) if name == "transfer_cost" else w).astype(np.float64)
                            if dyn:
                                v = v.astype(np.float32).astype(np.float64)
                        if dyn:
                            v = np.concatenate([v, np.full(width - slots, np.nan)])
                        vals.append(v)
            out.append(np.concatenate(vals) if vals else np.zeros(0))
        return np.asarray(out, np.float64)

    def read_frame(self, replica=0):
        return self._envs[replica].frame()

    def ticks(self):
        return np.asarray([e.tick() for e in self._envs], np.int32)


class _EmulScenarioBatch:
    """shared parts of the citi_bike / vm_scheduling adapters: one single-replica emulator per replica, queries answered
    from its snapshots with the static backend's conventions"""

    _FLOAT = ()

    def _make(self):
        raise NotImplementedError

    def _setup(self, topology, n_replicas, max_actions, layout, frame_words, met_words):
        self.topology = topology
        self.n_replicas, self.max_actions = int(n_replicas), int(max_actions)
        self._lay, self.frame_words = layout, frame_words
        self._attrs = {n: list(a) for n, a in layout.items()}
        self._envs = [self._make() for _ in range(self.n_replicas)]
        self.dec_words = self._envs[0].dec_words
        self.decisions = np.zeros((self.n_replicas, self.dec_words), np.int32)
        self.metrics = np.zeros((self.n_replicas, met_words), np.int64)

    def node_counts(self):
        return {n: next(iter(a.values()))[1] for n, a in self._lay.items()}

    def close(self):
        self._envs = []

    def set_stream(self, ptr):
        pass

    def reset(self, mask=None):
        for i in range(self.n_replicas):
            if mask is None or mask[i]:
                self._envs[i] = self._make()

    def _step_one(self, e, actions):
        raise NotImplementedError

    def step(self, actions=None, n_actions=None, active=None):
        for i, e in enumerate(self._envs):
            if active is not None and not active[i]:
                self.decisions[i] = 0
                self.decisions[i, 6] = _abi.STATUS_INACTIVE
                continue
            a = None
            if actions is not None:
                n = 1 if n_actions is None else int(np.asarray(n_actions).reshape(-1)[i])
                a = np.asarray(actions, np.int32).reshape(self.n_replicas, self.max_actions, 4)[i, :n]
            st, d, m = self._step_one(e, a)
            self.decisions[i], self.metrics[i] = d, m
        return self.decisions, self.metrics

    def attr_id(self, node, name):
        if name not in self._lay[node]:
            raise KeyError(f"{node}.{name}")
        return self._attrs[node].index(name)

    def attr_slots(self, node, attr_id):
        return self._lay[node][self._attrs[node][attr_id]][2]

    def snapshot_frames(self, replica=0):
        e = self._envs[replica]
        return np.asarray([f for f in range(self._total_frames) if e.snapshot(f) is not None], np.int32)

    def query(self, node, frame_indices, nodes, attrs, replicas=None):
        reps = range(self.n_replicas) if replicas is None else replicas
        out = []
        for r in reps:
            vals = []
            for f in frame_indices:
                row = self._envs[r].snapshot(int(f)) if f >= 0 else None
                for nd in nodes:
                    for a in attrs:
                        name = a if isinstance(a, str) else self._attrs[node][int(a)]
                        off, _, slots = self._lay[node][name]
                        if row is None:
                            vals.append(np.zeros(slots))
                        else:
                            w = row[off + nd * slots: off + (nd + 1) * slots]
                            vals.append((w.view(np.float32) if name in self._FLOAT else w).astype(np.float64))
            out.append(np.concatenate(vals) if vals else np.zeros(0))
        return np.asarray(out, np.float64)

    def read_frame(self, replica=0):
        return self._envs[replica].frame()

    def ticks(self):
        return np.asarray([e.tick() for e in self._envs], np.int32)


class EmulBikeBatch(_EmulScenarioBatch):
    def __init__(self, topology, n_replicas, snapshot_resolution=1, max_snapshots=None, device=0, max_actions=1, queue_capacity=0):
        from emul import BikeEmulEnv

        self._mk = lambda: BikeEmulEnv(topology, snapshot_resolution, max_snapshots, max_actions=max_actions)
        lay, fw = _abi.bike_frame_layout(topology.n_stations)
        self._total_frames = -(-(topology.max_tick - topology.start_tick) // snapshot_resolution)
        self._setup(topology, n_replicas, max_actions, lay, fw, 3)

    def _make(self):
        return self._mk()

    def _step_one(self, e, a):
        return e.step1(a)


class EmulVmBatch(_EmulScenarioBatch):
    _FLOAT = _abi.VM_FLOAT_ATTRS

    # the float64 lift of the product wrapper (maro_b200.batch.VmBatch.query) on top of the emulated raw query
    def query(self, node, frame_indices, nodes, attrs, replicas=None):
        from maro_b200.batch import BikeBatch, VmBatch

        class _Raw(BikeBatch):  # stand-in for `super().query` inside VmBatch.query
            pass

        raw = _EmulScenarioBatch.query
        outer = self

        class _Shim(VmBatch):
            def __init__(self):
                self.topology = outer.topology

        shim = _Shim()
        # VmBatch.query calls super().query (BikeBatch.query): route that to the emulated raw query
        orig = BikeBatch.query
        BikeBatch.query = lambda _self, n, f, nd, a, r=None: raw(outer, n, f, nd, a, r)
        try:
            return VmBatch.query(shim, node, frame_indices, nodes, attrs, replicas)
        finally:
            BikeBatch.query = orig

    def __init__(self, topology, n_replicas, snapshot_resolution=1, max_snapshots=None, device=0, max_actions=1, queue_capacity=0):
        from emul import VmEmulEnv

        if getattr(topology, "error", None):
            raise Exception(topology.error)
        self._mk = lambda: VmEmulEnv(topology, snapshot_resolution, max_snapshots, max_actions=max_actions)
        lay, fw = _abi.vm_frame_layout(topology)
        self._total_frames = -(-(topology.max_tick - topology.start_tick) // snapshot_resolution)
        self._setup(topology, n_replicas, max_actions, lay, fw, 16)

    def _make(self):
        return self._mk()

    def _step_one(self, e, a):
        return e.step(None if a is None else a)
