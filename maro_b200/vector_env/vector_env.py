"""``VectorEnv`` — drop-in for ``maro.vector_env.VectorEnv`` (maro/vector_env/vector_env.py:20-232).

The reference forks one OS process per environment and caps ``batch_num`` at ``os.cpu_count()``; here the batch is
``batch_num`` replicas of one CUDA handle (no cap), and one ``step`` is one kernel launch.  ``step`` keeps the
reference's list / dict / broadcast conventions and return shapes; ``step_columnar`` is the allocation-free variant
that returns the raw decision / metric arrays."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .. import _abi
from ..batch import BikeBatch, CimBatch, VmBatch
from ..scenarios.cim.common import ActionScope, DecisionEvent, action_row, encode_action
from ..scenarios.cim.topology import build_topology, load_config, next_topology_seed
from ..simulator.env import DecisionMode, SnapshotList, make_metrics, parse_query_key


class VectorEnv:
    class SnapshotListNodeWrapper:
        def __init__(self, env, node_name: str):
            self.node_name, self._env = node_name, env

        def __getitem__(self, args) -> List[np.ndarray]:
            return self._env._query(self.node_name, args)

    class SnapshotListWrapper:
        def __init__(self, env):
            self._env = env

        def __getitem__(self, node_name: str):
            return VectorEnv.SnapshotListNodeWrapper(self._env, node_name)

    def __init__(self, batch_num: int, scenario: str = None, topology: str = None, start_tick: int = 0,
                 durations: int = 100, snapshot_resolution: int = 1, max_snapshots: int = None,
                 decision_mode=DecisionMode.Sequential, business_engine_cls: type = None,
                 disable_finished_events: bool = False, options: dict = {}, device: int = 0, seeds=None):
        assert batch_num > 0
        self._joint = int(decision_mode) == int(DecisionMode.Joint)
        if scenario not in ("cim", "citi_bike", "vm_scheduling") or business_engine_cls is not None or (self._joint and scenario != "cim"):
            raise NotImplementedError("the CUDA core implements scenario='cim' / 'citi_bike' / 'vm_scheduling' (Joint mode: cim)")
        self._batch_num = batch_num
        self._scenario = scenario
        self._start_tick, self._resolution = start_tick, snapshot_resolution
        if scenario == "citi_bike":
            from ..scenarios.citi_bike.data import build_bike_topology, load_bike_config

            topo = build_bike_topology(load_bike_config(topology), start_tick, start_tick + durations,
                                       transfer_seed=int((options or {}).get("transfer_seed", 0)))
            self._batch = BikeBatch(topo, batch_num, snapshot_resolution, max_snapshots, device=device, max_actions=4)
            self._finish_init(batch_num, start_tick)
            return
        if scenario == "vm_scheduling":
            from ..scenarios.vm_scheduling.data import build_vm_topology, load_vm_config

            topo = build_vm_topology(load_vm_config(topology), start_tick, start_tick + durations)
            self._batch = VmBatch(topo, batch_num, snapshot_resolution, max_snapshots, device=device, max_actions=4)
            self._finish_init(batch_num, start_tick)
            return
        conf = load_config(topology)
        if seeds is None:
            topos, rt = [build_topology(conf, start_tick + durations)], None
        else:  # one topology instance per distinct seed (per-process seeds of the reference's VectorEnv)
            seeds = [int(s) for s in seeds]
            assert len(seeds) == batch_num
            uniq = sorted(set(seeds))
            topos = [build_topology(conf, start_tick + durations, seed=s) for s in uniq]
            rt = np.asarray([uniq.index(s) for s in seeds], np.int32)
        self._conf, self._max_tick, self._rt = conf, start_tick + durations, rt
        self._ctor = (batch_num, start_tick, snapshot_resolution, max_snapshots, device)
        self._batch = CimBatch(topos, batch_num, start_tick, snapshot_resolution, max_snapshots, device=device,
                               max_actions=4, replica_topology=rt, decision_mode=int(self._joint))
        self._finish_init(batch_num, start_tick)

    def _finish_init(self, batch_num, start_tick):
        import os

        self._backend_name = os.environ.get("DEFAULT_BACKEND_NAME", "static")  # maro/backends/frame.pyx:496-504
        if self._backend_name == "dynamic":
            self._batch.set_query_layout("dynamic")
        self._snapshot_wrapper = VectorEnv.SnapshotListWrapper(self)
        self._snapshot_lists = [SnapshotList(self._batch, i) for i in range(batch_num)]
        self._done = np.zeros(batch_num, bool)
        self._ticks = np.full(batch_num, start_tick, np.int64)
        self._act = np.zeros((batch_num, self._batch.max_actions, 4), np.int32)
        self._nact = np.zeros(batch_num, np.int32)
        self._active = np.ones(batch_num, np.uint8)

    # ---- reference surface
    @property
    def batch_number(self) -> int:
        return self._batch_num

    @property
    def snapshot_list(self):
        return self._snapshot_wrapper

    @property
    def tick(self) -> List[int]:
        return [int(t) for t in self._ticks]

    @property
    def frame_index(self) -> List[int]:
        return [int((t - self._start_tick) // self._resolution) for t in self._ticks]

    def _encode(self, i, action):
        if action is None:
            self._nact[i] = 0
            return
        acts = action if isinstance(action, list) else [action]
        for k, a in enumerate(acts):
            if self._scenario == "citi_bike":
                from ..scenarios.citi_bike.common import encode_bike_action

                encode_bike_action(a, self._act[i, k])
            elif self._scenario == "vm_scheduling":
                from ..scenarios.vm_scheduling.common import encode_vm_action

                encode_vm_action(a, self._act[i, k])
            else:
                encode_action(a, self._act[i, k])
        self._nact[i] = len(acts)

    def _encode_joint(self, i, answers):
        """Joint mode: env i's answers = None or a list with one entry (Action or None) per decision event of its last step"""
        answers = [] if answers is None else (answers if isinstance(answers, list) else [answers])
        if len(answers) > self._act.shape[1]:
            raise ValueError("more answers than decision events")
        for k, a in enumerate(answers):
            self._act[i, k] = (0, 0, 0, 2) if a is None else action_row(a)
        self._nact[i] = len(answers)

    def _encode_all_cim(self, actions):
        """one action (or None / list) per env -> the int32 action rows, filled with ONE array assignment per column block
        instead of four element writes per env (the per-env Python objects are the caller's, the loop stays)"""
        A = self._act.shape[1]
        rows = [[(0, 0, 0, 0)] * A for _ in range(self._batch_num)]
        counts = [0] * self._batch_num
        for i, a in enumerate(actions):
            if a is None:
                continue
            if isinstance(a, list):
                if len(a) > A:
                    raise ValueError("too many actions for one decision event")
                for k, x in enumerate(a):
                    rows[i][k] = action_row(x)
                counts[i] = len(a)
            else:
                rows[i][0] = action_row(a)
                counts[i] = 1
        self._act[:] = rows
        self._nact[:] = counts

    def step(self, action):
        """list -> one action per env; dict -> only those envs advance; anything else -> broadcast."""
        B = self._batch_num
        cim = self._scenario == "cim" and not self._joint
        if self._joint:  # list: one answer list per env; dict: only those envs; None: broadcast "no answers"
            if type(action) is dict:
                self._active[:] = 0
                for i, a in action.items():
                    self._encode_joint(i, a)
                    self._active[i] = 1
            else:
                per_env = action if (type(action) is list and len(action) == B and all(a is None or isinstance(a, list) for a in action)) else [action] * B
                for i in range(B):
                    self._encode_joint(i, per_env[i])
                self._active[:] = 1
            dec, met = self._batch.step(self._act, self._nact, self._active)
            return self._decode_cim(dec, met)
        if type(action) is list:
            assert len(action) == B
            if cim:
                self._encode_all_cim(action)
            else:
                for i in range(B):
                    self._encode(i, action[i])
            self._active[:] = 1
        elif type(action) is dict:
            self._active[:] = 0
            for i, a in action.items():
                self._encode(i, a)
                self._active[i] = 1
        else:
            if cim:
                self._encode_all_cim([action] * B)
            else:
                for i in range(B):
                    self._encode(i, action)
            self._active[:] = 1
        dec, met = self._batch.step(self._act, self._nact, self._active)
        if cim:
            return self._decode_cim(dec, met)
        # citi_bike / vm_scheduling: the decision rows become Python ints in one ``tolist()`` (per-element numpy indexing costs more
        # than building the objects); the metrics rows stay arrays (vm_scheduling keeps float64 bit patterns in them)
        metrics, events = [], []
        bike = self._scenario == "citi_bike"
        if bike:
            from ..scenarios.citi_bike.common import decode_bike_decision
        else:
            from ..scenarios.vm_scheduling.common import decode_vm_decision, decode_vm_metrics
        rows, mets = dec.tolist(), (met.tolist() if bike else None)
        for i in (range(B) if self._active.all() else np.flatnonzero(self._active).tolist()):
            d = rows[i]
            st = d[_abi.DEC_STATUS]
            if st == _abi.STATUS_BAD_ACTION:
                raise AssertionError(f"env {i}: invalid action (outside the action scope / unknown VM or PM id)")
            if st == _abi.STATUS_QUEUE_OVERFLOW:
                raise RuntimeError(f"env {i}: event queue overflow")
            if st == _abi.STATUS_FINISHED:  # env_process.py:37-40
                metrics.append(None)
                events.append(None)
                continue
            self._ticks[i] = d[0]
            if bike:
                m = mets[i]
                metrics.append({"trip_requirements": m[0], "bike_shortage": m[1], "operation_number": m[2]})
                if st == _abi.STATUS_DONE:
                    self._done[i] = True
                    events.append(None)
                else:
                    events.append(decode_bike_decision(d, self._snapshot_lists[i]))
                continue
            metrics.append(decode_vm_metrics(met[i]))
            if st == _abi.STATUS_DONE:
                self._done[i] = True
                events.append(None)
            else:
                events.append(decode_vm_decision(d))
        return metrics, events, bool(self._done.all())

    def _decode_cim(self, dec, met):
        """decision / metrics rows -> the reference's per-env Python objects; the arrays are converted to Python ints in
        one ``tolist()`` each (per-element numpy indexing costs more than building the objects)"""
        st_col = dec[:, _abi.DEC_STATUS]
        active = self._active.astype(bool)
        bad = active & (st_col < 0)
        if bad.any():
            i = int(np.argmax(bad))
            if st_col[i] == _abi.STATUS_BAD_ACTION:
                raise AssertionError(f"env {i}: invalid action (outside the action scope / unknown VM or PM id)")
            raise RuntimeError(f"env {i}: event queue overflow")
        live = active & ((st_col == _abi.STATUS_DECISION) | (st_col == _abi.STATUS_DONE))
        self._ticks[live] = dec[live, 0]
        self._done |= active & (st_col == _abi.STATUS_DONE)
        rows, mets, snaps = dec.tolist(), met.tolist(), self._snapshot_lists
        metrics, events = [], []
        ST_DEC, ST_DONE = _abi.STATUS_DECISION, _abi.STATUS_DONE
        for i in (range(self._batch_num) if active.all() else np.flatnonzero(active).tolist()):
            d = rows[i]
            st = d[6]
            if st == ST_DEC and self._joint:  # every decision event of the tick: rows of 8 words up to the first non-decision row
                metrics.append(make_metrics(mets[i]))
                evs = []
                for k in range(0, len(d), 8):
                    if d[k + 6] != ST_DEC:
                        break
                    evs.append(DecisionEvent(d[k], d[k + 1], d[k + 2], snaps[i], ActionScope(d[k + 3], d[k + 4]), d[k + 5]))
                events.append(evs)
            elif st == ST_DEC:
                metrics.append(make_metrics(mets[i]))
                events.append(DecisionEvent(d[0], d[1], d[2], snaps[i], ActionScope(d[3], d[4]), d[5]))
            elif st == ST_DONE:
                metrics.append(make_metrics(mets[i]))
                events.append(None)
            else:  # STATUS_FINISHED: (None, None, True) from that env (env_process.py:37-40)
                metrics.append(None)
                events.append(None)
        return metrics, events, bool(self._done.all())

    def step_columnar(self, actions: Optional[np.ndarray] = None, n_actions=None, active=None):
        """Array-in / array-out step: actions int32 [B][4][4] (or None), returns (decisions [B][8], metrics [B][3], done)."""
        dec, met = self._batch.step(actions, n_actions, active)
        st = dec[:, _abi.DEC_STATUS]
        live = (st == _abi.STATUS_DECISION) | (st == _abi.STATUS_DONE)
        self._ticks[live] = dec[live, 0]
        self._done |= (st == _abi.STATUS_DONE) | (st == _abi.STATUS_FINISHED)
        return dec, met, bool(self._done.all())

    def reset(self, keep_seed: bool = False):
        """Every env process of the reference calls ``env.reset()`` (env_process.py:55-57), i.e. ``keep_seed=False``: each
        environment draws a new topology seed from its own route_init stream (cim_data_container_helpers.py:56-66).  Here:
        one re-seeded instance per distinct topology of the batch (envs that shared a seed keep sharing the new one, exactly
        like identical processes do).  ``keep_seed=True`` replays the same instances."""
        if self._scenario == "cim" and not keep_seed:
            new = [build_topology(self._conf, self._max_tick, seed=next_topology_seed(t)) for t in self._batch.topologies]
            try:
                for k, t in enumerate(new):
                    self._batch.set_topology(k, t)
            except RuntimeError:  # a new instance does not fit the handle's event horizon / pool: fresh handle
                B, start_tick, res, max_snaps, device = self._ctor
                self._batch.close()
                self._batch = CimBatch(new, B, start_tick, res, max_snaps, device=device, max_actions=4, replica_topology=self._rt,
                                       decision_mode=int(self._joint))
                if self._backend_name == "dynamic":
                    self._batch.set_query_layout("dynamic")
                self._snapshot_lists = [SnapshotList(self._batch, i) for i in range(B)]
        self._batch.reset()
        self._done[:] = False
        self._ticks[:] = self._start_tick

    def stop(self):
        self._batch.close()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.stop()

    def __del__(self):
        try:
            self.stop()
        except Exception:
            pass

    def _query(self, node_name: str, args: slice):
        """``env.snapshot_list[node][ticks:nodes:attrs]`` -> one array per env (vector_env.py:177-217 asks every process);
        here ONE batched device gather + one copy for all envs when the ticks are given."""
        ticks, nodes, attrs = parse_query_key(args)
        if attrs is None:
            return [None] * self._batch_num
        if len(ticks) == 0:  # "all frames" differs per env (each has its own frame list)
            return [sl[node_name][args] for sl in self._snapshot_lists]
        node = self._snapshot_lists[0][node_name]
        if node is None:
            return [None] * self._batch_num
        if len(nodes) == 0:
            nodes = list(range(len(node)))
        try:
            ids = [self._batch.attr_id(node_name, a) for a in attrs]
        except KeyError:
            raise KeyError(f"invalid attribute for {node_name}: {attrs}")
        out = self._batch.query(node_name, ticks, nodes, ids)
        if self._batch.query_layout == "dynamic":
            shape = self._batch.query_shape(node_name, ids, len(ticks), len(nodes))
            return [out[i].reshape(shape) for i in range(self._batch_num)]
        return [out[i] for i in range(self._batch_num)]

    @property
    def batch(self) -> CimBatch:
        return self._batch
