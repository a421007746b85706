"""TEST INFRASTRUCTURE — host emulation of the kernel's per-replica logic (tests/_emul_src/emul.cpp).

Same device source (maro_b200/csrc/cim_core.cuh) compiled for the CPU with one lane.  Used only by the CPU test
suite to check the kernel logic against the golden traces without a GPU; the package never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np

from maro_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_emul", "libmaro_emul.so")
SRC = os.path.join(HERE, "_emul_src", "emul.cpp")
CORE = os.path.join(HERE, "..", "maro_b200", "csrc")
_lib = None


def lib():
    global _lib
    if _lib is None:
        deps = [SRC, os.path.join(os.path.dirname(SRC), "warp_emul.hpp"), os.path.join(CORE, "cim_core.cuh"),
                os.path.join(CORE, "cim_host.hpp"), os.path.join(CORE, "bike_core.cuh"), os.path.join(CORE, "bike_host.hpp"),
                os.path.join(CORE, "vm_core.cuh"), os.path.join(CORE, "vm_host.hpp")]
        if not os.path.isfile(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
            os.makedirs(os.path.dirname(LIB), exist_ok=True)
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-ffp-contract=off",
                                   "-DMARO_HOST_EMULATION", "-I", os.path.dirname(SRC), "-I", os.path.join(HERE, "..", "include"),
                                   "-shared", "-fPIC", SRC,
                                   "-o", LIB])
        _lib = C.CDLL(LIB)
        _lib.emul_create.restype = C.c_void_p
        _lib.emul_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        for name, args in [("emul_destroy", 1), ("emul_reset", 1)]:
            getattr(_lib, name).argtypes = [C.c_void_p]
        _lib.emul_step.argtypes = [C.c_void_p] * 5
        _lib.emul_frame_words.argtypes = [C.c_void_p]
        _lib.emul_read_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.emul_read_snapshot.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.emul_tick.argtypes = [C.c_void_p, C.c_int]
        _lib.emul_counters.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.bike_emul_create.restype = C.c_void_p
        _lib.bike_emul_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.bike_emul_destroy.argtypes = [C.c_void_p]
        _lib.bike_emul_reseed.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        _lib.bike_emul_dec_words.argtypes = [C.c_void_p]
        _lib.bike_emul_frame_words.argtypes = [C.c_void_p]
        _lib.bike_emul_step.argtypes = [C.c_void_p] * 5
        _lib.bike_emul_read_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.bike_emul_read_snapshot.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.bike_emul_tick.argtypes = [C.c_void_p, C.c_int]
        _lib.bike_emul_counters.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.vm_emul_create.restype = C.c_void_p
        _lib.vm_emul_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.vm_emul_destroy.argtypes = [C.c_void_p]
        _lib.vm_emul_reset.argtypes = [C.c_void_p]
        _lib.vm_emul_dec_words.argtypes = [C.c_void_p]
        _lib.vm_emul_frame_words.argtypes = [C.c_void_p]
        _lib.vm_emul_step.argtypes = [C.c_void_p] * 5
        _lib.vm_emul_read_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.vm_emul_read_snapshot.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.vm_emul_tick.argtypes = [C.c_void_p, C.c_int]
        _lib.vm_emul_counters.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return _lib


class EmulEnv:
    def __init__(self, topos, n_replicas=1, start_tick=0, snapshot_resolution=1, max_snapshots=None, max_actions=2,
                 replica_topology=None, lanes=0, decision_mode=0):
        if not isinstance(topos, (list, tuple)):
            topos = [topos]
        self._keep = []
        arr = (_abi.MaroCimTopology * len(topos))()
        for i, t in enumerate(topos):
            s, keep = _abi.topology_struct(t)
            arr[i] = s
            self._keep.append(keep)
        cfg = _abi.MaroCimConfig()
        cfg.n_replicas = n_replicas
        cfg.start_tick = start_tick
        cfg.snapshot_resolution = snapshot_resolution
        cfg.max_snapshots = int(max_snapshots) if max_snapshots else 0
        cfg.max_actions = max_actions
        cfg.decision_mode = int(decision_mode)
        if replica_topology is not None:
            rt = np.ascontiguousarray(replica_topology, np.int32)
            self._keep.append(rt)
            cfg.replica_topology = rt.ctypes.data_as(C.POINTER(C.c_int32))
        self.B = n_replicas
        self._h = lib().emul_create(arr, len(topos), C.byref(cfg), lanes)
        assert self._h
        self.frame_words = lib().emul_frame_words(self._h)
        self.A = lib().emul_max_actions(self._h)
        self.DW = lib().emul_dec_words(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emul_destroy(self._h)
            self._h = None

    def step_joint(self, answers=None):
        """DecisionMode.Joint, one replica: answers = list of (vessel, port, qty, type) or None -> (status, rows [n][8], metrics)"""
        if answers is None:
            dec, met = self.step(None)
        else:
            rows = np.asarray([(0, 0, 0, 2) if a is None else a for a in answers], np.int32).reshape(1, -1, 4)
            dec, met = self.step(rows if rows.shape[1] else np.zeros((1, 1, 4), np.int32), np.asarray([rows.shape[1]], np.int32))
        d = dec[0].reshape(-1, 8)
        st = int(d[0, 6])
        n = 0
        while st == 0 and n < len(d) and d[n, 6] == 0:
            n += 1
        return st, d[:n] if st == 0 else d[:1], met[0]

    def step(self, actions=None, n_actions=None):
        dec = np.zeros((self.B, self.DW), np.int32)
        met = np.zeros((self.B, 3), np.int64)
        if actions is None:
            lib().emul_step(self._h, None, None, dec.ctypes.data, met.ctypes.data)
        else:
            a = np.zeros((self.B, self.A, 4), np.int32)
            src = np.asarray(actions, np.int32).reshape(self.B, -1, 4)
            a[:, :src.shape[1]] = src
            n = np.full(self.B, src.shape[1], np.int32) if n_actions is None else np.ascontiguousarray(n_actions, np.int32)
            lib().emul_step(self._h, a.ctypes.data, n.ctypes.data, dec.ctypes.data, met.ctypes.data)
        return dec, met

    def step1(self, actions=None):
        """single-replica convenience: (status, dec[8], met[3])"""
        dec, met = self.step(None if actions is None else np.asarray(actions, np.int32).reshape(1, -1, 4))
        return int(dec[0, 6]), dec[0], met[0]

    def frame(self, rep=0):
        out = np.zeros(self.frame_words, np.int32)
        lib().emul_read_frame(self._h, rep, out.ctypes.data)
        return out

    def snapshot(self, frame_index, rep=0):
        out = np.zeros(self.frame_words, np.int32)
        return out if lib().emul_read_snapshot(self._h, rep, frame_index, out.ctypes.data) else None

    def tick(self, rep=0):
        return lib().emul_tick(self._h, rep)

    def counters(self, rep=0):
        out = np.zeros(4, np.int64)
        lib().emul_counters(self._h, rep, out.ctypes.data)
        return out


class BikeEmulEnv:
    """citi_bike device logic under the thread-per-lane emulator (one replica)."""

    def __init__(self, topo, snapshot_resolution=1, max_snapshots=None, max_actions=2, lanes=0):
        self._struct, self._keep = _abi.bike_topology_struct(topo)
        cfg = _abi.MaroCimConfig()
        cfg.n_replicas = 1
        cfg.start_tick = topo.start_tick
        cfg.snapshot_resolution = snapshot_resolution
        cfg.max_snapshots = int(max_snapshots) if max_snapshots else 0
        cfg.max_actions = max_actions
        self.A = max_actions
        self._h = lib().bike_emul_create(C.byref(self._struct), C.byref(cfg), lanes)
        assert self._h
        self.dec_words = lib().bike_emul_dec_words(self._h)
        self.frame_words = lib().bike_emul_frame_words(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bike_emul_destroy(self._h)
            self._h = None

    def step1(self, actions=None):
        dec = np.zeros((1, self.dec_words), np.int32)
        met = np.zeros((1, 3), np.int64)
        if actions is None:
            lib().bike_emul_step(self._h, None, None, dec.ctypes.data, met.ctypes.data)
        else:
            a = np.zeros((1, self.A, 4), np.int32)
            src = np.asarray(actions, np.int32).reshape(1, -1, 4)
            a[:, :src.shape[1]] = src
            n = np.full(1, src.shape[1], np.int32)
            lib().bike_emul_step(self._h, a.ctypes.data, n.ctypes.data, dec.ctypes.data, met.ctypes.data)
        return int(dec[0, 6]), dec[0], met[0]

    def frame(self):
        out = np.zeros(self.frame_words, np.int32)
        lib().bike_emul_read_frame(self._h, 0, out.ctypes.data)
        return out

    def snapshot(self, frame_index):
        out = np.zeros(self.frame_words, np.int32)
        return out if lib().bike_emul_read_snapshot(self._h, 0, frame_index, out.ctypes.data) else None

    def tick(self):
        return lib().bike_emul_tick(self._h, 0)

    def counters(self):
        out = np.zeros(4, np.int64)
        lib().bike_emul_counters(self._h, 0, out.ctypes.data)
        return out


class VmEmulEnv:
    """vm_scheduling device logic under the thread-per-lane emulator (one replica)."""

    def __init__(self, topo, snapshot_resolution=1, max_snapshots=None, max_actions=2, lanes=0):
        self._struct, self._keep = _abi.vm_topology_struct(topo)
        cfg = _abi.MaroCimConfig()
        cfg.n_replicas = 1
        cfg.start_tick = topo.start_tick
        cfg.snapshot_resolution = snapshot_resolution
        cfg.max_snapshots = int(max_snapshots) if max_snapshots else 0
        cfg.max_actions = max_actions
        self.A = max_actions
        self._h = lib().vm_emul_create(C.byref(self._struct), C.byref(cfg), lanes)
        assert self._h
        self.dec_words = lib().vm_emul_dec_words(self._h)
        self.frame_words = lib().vm_emul_frame_words(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vm_emul_destroy(self._h)
            self._h = None

    def reset(self):
        lib().vm_emul_reset(self._h)

    def step(self, actions=None):
        dec = np.zeros((1, self.dec_words), np.int32)
        met = np.zeros((1, 16), np.int64)
        if actions is None:
            lib().vm_emul_step(self._h, None, None, dec.ctypes.data, met.ctypes.data)
        else:
            src = np.asarray(actions, np.int32).reshape(1, -1, 4)
            a = np.zeros((1, max(self.A, src.shape[1]), 4), np.int32)
            a[:, :src.shape[1]] = src
            a = np.ascontiguousarray(a[:, :self.A])
            n = np.full(1, src.shape[1], np.int32)
            lib().vm_emul_step(self._h, a.ctypes.data, n.ctypes.data, dec.ctypes.data, met.ctypes.data)
        return int(dec[0, 6]), dec[0], met[0]

    def frame(self):
        out = np.zeros(self.frame_words, np.int32)
        lib().vm_emul_read_frame(self._h, 0, out.ctypes.data)
        return out

    def snapshot(self, frame_index):
        out = np.zeros(self.frame_words, np.int32)
        return out if lib().vm_emul_read_snapshot(self._h, 0, frame_index, out.ctypes.data) else None

    def tick(self):
        return lib().vm_emul_tick(self._h, 0)

    def counters(self):
        out = np.zeros(4, np.int64)
        lib().vm_emul_counters(self._h, 0, out.ctypes.data)
        return out
