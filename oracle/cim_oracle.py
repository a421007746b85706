"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/cim_oracle.c (the CPU restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcim_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cim_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                               "-I", os.path.join(_HERE, "..", "include"), src, "-o", _LIB_PATH, "-lm"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()  # rebuilds when the C source is newer than the library
        _lib = C.CDLL(_LIB_PATH)
        _lib.cim_oracle_create.restype = C.c_void_p
        _lib.cim_oracle_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.cim_oracle_destroy.argtypes = [C.c_void_p]
        _lib.cim_oracle_reset.argtypes = [C.c_void_p]
        _lib.cim_oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.cim_oracle_frame_words.argtypes = [C.c_void_p]
        _lib.cim_oracle_read_frame.argtypes = [C.c_void_p, C.c_void_p]
        _lib.cim_oracle_tick.argtypes = [C.c_void_p]
        _lib.cim_oracle_counters.argtypes = [C.c_void_p, C.c_void_p]
        _lib.cim_oracle_read_snapshot.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.cim_oracle_run_episode.restype = C.c_int64
        _lib.cim_oracle_run_episode.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.cim_oracle_step_joint.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.cim_policy_random.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    return _lib


class CimOracle:
    """One replica of the CIM scenario on the CPU oracle."""

    def __init__(self, topo, start_tick: int = 0, snapshot_resolution: int = 1, max_snapshots=None):
        from maro_b200._abi import topology_struct  # struct layout only (a data format, not compute)

        self._topo = topo
        self._struct, self._keep = topology_struct(topo)
        self._h = lib().cim_oracle_create(C.byref(self._struct), start_tick, snapshot_resolution,
                                          int(max_snapshots) if max_snapshots else 0)
        self.frame_words = lib().cim_oracle_frame_words(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cim_oracle_destroy(self._h)
            self._h = None

    def reset(self):
        lib().cim_oracle_reset(self._h)

    def step(self, actions=None):
        """actions: None or array-like [n][4] {vessel, port, qty, type}.  Returns (status, dec[8], metrics[3])."""
        dec = np.zeros(8, np.int32)
        met = np.zeros(3, np.int64)
        if actions is None:
            st = lib().cim_oracle_step(self._h, None, 0, dec.ctypes.data, met.ctypes.data)
        else:
            a = np.ascontiguousarray(actions, np.int32).reshape(-1, 4)
            st = lib().cim_oracle_step(self._h, a.ctypes.data, a.shape[0], dec.ctypes.data, met.ctypes.data)
        return st, dec, met

    def step_joint(self, answers=None):
        """DecisionMode.Joint: answers = list of (vessel, port, qty, type) or None per pending decision (may be shorter).
        Returns (status, decision rows [n][8], metrics[3])."""
        V = self._topo.n_vessels
        dec = np.zeros((V, 8), np.int32)
        met = np.zeros(3, np.int64)
        n = C.c_int32()
        rows = np.zeros((max(1, len(answers or [])), 4), np.int32)
        for k, a in enumerate(answers or []):
            rows[k] = (0, 0, 0, 2) if a is None else a
        st = lib().cim_oracle_step_joint(self._h, rows.ctypes.data, len(answers or []), dec.ctypes.data, C.byref(n), met.ctypes.data)
        return st, dec[:n.value] if st == 0 else dec[:1], met

    @property
    def tick(self) -> int:
        return lib().cim_oracle_tick(self._h)

    def frame(self) -> np.ndarray:
        out = np.zeros(self.frame_words, np.int32)
        lib().cim_oracle_read_frame(self._h, out.ctypes.data)
        return out

    def snapshot(self, frame_index: int):
        out = np.zeros(self.frame_words, np.int32)
        ok = lib().cim_oracle_read_snapshot(self._h, frame_index, out.ctypes.data)
        return out if ok else None

    def counters(self) -> np.ndarray:
        out = np.zeros(4, np.int64)
        lib().cim_oracle_counters(self._h, out.ctypes.data)
        return out

    def run_episode(self, policy: int = 0, seed: int = 0, replica: int = 0):
        met = np.zeros(3, np.int64)
        n = lib().cim_oracle_run_episode(self._h, policy, seed, replica, met.ctypes.data)
        return int(n), met


def policy_random(dec: np.ndarray, seed: int, replica: int, step: int) -> np.ndarray:
    act = np.zeros(4, np.int32)
    d = np.ascontiguousarray(dec, np.int32)
    lib().cim_policy_random(d.ctypes.data, seed, replica, step, act.ctypes.data)
    return act
