"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/vm_oracle.c (CPU restatement of the vm_scheduling path)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libvm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vm_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                               "-I", os.path.join(_HERE, "..", "include"), src, "-o", _LIB_PATH, "-lm"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        _lib.vm_oracle_create.restype = vp
        _lib.vm_oracle_create.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        for n in ("vm_oracle_destroy", "vm_oracle_reset", "vm_oracle_frame_words", "vm_oracle_tick", "vm_oracle_decision_words"):
            getattr(_lib, n).argtypes = [vp]
        _lib.vm_oracle_step.argtypes = [vp, vp, C.c_int, vp, vp]
        _lib.vm_oracle_read_frame.argtypes = [vp, vp]
        _lib.vm_oracle_counters.argtypes = [vp, vp]
        _lib.vm_oracle_read_snapshot.argtypes = [vp, C.c_int, vp]
        _lib.vm_oracle_run_episode.restype = C.c_int64
        _lib.vm_oracle_run_episode.argtypes = [vp, C.c_int, vp]
        _lib.vm_policy_best_fit.argtypes = [vp, vp, vp]
    return _lib


class VmOracle:
    def __init__(self, topo, snapshot_resolution: int = 1, max_snapshots=None):
        from maro_b200._abi import vm_topology_struct

        if topo.error:
            raise Exception(topo.error)
        self._struct, self._keep = vm_topology_struct(topo)
        self._h = lib().vm_oracle_create(C.byref(self._struct), topo.start_tick, snapshot_resolution,
                                         int(max_snapshots) if max_snapshots else 0)
        self.frame_words = lib().vm_oracle_frame_words(self._h)
        self.dec_words = lib().vm_oracle_decision_words(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vm_oracle_destroy(self._h)
            self._h = None

    def reset(self):
        lib().vm_oracle_reset(self._h)

    def step(self, actions=None):
        dec = np.zeros(self.dec_words, np.int32)
        met = np.zeros(16, np.int64)
        if actions is None:
            st = lib().vm_oracle_step(self._h, None, 0, dec.ctypes.data, met.ctypes.data)
        else:
            a = np.ascontiguousarray(actions, np.int32).reshape(-1, 4)
            st = lib().vm_oracle_step(self._h, a.ctypes.data, a.shape[0], dec.ctypes.data, met.ctypes.data)
        return st, dec, met

    @property
    def tick(self):
        return lib().vm_oracle_tick(self._h)

    def frame(self):
        out = np.zeros(self.frame_words, np.int32)
        lib().vm_oracle_read_frame(self._h, out.ctypes.data)
        return out

    def snapshot(self, frame_index):
        out = np.zeros(self.frame_words, np.int32)
        return out if lib().vm_oracle_read_snapshot(self._h, frame_index, out.ctypes.data) else None

    def counters(self):
        out = np.zeros(4, np.int64)
        lib().vm_oracle_counters(self._h, out.ctypes.data)
        return out

    def best_fit(self, dec):
        act = np.zeros(4, np.int32)
        d = np.ascontiguousarray(dec, np.int32)
        lib().vm_policy_best_fit(self._h, d.ctypes.data, act.ctypes.data)
        return act

    def run_episode(self, policy=1):
        met = np.zeros(16, np.int64)
        return int(lib().vm_oracle_run_episode(self._h, policy, met.ctypes.data)), met
