"""Frame / snapshot dump in the reference's on-disk format.

``NumpyBackend.dump`` (maro/backends/np_backend.pyx:391-401, reached through ``FrameBase.dump``, frame.pyx:642-649) writes,
per node type, ``<node>.npy`` — ONE structured array ``[1 + total_snapshots][node_number]`` whose row 0 is the live frame and
whose rows 1.. are the snapshot rows in the ring order of ``NPSnapshotList.take_snapshot`` (np_backend.pyx:481-518: row
``1 + frame % total_snapshots`` for consecutive frames) — and ``<node>.meta`` (attribute names, then slot counts).  The
field dtypes are what the reference derives from the attribute declarations as built here: the decoded AttributeType NAME
is handed to numpy (np_backend.pyx:177, 293), so ``"i"`` -> ``"int"`` -> int64, ``"i2"`` -> ``"short"`` -> int16,
``"f"`` -> ``"float"`` -> float64; fields are in the backend's registration order (alphabetical, frame.pyx:700).

The RawBackend's CSV dumps (raw/frame.cpp:298-369, raw/snapshotlist.cpp:420-487) are not restated: the reference's own
``Frame::dump`` crashes (segmentation fault) in the build of this image, so there is nothing to pin them against.
"""
from __future__ import annotations

import os

import numpy as np

from .. import _abi

# attribute declarations that are not the default "i" (maro/simulator/scenarios/*/{port,vessel,station,physical_machine,...}.py)
_DECL = {
    "cim": {"ports": {"transfer_cost": "f"}, "vessels": {"is_parking": "i2"}, "matrices": {}},
    "citi_bike": {"stations": {"weekday": "i2", "temperature": "i2", "weather": "i2", "holiday": "i2"}, "matrices": {}},
    "vm_scheduling": {
        "pms": {k: "i2" for k in ("cpu_cores_capacity", "memory_capacity", "pm_type", "cpu_cores_allocated", "memory_allocated",
                                  "oversubscribable", "region_id", "zone_id", "data_center_id", "cluster_id")}
        | {"cpu_utilization": "f", "energy_consumption": "f"},
        "racks": {k: "i2" for k in ("region_id", "zone_id", "data_center_id", "cluster_id")},
        "clusters": {k: "i2" for k in ("id", "region_id", "zone_id", "data_center_id")},
        "data_centers": {k: "i2" for k in ("id", "region_id", "zone_id")},
        "zones": {k: "i2" for k in ("id", "region_id")},
        "regions": {"id": "i2"},
    },
}
_NUMPY_NAME = {"i": "int", "i2": "short", "i4": "int", "i8": "long", "f": "float", "d": "double"}  # decoded AttributeType names


def _scenario_of(batch):
    name = type(batch).__name__
    return "vm_scheduling" if "Vm" in name else ("citi_bike" if "Bike" in name else "cim")


def _layout(batch, scenario):
    if scenario == "cim":
        t = batch.topologies[0]
        return _abi.frame_layout(t.n_ports, t.n_vessels, t.past_stop_number, t.future_stop_number)[0]
    if scenario == "citi_bike":
        return _abi.bike_frame_layout(batch.topology.n_stations)[0]
    return _abi.vm_frame_layout(batch.topology)[0]


def dump_snapshots(batch, replica: int, folder: str) -> None:
    """write ``<node>.npy`` + ``<node>.meta`` for every node type of ``batch``'s replica ``replica`` into ``folder``"""
    scenario = _scenario_of(batch)
    if scenario == "vm_scheduling":
        # cpu_utilization / energy_consumption are float64 in the reference's arrays but float32 words in the device ring; the
        # exact lift exists for queries (VmBatch.query) only
        raise NotImplementedError("frame dump is implemented for the cim and citi_bike scenarios")
    layout = _layout(batch, scenario)
    frames = [int(f) for f in batch.snapshot_frames(replica)]
    ring = int(batch.ring_rows())
    rows = {0: batch.read_frame(replica)}
    for f in frames:
        rows[1 + f % ring] = batch.snapshot_row(f, replica)
    float_words = set(getattr(_abi, "VM_FLOAT_ATTRS", ())) if scenario == "vm_scheduling" else {"transfer_cost"} if scenario == "cim" else set()
    for node, attrs in layout.items():
        names = sorted(attrs)
        n_nodes = attrs[names[0]][1]
        fields = []
        for a in names:
            dt = _NUMPY_NAME[_DECL[scenario].get(node, {}).get(a, "i")]
            slots = attrs[a][2]
            fields.append((a, dt) if slots == 1 else (a, dt, slots))
        arr = np.zeros((ring + 1, n_nodes), np.dtype(fields))
        for r, words in rows.items():
            for a in names:
                off, n, slots = attrs[a]
                w = np.asarray(words[off:off + n * slots], np.int32)
                v = w.view(np.float32).astype(np.float64) if a in float_words else w
                arr[r][a] = v.reshape(n, slots) if slots > 1 else v
        with open(os.path.join(folder, node + ".npy"), "wb+") as fp:
            np.save(fp, arr)
        with open(os.path.join(folder, node + ".meta"), "wt+") as fp:
            fp.write(",".join(names) + "\n")
            fp.write(",".join(str(attrs[a][2]) for a in names))
