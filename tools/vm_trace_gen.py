"""Synthetic vm_scheduling trace at azure.2019.10k scale, written in the MARO .bin schema (vmtable + one cpu-readings
file).  The real dataset is a network download (vm_scheduling/topologies/azure.2019.10k/config.yml:14) and unavailable
offline, so BASELINE config #5 runs on this trace: same item layout, 10 000 VMs / 8 638 ticks / 100 PMs by default.
The binary meta blocks are taken from the small fixtures under tests/golden/vm_synth (same schema).  Deterministic.

    python tools/vm_trace_gen.py OUT_DIR [n_vm] [ticks]
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEMPLATES = os.path.join(ROOT, "tests", "golden", "vm_synth")
HDR = struct.Struct("<4s b I Q I QQ QQ qq")


def _template(path):
    b = open(path, "rb").read()
    h = HDR.unpack_from(b)
    return h, b[h[5]:h[5] + h[6]]


def _write_bin(path, template, items, starttime, endtime):
    h, meta = template
    name, ftype, ver, _, isize, moff, msize = h[:7]
    assert items.dtype.itemsize == isize
    hdr = HDR.pack(name, ftype, ver, len(items), isize, moff, msize, moff + msize, len(items) * isize, starttime, endtime)
    with open(path, "wb") as fp:
        fp.write(hdr + meta)
        fp.write(items.tobytes())


def generate(out_dir, n_vm=10000, ticks=8638, seed=2019, mean_concurrent_cores=2300.0):
    """-> (vm table path, cpu readings path).  Lifetimes are scaled so that about `mean_concurrent_cores` cores are
    requested at any time (100 PMs x 32 cores = 3 200 in the azure.2019.10k topology: busy but rarely full)."""
    vm_path = os.path.join(out_dir, "vmtable.bin")
    cpu_path = os.path.join(out_dir, "vm_cpu_readings-file-1-of-1.bin")
    stamp = os.path.join(out_dir, f".done_{n_vm}_{ticks}_{seed}")
    if os.path.exists(stamp):
        return vm_path, cpu_path
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(seed)
    vm_dt = np.dtype([("timestamp", "<i4"), ("vm_id", "<i4"), ("sub_id", "<i4"), ("deploy_id", "<i4"), ("vm_lifetime", "<i4"),
                      ("vm_deleted", "<i4"), ("vm_category", "<i4"), ("vm_cpu_cores", "<i4"), ("vm_memory", "<i4")])
    cpu_dt = np.dtype([("timestamp", "<i4"), ("vm_id", "<i4"), ("cpu_utilization", "<f4")])
    created = np.sort(rng.integers(0, max(1, ticks - 10), n_vm)).astype(np.int64)
    cores = rng.choice([1, 2, 4, 8], n_vm, p=[0.4, 0.35, 0.2, 0.05]).astype(np.int64)
    # heavy-tailed lifetimes (most VMs short, a few live for days), scaled to the concurrency target
    raw = rng.lognormal(0.0, 1.3, n_vm)
    scale = mean_concurrent_cores * ticks / float((raw * cores).sum())
    life = np.clip(np.rint(raw * scale), 1, ticks).astype(np.int64)
    ids = rng.permutation(np.arange(1, 4 * n_vm + 1))[:n_vm].astype(np.int64)
    vms = np.zeros(n_vm, vm_dt)
    vms["timestamp"], vms["vm_id"] = created, ids
    vms["sub_id"], vms["deploy_id"] = rng.integers(0, 500, n_vm), rng.integers(0, 2000, n_vm)
    vms["vm_lifetime"], vms["vm_deleted"], vms["vm_category"] = life, created + life, rng.integers(0, 3, n_vm)
    vms["vm_cpu_cores"], vms["vm_memory"] = cores, cores * rng.choice([2, 4], n_vm)
    # one reading per VM per tick from its creation until a few ticks after its deletion; 2% of the later ones missing
    span = np.minimum(life + 4, ticks + 1 - created)
    owner = np.repeat(np.arange(n_vm), span)
    first = np.repeat(np.cumsum(span) - span, span)
    t = created[owner] + (np.arange(len(owner)) - first)
    base = rng.uniform(3, 45, n_vm)[owner]
    u = np.clip(base + rng.normal(0, 9, len(owner)), 0.0, 100.0)
    spike = rng.random(len(owner)) < 0.01
    u[spike] = rng.uniform(80, 100, int(spike.sum()))
    keep = (t == created[owner]) | (rng.random(len(owner)) >= 0.02)
    cpu = np.zeros(int(keep.sum()), cpu_dt)
    cpu["timestamp"], cpu["vm_id"], cpu["cpu_utilization"] = t[keep], ids[owner][keep], u[keep].astype(np.float32)
    cpu = cpu[np.argsort(cpu["timestamp"], kind="stable")]
    _write_bin(vm_path, _template(os.path.join(TEMPLATES, "vmtable_synth.bin")), vms, 0, int(created.max()))
    _write_bin(cpu_path, _template(os.path.join(TEMPLATES, "vm_cpu_readings-file-1-of-synth.bin")), cpu, 0,
               int(cpu["timestamp"].max()))
    open(stamp, "w").write("ok")
    return vm_path, cpu_path


def azure_like_config(vm_path, cpu_path, n_pm=100):
    """The azure.2019.10k topology (config.yml:1-60): racks of 10 x (32 cores, 128 GB) in one cluster."""
    return {
        "BUFFER_TIME_BUDGET": 0, "DELAY_DURATION": 1, "TICKS_PER_HOUR": 12, "VM_TABLE": vm_path, "CPU_READINGS": cpu_path,
        "PROCESSED_DATA_URL": "", "KILL_ALL_VMS_IF_OVERLOAD": True, "MAX_CPU_OVERSUBSCRIPTION_RATE": 1,
        "MAX_MEM_OVERSUBSCRIPTION_RATE": 1, "MAX_UTILIZATION_RATE": 1, "PRICE_PER_CPU_CORES_PER_HOUR": 0.0698,
        "PRICE_PER_MEMORY_PER_HOUR": 0.0078, "UNIT_ENERGY_PRICE_PER_KWH": 0.07, "POWER_USAGE_EFFICIENCY": 1.7,
        "components": {
            "pm": [{"pm_type": 0, "cpu": 32, "memory": 128,
                    "power_curve": {"calibration_parameter": 1.4, "busy_power": 185, "idle_power": 120}}],
            "rack": [{"type": "a", "pm": [{"pm_type": 0, "pm_amount": 10}]}],
            "cluster": [{"type": "JP1", "rack": [{"rack_type": "a", "rack_amount": max(1, n_pm // 10)}]}]},
        "architecture": {"region": [{"name": "APAC", "zone": [{"name": "asia-northeast1", "data_center": [
            {"name": "Japan", "cluster": [{"type": "JP1", "cluster_amount": 1}]}]}]}]},
    }


if __name__ == "__main__":
    out = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    tk = int(sys.argv[3]) if len(sys.argv) > 3 else 8638
    print(generate(out, n, tk))
