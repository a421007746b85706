"""Write a synthetic vm_scheduling trace in the MARO .bin schema (same meta blocks as the reference's own fixtures
tests/data/vm_scheduling/*.bin, which are copied verbatim as templates).  Deterministic (numpy seed).  The reference's
real dataset (azure.2019.10k) is a network download and unavailable offline.

    python tests/golden/vm_synth_gen.py        # rewrites tests/golden/vm_synth/
"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/tests/data/vm_scheduling"
HDR = struct.Struct("<4s b I Q I QQ QQ qq")


def _template(path):
    b = open(path, "rb").read()
    h = HDR.unpack_from(b)
    return h, b[h[5]:h[5] + h[6]]  # header tuple, meta bytes


def write_bin(path, template, items: np.ndarray, starttime, endtime):
    h, meta = template
    name, ftype, ver, _, isize, moff, msize, _, _, _, _ = h
    assert items.dtype.itemsize == isize
    doff = moff + msize
    hdr = HDR.pack(name, ftype, ver, len(items), isize, moff, msize, doff, len(items) * isize, starttime, endtime)
    with open(path, "wb") as fp:
        fp.write(hdr + meta + items.tobytes())


def generate(out_dir, n_vm=260, ticks=160, seed=7, split_at=80):
    rng = np.random.default_rng(seed)
    vm_dt = np.dtype([("timestamp", "<i4"), ("vm_id", "<i4"), ("sub_id", "<i4"), ("deploy_id", "<i4"), ("vm_lifetime", "<i4"),
                      ("vm_deleted", "<i4"), ("vm_category", "<i4"), ("vm_cpu_cores", "<i4"), ("vm_memory", "<i4")])
    cpu_dt = np.dtype([("timestamp", "<i4"), ("vm_id", "<i4"), ("cpu_utilization", "<f4")])
    created = np.sort(rng.integers(0, ticks - 10, n_vm)).astype(np.int32)
    life = rng.integers(1, 50, n_vm).astype(np.int32)
    ids = rng.permutation(np.arange(1000, 1000 + 4 * n_vm))[:n_vm].astype(np.int32)
    vms = np.zeros(n_vm, vm_dt)
    vms["timestamp"], vms["vm_id"], vms["sub_id"], vms["deploy_id"] = created, ids, rng.integers(0, 20, n_vm), rng.integers(0, 50, n_vm)
    vms["vm_lifetime"], vms["vm_deleted"], vms["vm_category"] = life, created + life, rng.integers(0, 3, n_vm)
    cores = rng.choice([1, 2, 4, 8, 16], n_vm, p=[0.25, 0.3, 0.25, 0.15, 0.05]).astype(np.int32)
    vms["vm_cpu_cores"], vms["vm_memory"] = cores, cores * rng.choice([2, 4, 8], n_vm)
    rows = []
    for i in range(n_vm):
        base = rng.uniform(2, 70)
        for t in range(created[i], min(ticks + 1, created[i] + life[i] + 14)):
            if t > created[i] and rng.random() < 0.08:
                continue  # missing reading -> the engine repeats the previous one
            u = float(np.clip(base + rng.normal(0, 12), 0.0, 100.0)) if rng.random() > 0.03 else float(rng.uniform(90, 100))
            rows.append((t, int(ids[i]), u))
    cpu = np.array(rows, cpu_dt)
    cpu = cpu[np.argsort(cpu["timestamp"], kind="stable")]
    os.makedirs(out_dir, exist_ok=True)
    write_bin(os.path.join(out_dir, "vmtable_synth.bin"), _template(os.path.join(REF, "vmtable_test.bin")), vms, 0, int(created.max()))
    tpl = _template(os.path.join(REF, "vm_cpu_readings-file-1-of-test.bin"))
    a, b = cpu[cpu["timestamp"] <= split_at], cpu[cpu["timestamp"] >= split_at]
    write_bin(os.path.join(out_dir, "vm_cpu_readings-file-1-of-synth.bin"), tpl, a, 0, split_at)
    write_bin(os.path.join(out_dir, "vm_cpu_readings-file-2-of-synth.bin"), tpl, b, split_at, int(cpu["timestamp"].max()))
    return len(vms), len(cpu)


if __name__ == "__main__":
    print(generate(os.path.join(HERE, "vm_synth")))
