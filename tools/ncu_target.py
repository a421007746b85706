"""Small driver for ncu captures: a few resident rollouts (or per-step launches) of one CIM configuration.
    python tools/ncu_target.py [topology] [replicas] [ticks] [chunk] [launches] [mode: rollout|step]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from maro_b200.batch import CimBatch  # noqa: E402
from maro_b200.scenarios.cim.topology import build_topology  # noqa: E402

a = sys.argv[1:]
topology = a[0] if len(a) > 0 else "toy.4p_ssdd_l0.0"
B = int(a[1]) if len(a) > 1 else 1024
ticks = int(a[2]) if len(a) > 2 else 1000
chunk = int(a[3]) if len(a) > 3 else 64
launches = int(a[4]) if len(a) > 4 else 3
mode = a[5] if len(a) > 5 else "rollout"
topo = build_topology(topology, ticks)
env = CimBatch(topo, B, max_snapshots=int(os.environ.get("MAX_SNAPSHOTS", "0")) or None)
env.set_stream(torch.cuda.current_stream().cuda_stream)
dec = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
met = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
act = torch.zeros((B, 1, 4), dtype=torch.int32, device="cuda")
for _ in range(launches):
    if mode == "rollout":
        env.rollout_device(dec.data_ptr(), met.data_ptr(), chunk, 1, 0, 0)
    else:
        for _ in range(chunk):
            env.random_policy_device(dec.data_ptr(), act.data_ptr(), 0, 0)
            env.step_device(dec.data_ptr(), met.data_ptr(), act.data_ptr())
torch.cuda.synchronize()
print("counters", env.counters().sum(0).tolist())
