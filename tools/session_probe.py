"""GPU probe (not a test, not the bench): where the time of a host-driven Env.step call goes.
    python tools/session_probe.py [replicas]
Times maro_cim_step_pinned for (a) every replica inactive (protocol only: command rows out, relay, rows back),
(b) step(None) for every replica, with the resident session on / off and a few poll back-off settings."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(B, env_vars, n=2000):
    for k, v in env_vars.items():
        os.environ[k] = v
    from maro_b200.batch import CimBatch
    from maro_b200.scenarios.cim.topology import build_topology

    topo = build_topology("toy.4p_ssdd_l0.0", 1000)
    env = CimBatch(topo, B)
    pa, pn, pact, pd, pm = env.pinned()
    out = {}
    pact[:] = 0
    env.step_pinned(use_actions=False)
    for _ in range(20):
        env.step_pinned(use_actions=False, use_active=True)
    t0 = time.perf_counter()
    for _ in range(n):
        env.step_pinned(use_actions=False, use_active=True)
    out["inactive_us"] = 1e6 * (time.perf_counter() - t0) / n
    env.reset()
    for _ in range(20):
        env.step_pinned(use_actions=False)
    c0 = env.counters().sum(0)
    t0 = time.perf_counter()
    m = min(n, 600)
    for _ in range(m):
        env.step_pinned(use_actions=False)
    dt = time.perf_counter() - t0
    c1 = env.counters().sum(0)
    out["null_step_us"] = 1e6 * dt / m
    out["ticks_per_step"] = float(c1[1] - c0[1]) / max(1.0, float(c1[0] - c0[0]))
    env.close()
    for k in env_vars:
        os.environ.pop(k, None)
    return out


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    for ev in ({"MARO_B200_SESSION": "0"}, {}, {"MARO_B200_POLL_NS": "200"}, {"MARO_B200_WAIT_NS": "0"},
               {"MARO_B200_WAIT_NS": "100"}, {"MARO_B200_RES_WARPS": "4"}, {"MARO_B200_RES_WARPS": "8"}):
        print(B, ev, run(B, ev), flush=True)
