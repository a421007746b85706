"""TEST / BENCH INFRASTRUCTURE — the unmodified reference's single in-process Env loop, one host core.
    python tools/ref_inprocess.py <scenario> <topology> <ticks> <seconds> <static|dynamic>
Prints one JSON object {value (env-steps/s), episodes, backend}.  The agents are the ones bench.py's reference arm uses."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scenario, topology, ticks, seconds, backend = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4]), sys.argv[5]
os.environ["SKIP_DEPLOYMENT"] = "TRUE"
os.environ["DEFAULT_BACKEND_NAME"] = backend
sys.path[:0] = [os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "oracle", "_ref", "_stubs"), ROOT]
from maro.simulator import Env  # noqa: E402

if scenario == "cim":
    from maro.simulator.scenarios.cim.common import Action, ActionType
    from tools.workloads import cim_policy_random

    env = Env("cim", topology, durations=ticks)

    def agent(d, step):
        row = [d.tick, d.port_idx, d.vessel_idx, d.action_scope.load, d.action_scope.discharge, d.early_discharge]
        v, p, q, t = cim_policy_random(row, 0, 0, step)
        return Action(v, p, q, ActionType.DISCHARGE if t else ActionType.LOAD)
else:
    from maro.simulator.scenarios.citi_bike.common import Action, DecisionType
    from tools.workloads import bike_greedy

    env = Env("citi_bike", topology, durations=ticks, snapshot_resolution=10)

    def agent(d, step):
        v, cand = bike_greedy(d)
        return Action(d.station_idx, cand, int(v)) if d.type == DecisionType.Supply else Action(cand, d.station_idx, int(v))

steps = eps = 0
t0 = time.perf_counter()
while time.perf_counter() - t0 < seconds:
    env.reset(keep_seed=True) if scenario == "cim" else env.reset()
    _, dec, done = env.step(None)
    steps += 1
    k = 0
    while not done:
        _, dec, done = env.step(agent(dec, k))
        k += 1
        steps += 1
    eps += 1
dt = time.perf_counter() - t0
print(json.dumps({"value": steps / dt, "unit": "env-steps/s", "cores": 1, "backend": backend, "episodes": eps,
                  "what": f"maro.simulator.Env loop, {scenario} {os.path.basename(topology)}, {ticks} ticks, in-process"}))
