"""Golden RL state / reward vectors from the UNMODIFIED reference (oracle/_ref): the shaping of the reference's CIM RL
example (examples/cim/rl/config.py, env_sampler.py:15-36 state, :66-80 reward) evaluated on ``maro.simulator.Env``.

    bash oracle/build_ref.sh && python tests/golden/gen_cim_rl_golden.py

Per decision: (tick, port, vessel), the 171-dim float64 state, a hashed model action (index into the example's 21-entry
action space) and the env Action the example's ``_translate_to_env_action`` (:38-64) makes of it — that Action drives
the episode; after the episode the float32 reward of every decision (``time_window`` snapshots after the action tick; frames past the end read as zeros).
Output: tests/golden/cim_rl_<case>.npz.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
from gen_cim_golden import hash_u32  # noqa: E402

# examples/cim/rl/config.py
PORT_ATTRIBUTES = ["empty", "full", "on_shipper", "on_consignee", "booking", "shortage", "fulfillment"]
VESSEL_ATTRIBUTES = ["empty", "full", "remaining_space"]
LOOK_BACK = 7
REWARD = dict(time_window=99, fulfillment_factor=1.0, shortage_factor=1.0, time_decay=0.97)
ACTION_SPACE = [(i - 10) / 10 for i in range(21)]
FINITE_VESSEL_SPACE, HAS_EARLY_DISCHARGE = True, True


def model_action_of(seed, replica, step):
    return hash_u32(seed ^ hash_u32(replica * 0x9E3779B9 + step * 0x85EBCA6B + 0x51ED27)) % len(ACTION_SPACE)


def env_action_of(env, event, model_action):
    """examples/cim/rl/env_sampler.py:38-64 -> (vessel, port, quantity, type: 0 load / 1 discharge)"""
    vsl_idx, action_scope = event.vessel_idx, event.action_scope
    vsl_snapshots = env.snapshot_list["vessels"]
    vsl_space = vsl_snapshots[env.tick:vsl_idx:VESSEL_ATTRIBUTES][2] if FINITE_VESSEL_SPACE else float("inf")
    percent = abs(ACTION_SPACE[model_action])
    zero_action_idx = len(ACTION_SPACE) / 2
    if model_action < zero_action_idx:
        return vsl_idx, event.port_idx, int(min(round(percent * action_scope.load), vsl_space)), 0
    early_discharge = vsl_snapshots[env.tick:vsl_idx:"early_discharge"][0] if HAS_EARLY_DISCHARGE else 0
    plan_action = percent * (action_scope.discharge + early_discharge) - early_discharge
    actual = round(plan_action) if plan_action > 0 else round(percent * action_scope.discharge)
    return vsl_idx, event.port_idx, int(actual), 1

CASES = {
    "toy4p_l00_560": dict(topology="toy.4p_ssdd_l0.0", durations=560, pseed=0, replica=0),   # the example's env_conf
    "toy5p_l03_200_ring40": dict(topology="toy.5p_ssddd_l0.3", durations=200, pseed=1, replica=2, max_snapshots=40),
}


def state_of(env, event):
    tick = env.tick
    vessel_snapshots, port_snapshots = env.snapshot_list["vessels"], env.snapshot_list["ports"]
    ticks = [max(0, tick - rt) for rt in range(LOOK_BACK - 1)]
    future_port_list = vessel_snapshots[tick:event.vessel_idx:"future_stop_list"].astype("int")
    return np.concatenate([port_snapshots[ticks:[event.port_idx] + list(future_port_list):PORT_ATTRIBUTES],
                           vessel_snapshots[tick:event.vessel_idx:VESSEL_ATTRIBUTES]])


def reward_of(env, port, tick):
    start = tick + 1
    ticks = list(range(start, start + REWARD["time_window"]))
    ps = env.snapshot_list["ports"]
    ff = ps[ticks:[port]:"fulfillment"].reshape(len(ticks), -1)
    fs = ps[ticks:[port]:"shortage"].reshape(len(ticks), -1)
    decay = [REWARD["time_decay"] ** i for i in range(REWARD["time_window"])]
    return np.float32(REWARD["fulfillment_factor"] * np.dot(ff.T, decay) - REWARD["shortage_factor"] * np.dot(fs.T, decay))[0]


def run_case(name, spec):
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    sys.path.insert(1, os.path.join(ROOT, "oracle", "_ref", "_stubs"))
    from maro.simulator import Env
    from maro.simulator.scenarios.cim.common import Action, ActionType

    env = Env("cim", spec["topology"], durations=spec["durations"], max_snapshots=spec.get("max_snapshots"))
    rows, states, acts, models = [], [], [], []
    metrics, dec, done = env.step(None)
    step = 0
    while not done:
        rows.append([env.tick, dec.port_idx, dec.vessel_idx])
        states.append(state_of(env, dec))
        m = model_action_of(spec["pseed"], spec["replica"], step)
        v, p, q, t = env_action_of(env, dec, m)
        models.append(m)
        acts.append([v, p, q, t])
        step += 1
        metrics, dec, done = env.step(Action(v, p, q, ActionType.DISCHARGE if t else ActionType.LOAD))
    rewards = [reward_of(env, p, t) for t, p, _ in rows]
    np.savez_compressed(os.path.join(HERE, f"cim_rl_{name}.npz"), steps=np.asarray(rows, np.int32),
                        states=np.asarray(states, np.float64), actions=np.asarray(acts, np.int32),
                        model_actions=np.asarray(models, np.int32),
                        rewards=np.asarray(rewards, np.float32))
    print(name, len(rows), "decisions, state dim", len(states[0]), "reward range", float(min(rewards)), float(max(rewards)))


if __name__ == "__main__":
    mp.set_start_method("spawn")
    for name, spec in CASES.items():
        p = mp.Process(target=run_case, args=(name, spec))
        p.start()
        p.join()
        assert p.exitcode == 0, name
