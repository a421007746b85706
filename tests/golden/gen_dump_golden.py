"""Frame dump of the UNMODIFIED reference (static backend): FrameBase.dump -> NumpyBackend.dump (np_backend.pyx:391-401).

    bash oracle/build_ref.sh && python tests/golden/gen_dump_golden.py

toy.4p_ssdd_l0.0, 20 ticks, max_snapshots 8, LOAD of half the scope at every second decision.  Output:
tests/golden/dump_toy4p_20/{ports,vessels,matrices}.{npy,meta}."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
TOPOLOGY, DURATIONS, MAX_SNAPSHOTS = "toy.4p_ssdd_l0.0", 20, 8


def drive(env, Action, ActionType):
    metrics, ev, done = env.step(None)
    k = 0
    while not done:
        act = Action(ev.vessel_idx, ev.port_idx, ev.action_scope.load // 2, ActionType.LOAD) if k % 2 == 0 else None
        k += 1
        metrics, ev, done = env.step(act)


if __name__ == "__main__":
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    sys.path[:0] = [os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "oracle", "_ref", "_stubs")]
    from maro.simulator import Env
    from maro.simulator.scenarios.cim.common import Action, ActionType

    env = Env("cim", TOPOLOGY, durations=DURATIONS, max_snapshots=MAX_SNAPSHOTS)
    drive(env, Action, ActionType)
    out = os.path.join(HERE, "dump_toy4p_20")
    os.makedirs(out, exist_ok=True)
    env._business_engine.frame.dump(out)
    print(sorted(os.listdir(out)))
