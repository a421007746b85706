"""TEST INFRASTRUCTURE — runs the reference's own example code, UNMODIFIED, either on the reference itself or on the
CUDA core through the ``maro`` import shim (maro_b200/shim.py), and prints one JSON summary.  tests/test_gpu_shim.py
runs it twice (``--mode reference`` / ``--mode shim``) in fresh processes and compares the summaries.

    python tests/run_reference_script.py --mode reference|shim [--emulate] --what hello_cim|hello_vector|rl_cim [--seed K]

The scripts come from oracle/_ref/examples (copied there, unmodified, by oracle/build_ref.sh; git-ignored):
  hello_cim     examples/hello_world/cim/hello.py          (Env, random agent from the `random` module, two episodes)
  hello_vector  examples/vector_env/hello.py               (VectorEnv: dict / list stepping, snapshot_list, reset)
  rl_cim        examples/cim/rl (rl_component_bundle + CIMEnvSampler on maro.rl's AbsEnvSampler): one sample() episode
                with the bundle's DQN policies, then TrainingManager.record_experiences + train_step
``--emulate`` (CPU suite) swaps the CUDA batch for the host emulation of the device code (tests/emul_batch.py).
"""
import argparse
import contextlib
import hashlib
import io
import json
import os
import random
import re
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref")


def digest(arrs) -> str:
    import numpy as np

    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", required=True, choices=["reference", "shim"])
    ap.add_argument("--what", required=True, choices=["hello_cim", "hello_vector", "rl_cim", "rl_cim_greedy", "rl_cim_batched"])
    ap.add_argument("--emulate", action="store_true")
    ap.add_argument("--seed", type=int, default=20240923)
    args = ap.parse_args()
    os.environ["SKIP_DEPLOYMENT"] = "TRUE"
    sys.path[:0] = [REF, os.path.join(REF, "_stubs")]   # maro.rl / maro.utils / examples.* always come from the reference
    if args.mode == "shim":
        sys.path.insert(0, ROOT)
        if args.emulate:
            sys.path.insert(0, HERE)
            import emul_batch
            import maro_b200.simulator.env as env_mod
            import maro_b200.vector_env.vector_env as venv_mod

            for mod in (env_mod, venv_mod):
                mod.CimBatch, mod.BikeBatch, mod.VmBatch = emul_batch.EmulCimBatch, emul_batch.EmulBikeBatch, emul_batch.EmulVmBatch
        import maro_b200.shim

        maro_b200.shim.install()
    random.seed(args.seed)
    out = {"mode": args.mode, "what": args.what}
    buf = io.StringIO()
    if args.what in ("hello_cim", "hello_vector"):
        rel = "hello_world/cim/hello.py" if args.what == "hello_cim" else "vector_env/hello.py"
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(REF, "examples", rel), run_name="__main__")
        text = buf.getvalue()
        if args.what == "hello_cim":
            out["lines"] = [ln for ln in text.splitlines() if ln.startswith("ep:")]
            m = re.search(r"'node_mapping': (\{.*?\}\})", text)
            out["summary_has_node_mapping"] = bool(m)
        else:
            out["lines"] = [ln for ln in text.splitlines() if ln.startswith(("When env 0", "Final"))]
    elif args.what in ("rl_cim_greedy", "rl_cim_batched"):
        # The reference sampler with exploration switched off (so that a batched run can be compared transition by transition)
        # vs maro_b200.rl_rollout.BatchedCimEnvSampler driving 16 replicas with the SAME per-port DQN policies.
        import numpy as np
        import torch

        np.random.seed(args.seed)
        torch.manual_seed(args.seed)
        torch.set_num_threads(1)
        with contextlib.redirect_stdout(buf):
            from examples.cim.rl.rl_component_bundle import rl_component_bundle as bundle
            from maro.rl.training import TrainingManager

            sampler = bundle.env_sampler
            policies = {p.name: p for p in bundle.policies}
            for p in policies.values():
                p.explore = p.exploit  # test driver only: greedy actions in sample() ...
                p._warmup = 0          # ... from the first call on (the bundle's DQN policies act uniformly at random for 100 calls)
            if args.what == "rl_cim_greedy":
                result = sampler.sample()
            else:
                from examples.cim.rl.config import env_conf
                from maro_b200.batch import CimBatch
                from maro_b200.rl_rollout import BatchedCimEnvSampler
                from maro_b200.scenarios.cim.topology import build_topology

                batch = CimBatch(build_topology(env_conf["topology"], env_conf["durations"]), 16)
                a2p = bundle.agent2policy

                def policy(states, decisions):  # per-port policies of the bundle, evaluated for the rows of each port
                    st = states.cpu().numpy().astype(np.float64)
                    ports = decisions[:, 1].cpu().numpy()
                    live = decisions[:, 6].cpu().numpy() == 0
                    out = np.zeros(len(st), np.int64)
                    for port in np.unique(ports[live]):
                        rows = np.flatnonzero(live & (ports == port))
                        pol = policies[a2p[int(port)]]
                        pol.eval()
                        with torch.no_grad():
                            out[rows] = np.asarray(pol.get_actions(st[rows])).reshape(-1)
                    return torch.as_tensor(out, device=states.device)

                bs = BatchedCimEnvSampler(batch, policy, policy_takes_decisions=True, use_graph=False)
                result = bs.sample()
                assert len(result["experiences"]) == 16
                first = result["experiences"][0]
                for r, other in enumerate(result["experiences"][1:], 1):  # identical replicas, greedy policies: identical lists
                    if len(other) != len(first):
                        raise AssertionError(f"replica {r}: {len(other)} experiences, replica 0 has {len(first)}")
                    for k, (a, b) in enumerate(zip(first, other)):
                        if a.tick != b.tick or not np.array_equal(a.state, b.state):
                            raise AssertionError(f"replica {r} transition {k}: tick {b.tick} vs {a.tick}, "
                                                 f"state diff at {np.flatnonzero(np.asarray(a.state) != np.asarray(b.state))[:8].tolist()}")
                result = {"experiences": result["experiences"][:2], "info": result["info"][:2]}
            exps = result["experiences"][0]
            tm = TrainingManager(rl_component_bundle=bundle, explicit_assign_device=True)
            tm.record_experiences(result["experiences"])
            tm.train_step()
            state = tm.get_policy_state()
        agent = lambda e: sorted(e.agent_state_dict)[0]
        out.update({
            "exp_class": type(exps[0]).__module__,
            "n_experiences": len(exps),
            "ticks": [int(e.tick) for e in exps],
            "agents": [int(agent(e)) for e in exps],
            "states": digest([np.asarray(e.state, np.float64) for e in exps]),
            "agent_states": digest([np.asarray(e.agent_state_dict[agent(e)], np.float64) for e in exps]),
            "next_agent_states": digest([np.asarray(e.next_agent_state_dict[agent(e)], np.float64) for e in exps]),
            "actions": [int(np.asarray(e.action_dict[agent(e)]).reshape(-1)[0]) for e in exps],
            "rewards": [float(e.reward_dict[agent(e)]) for e in exps],
            "terminals": [bool(e.terminal_dict[agent(e)]) for e in exps],
            "env_metric": {k: int(v) for k, v in result["info"][0]["env_metric"].items()},
            "trained": sorted(state),
        })
    else:
        import numpy as np
        import torch

        np.random.seed(args.seed)
        torch.manual_seed(args.seed)
        torch.set_num_threads(1)
        with contextlib.redirect_stdout(buf):
            from examples.cim.rl.rl_component_bundle import rl_component_bundle as bundle
            from maro.rl.training import TrainingManager

            sampler = bundle.env_sampler
            result = sampler.sample()
            exps = result["experiences"][0]
            sampler.post_collect(result["info"], 1)
            tm = TrainingManager(rl_component_bundle=bundle, explicit_assign_device=True)
            tm.record_experiences(result["experiences"])
            tm.train_step()
            state = tm.get_policy_state()
        rewards = [float(sum(e.reward_dict.values())) for e in exps]
        out.update({
            "env_class": type(sampler.env).__module__,
            "n_experiences": len(exps),
            "ticks": [int(e.tick) for e in exps],
            "states": digest([e.state for e in exps]),
            "actions": digest([np.asarray([np.asarray(v).ravel() for v in e.action_dict.values()]) for e in exps]),
            "reward_sum": float(np.sum(rewards)),
            "rewards": [round(r, 3) for r in rewards[:50]],
            "env_metric": {k: int(v) for k, v in result["info"][0]["env_metric"].items()},
            "policy_state": digest([t.detach().cpu().numpy() for name in sorted(state) for ps in [state[name]]
                                    for t in (ps.values() if isinstance(ps, dict) else []) if hasattr(t, "detach")]),
        })
    print(json.dumps(out))


if __name__ == "__main__":
    main()
