"""GPU (-m gpu): device-resident RL state / reward shaping (maro_cim_rl_state_device / maro_cim_rl_reward_device,
maro_b200.rl_shaping.CimShaper) against vectors recorded from the unmodified reference and the numpy restatement."""
import numpy as np
import pytest

from rl_helpers import RL_CASES, SnapshotView, load_rl_golden, reward_numpy, state_numpy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(RL_CASES))
def test_device_shaping_matches_reference_vectors(name):
    import torch

    from maro_b200.batch import CimBatch
    from maro_b200.rl_shaping import CimShaper
    from maro_b200.scenarios.cim.topology import build_topology

    spec, gold = RL_CASES[name], load_rl_golden(name)
    topo = build_topology(spec["topology"], spec["durations"])
    B = 4
    env = CimBatch(topo, B, max_snapshots=spec.get("max_snapshots"))
    env.set_stream(torch.cuda.current_stream().cuda_stream)
    shaper = CimShaper(env)
    assert shaper.state_dim == gold["states"].shape[1]
    dec = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
    met = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
    act = torch.zeros((B, 1, 4), dtype=torch.int32, device="cuda")
    env.step_device(dec.data_ptr(), met.data_ptr())
    for k in range(len(gold["steps"])):
        d = dec.cpu().numpy()
        assert d[:, 6].tolist() == [0] * B and d[0, :3].tolist() == gold["steps"][k].tolist()
        s = shaper.states(dec).cpu().numpy()
        assert np.array_equal(s, np.broadcast_to(gold["states"][k], s.shape)), k   # integers -> float64: exact
        model = torch.full((B,), int(gold["model_actions"][k]), dtype=torch.int32, device="cuda")
        act = shaper.env_actions(dec, model)  # the example's action translation, on the device
        assert act.cpu().numpy()[:, 0].tolist() == [gold["actions"][k].tolist()] * B, k
        env.step_device(dec.data_ptr(), met.data_ptr(), act.data_ptr())
    assert dec.cpu().numpy()[:, 6].tolist() == [1] * B
    assert (shaper.states(dec).cpu().numpy() == 0).all()  # no decision pending -> zero rows
    # rewards of every decision, four at a time (replica r scores decision k + r)
    got = np.zeros(len(gold["steps"]), np.float32)
    for k in range(0, len(gold["steps"]), B):
        idx = [min(k + r, len(gold["steps"]) - 1) for r in range(B)]
        ticks = torch.tensor(gold["steps"][idx, 0], dtype=torch.int32, device="cuda")
        ports = torch.tensor(gold["steps"][idx, 1], dtype=torch.int32, device="cuda")
        got[idx] = shaper.rewards(ticks, ports).cpu().numpy()
    # float rewards within 1e-6 (BASELINE north_star): float64 dot products in a different association, cast to float32
    assert np.allclose(got, gold["rewards"], rtol=1e-6, atol=1e-3), np.abs(got - gold["rewards"]).max()
    env.close()


def test_device_shaping_at_batch_size_matches_restatement():
    """1024 replicas with different action histories (hashed random agent on the device; the noise-free schedule keeps
    them at the same tick, their container counts differ): sampled replicas against the restatement."""
    import torch

    from maro_b200.batch import CimBatch
    from maro_b200.rl_shaping import CimShaper
    from maro_b200.scenarios.cim.topology import build_topology

    topo = build_topology("toy.4p_ssdd_l0.0", 400)
    B = 1024
    env = CimBatch(topo, B)
    env.set_stream(torch.cuda.current_stream().cuda_stream)
    shaper = CimShaper(env)
    dec = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
    met = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
    act = torch.zeros((B, 1, 4), dtype=torch.int32, device="cuda")
    env.step_device(dec.data_ptr(), met.data_ptr())
    for _ in range(150):
        env.random_policy_device(dec.data_ptr(), act.data_ptr(), 11, 0)
        env.step_device(dec.data_ptr(), met.data_ptr(), act.data_ptr())
    d = dec.cpu().numpy()
    states = shaper.states(dec).cpu().numpy()
    s32 = torch.full(states.shape, -1.0, dtype=torch.float32, device="cuda")
    assert shaper.states(dec, out=s32) is s32 and np.array_equal(s32.cpu().numpy(), states.astype(np.float32))  # maro_cim_rl_state_f32_device
    # maro_cim_rl_action_ex_device: int64 policy output, int32 record and the running maximum of the metrics in the same launch
    m32 = torch.remainder(dec[:, 7] * 7 + torch.arange(B, device="cuda", dtype=torch.int32), 21).to(torch.int32).contiguous()
    plain = shaper.env_actions(dec, m32).clone()
    record = torch.full((B,), -5, dtype=torch.int32, device="cuda")
    fin = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
    fin[::2] = 1 << 40
    want_fin = torch.maximum(fin, met)
    ex = shaper.env_actions(dec, m32.to(torch.int64), record=record, metrics=met, final_metrics=fin)
    assert torch.equal(ex, plain) and torch.equal(record, m32) and torch.equal(fin, want_fin) and int(met.sum()) > 0
    ticks = torch.tensor(np.maximum(d[:, 0] - 120, 0), dtype=torch.int32, device="cuda")
    ports = torch.tensor(d[:, 1] % topo.n_ports, dtype=torch.int32, device="cuda")
    rewards = shaper.rewards(ticks, ports).cpu().numpy()
    assert len({states[r].tobytes() for r in range(B)}) > B // 2  # the replicas' states differ
    for rep in (0, 1, 511, 1023):
        view = SnapshotView(lambda f, rep=rep: env.snapshot_row(f, rep), topo, cache=True)
        assert d[rep, 6] == 0
        assert np.array_equal(states[rep], state_numpy(view, int(d[rep, 0]), int(d[rep, 1]), int(d[rep, 2]))), rep
        want = reward_numpy(view, int(ports[rep]), int(ticks[rep]))
        assert abs(float(rewards[rep]) - float(want)) <= 1e-6 * max(1.0, abs(float(want))), (rep, rewards[rep], want)
    env.close()


def test_device_rollout_loop_matches_oracle_replay():
    """CimDeviceRollout (state -> torch policy -> action translation -> step, all on the device) for 256 replicas;
    replica 3's trajectory is replayed on the C oracle with the numpy restatement of the same shaping and policy."""
    import torch

    from maro_b200.batch import CimBatch
    from maro_b200.rl_rollout import CimDeviceRollout
    from maro_b200.scenarios.cim.topology import build_topology
    from oracle.cim_oracle import CimOracle
    from rl_helpers import action_numpy

    topo = build_topology("toy.4p_ssdd_l0.0", 300)
    B = 256
    env = CimBatch(topo, B)
    env.set_stream(torch.cuda.current_stream().cuda_stream)
    w = torch.linspace(-1.0, 1.0, 171, device="cuda", dtype=torch.float32)
    rid = torch.arange(B, device="cuda", dtype=torch.float32)

    def policy(states):  # deterministic, replica dependent, exactly representable in float32 and float64
        score = torch.floor((states * w).sum(1).abs() / 64.0) + rid
        return torch.remainder(score, 21).to(torch.int32)

    out = CimDeviceRollout(env, policy).run_episode()
    T_all = out["ticks"].shape[0]
    T = int(out["valid"][:, 0].sum())
    # noise-free schedule: every replica decides at the same steps; the loop runs a fixed, sync-free number of steps (an
    # upper bound of the episode length), the tail answers no-op rows
    assert out["valid"][:T].all() and not out["valid"][T:].any() and out["states"].shape == (T_all, B, 171) and T_all - T <= 4
    rep = 3
    o = CimOracle(topo)
    view = SnapshotView(o.snapshot, topo)
    st, dec, met = o.step(None)
    wn = np.linspace(-1.0, 1.0, 171, dtype=np.float32)
    for t in range(T):
        assert st == 0 and [int(dec[0]), int(dec[1]), int(dec[2])] == [int(out["ticks"][t, rep]), int(out["ports"][t, rep]), int(out["vessels"][t, rep])]
        s = state_numpy(view, int(dec[0]), int(dec[1]), int(dec[2])).astype(np.float32)
        assert np.array_equal(s, out["states"][t, rep].cpu().numpy())
        m = int(out["model_actions"][t, rep])
        assert 0 <= m < 21
        st, dec, met = o.step(np.asarray(action_numpy(view, dec, m), np.int32).reshape(1, 4))
    assert st == 1 and met.tolist() == out["metrics"][rep].cpu().tolist()
    view = SnapshotView(o.snapshot, topo, cache=True)
    for t in (0, T // 2, T - 1):
        want = reward_numpy(view, int(out["ports"][t, rep]), int(out["ticks"][t, rep]))
        got = float(out["rewards"][t, rep])
        assert abs(got - float(want)) <= 1e-6 * max(1.0, abs(float(want))), (t, got, want)
    # the policy depends on replica index mod 21: at least that many distinct trajectories
    assert len({out["model_actions"][:T, r].cpu().numpy().tobytes() for r in range(B)}) >= 16
    env.close()
