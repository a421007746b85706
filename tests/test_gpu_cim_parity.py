"""GPU (-m gpu): the CUDA path, called through the C ABI, against the golden reference traces and the oracle."""
import numpy as np
import pytest

from helpers import CASES, assert_snapshots_equal, case_topology, drive, load_golden, named_frames

pytestmark = pytest.mark.gpu


def _batch(*a, **k):
    from maro_b200.batch import CimBatch

    return CimBatch(*a, **k)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_reference_trace(name):
    spec = CASES[name]
    topo = case_topology(spec)
    gold = load_golden(name)
    B = 5  # identical replicas, same action tape -> every replica must reproduce the reference
    env = _batch(topo, B, spec.get("start_tick", 0), spec.get("snapshot_resolution", 1), spec.get("max_snapshots"),
                 max_actions=2)

    def step_fn(act):
        if act is None:
            dec, met = env.step(None)
        else:
            a = np.zeros((B, 2, 4), np.int32)
            a[:, :1] = np.asarray(act, np.int32).reshape(1, 1, 4)
            dec, met = env.step(a, np.ones(B, np.int32))
        assert (dec == dec[0]).all() and (met == met[0]).all()
        return int(dec[0, 6]), dec[0].copy(), met[0].copy()

    rows, final, dec, st = drive(step_fn, spec)
    assert rows.shape == gold["steps"].shape
    if not np.array_equal(rows, gold["steps"]):
        bad = np.argwhere(rows != gold["steps"])[0]
        raise AssertionError(f"step {bad[0]} col {bad[1]}: got {rows[bad[0]]} want {gold['steps'][bad[0]]}")
    assert final.tolist() == gold["final_metrics"].tolist()
    assert env.ticks().tolist() == [int(gold["final_tick"])] * B
    assert st == 1
    assert step_fn(None)[0] == 2  # finished env -> (None, None, True)
    if "frames" in gold:
        held = env.snapshot_frames(B - 1).tolist()
        if spec.get("keep_frames"):  # the trace keeps every n-th frame only
            assert set(gold["frames"].tolist()) <= set(held)
        else:
            assert held == gold["frames"].tolist()
        assert_snapshots_equal(lambda f: env.snapshot_row(f, B - 1), gold, topo)
    env.close()


def test_cuda_multi_action_known_answers():
    """tests/cim/test_cim_scenario.py:391-435 through the CUDA path ([LOAD 1201, DISCHARGE 1] in one step)."""
    spec = CASES["case22p_200_null"]
    topo = case_topology(spec)
    env = _batch(topo, 2, max_actions=2)
    dec, met = env.step(None)
    assert (dec[0, 3], dec[0, 4], dec[0, 5]) == (1240, 0, 0)
    v, p = int(dec[0, 2]), int(dec[0, 1])
    a = np.zeros((2, 2, 4), np.int32)
    a[:, 0] = [v, p, 1201, 0]
    a[:, 1] = [v, p, 1, 1]
    dec, met = env.step(a, np.full(2, 2, np.int32))
    history = []
    while dec[0, 6] == 0:
        dec, met = env.step(None)
        if dec[0, 6] == 0 and dec[0, 2] == 35:
            fr = named_frames([env.read_frame(1)], topo)
            history.append((int(fr["vessels/full"][0, 35, 0]), int(fr["vessels/empty"][0, 35, 0]),
                            int(fr["vessels/early_discharge"][0, 35, 0])))
    assert history == [(465, 838, 362), (756, 547, 291), (1261, 42, 505), (1303, 0, 42), (1303, 0, 0),
                       (1303, 0, 0), (803, 0, 0)]
    env.close()


@pytest.mark.parametrize("topology,durations,B", [("toy.4p_ssdd_l0.0", 200, 256), ("toy.4p_ssdd_l0.8", 120, 96),
                                                   ("global_trade.22p_l0.8", 40, 24)])
def test_cuda_independent_replicas_match_oracle(topology, durations, B):
    """Every replica follows its own hashed random action tape; sampled replicas are replayed on the oracle."""
    from maro_b200.scenarios.cim.topology import build_topology
    from oracle.cim_oracle import CimOracle, policy_random

    topo = build_topology(topology, durations)
    env = _batch(topo, B)
    sample = sorted(set([0, 1, B // 2, B - 1] + list(range(0, B, max(1, B // 8)))))
    oracles = {i: CimOracle(topo) for i in sample}
    dec, met = env.step(None)
    o_out = {i: o.step(None) for i, o in oracles.items()}
    step = 0
    while (dec[:, 6] == 0).any():
        acts = np.zeros((B, 1, 4), np.int32)
        for i in range(B):
            acts[i, 0] = policy_random(dec[i], 11, i, step)
        for i, (st, d, m) in o_out.items():
            assert d[:7].tolist() == dec[i, :7].tolist(), (step, i, d, dec[i])
            assert m.tolist() == met[i].tolist()
        o_out = {i: o.step(acts[i]) for i, o in oracles.items()}
        dec, met = env.step(acts)
        step += 1
    cnt = env.counters()
    for i, (st, d, m) in o_out.items():
        assert st == 1 and dec[i, 6] == 1
        assert m.tolist() == met[i].tolist()
        assert np.array_equal(env.read_frame(i), oracles[i].frame())
        assert cnt[i].tolist() == oracles[i].counters().tolist()
    env.close()


def test_cuda_subset_stepping_and_bad_action():
    """VectorEnv dict semantics: replicas outside the active mask do not advance (vector_env.py:131-140);
    an over-scope action is flagged where the reference would raise AssertionError (business_engine.py:731,736)."""
    from maro_b200.scenarios.cim.topology import build_topology

    topo = build_topology("toy.4p_ssdd_l0.0", 50)
    env = _batch(topo, 4)
    dec, met = env.step(None)
    t0 = env.ticks().copy()
    act = np.array([1, 0, 1, 0], np.uint8)
    dec, met = env.step(None, active=act)
    assert dec[1, 6] == 3 and dec[3, 6] == 3 and dec[0, 6] == 0
    dec2, _ = env.step(None, active=1 - act)  # now the others catch up
    dec3, _ = env.step(None, active=act)
    dec4, _ = env.step(None, active=1 - act)
    assert env.ticks()[0] == env.ticks()[1]
    bad = np.zeros((4, 1, 4), np.int32)
    d, _ = env.step(None)
    bad[:, 0] = [d[0, 2], d[0, 1], 10 ** 7, 0]  # load far beyond scope
    d, _ = env.step(bad, active=np.array([1, 0, 0, 0], np.uint8))
    assert d[0, 6] == -1
    d, _ = env.step(None)
    assert d[0, 6] == 2 and d[1, 6] == 0
    env.close()


def test_cuda_reset_reproduces_episode():
    from maro_b200.scenarios.cim.topology import build_topology

    topo = build_topology("toy.4p_ssdd_l0.8", 80)
    env = _batch(topo, 8)

    def run():
        out = []
        dec, met = env.step(None)
        while (dec[:, 6] == 0).any():
            out.append((dec[:, :7].copy(), met.copy()))
            dec, met = env.step(None)
        out.append((dec[:, :7].copy(), met.copy()))
        return out

    a = run()
    env.reset()
    b = run()
    assert len(a) == len(b)
    for (d1, m1), (d2, m2) in zip(a, b):
        assert np.array_equal(d1, d2) and np.array_equal(m1, m2)
    env.close()


def test_cuda_full_size_properties():
    """BASELINE config #2 size (1024 replicas x 1000 ticks, random actions): size-independent invariants
    (container conservation, booking = fulfillment + shortage, monotone ticks) + oracle replay of 3 replicas."""
    import torch

    from maro_b200.scenarios.cim.topology import build_topology
    from oracle.cim_oracle import CimOracle, policy_random

    topo = build_topology("toy.4p_ssdd_l0.0", 1000)
    B = 1024
    env = _batch(topo, B, max_snapshots=64)
    env.set_stream(torch.cuda.current_stream().cuda_stream)  # legacy default stream (handle 0), external
    dec = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
    met = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
    act = torch.zeros((B, 1, 4), dtype=torch.int32, device="cuda")
    env.step_device(dec.data_ptr(), met.data_ptr())
    sample = [0, 511, 1023]
    tapes = {i: [] for i in sample}
    n_total = CimOracle(topo).run_episode(0)[0]  # decisions + the final step; static for a given stop table
    for step in range(n_total - 1):
        env.random_policy_device(dec.data_ptr(), act.data_ptr(), 5)
        a = act[sample].cpu().numpy()
        d = dec[sample].cpu().numpy()
        for k, i in enumerate(sample):
            tapes[i].append((d[k].copy(), a[k, 0].copy()))
        env.step_device(dec.data_ptr(), met.data_ptr(), act.data_ptr())
    torch.cuda.synchronize()
    d = dec.cpu().numpy()
    assert (d[:, 6] == 1).all(), np.unique(d[:, 6])
    assert (env.ticks() == 999).all()
    P, V = topo.n_ports, topo.n_vessels
    lay_frames = np.stack([env.read_frame(i) for i in range(0, B, 37)])
    fr = named_frames(lay_frames, topo)
    total = (fr["ports/empty"] + fr["ports/full"] + fr["ports/on_shipper"] + fr["ports/on_consignee"]).sum(1) + \
        (fr["vessels/empty"][:, :, 0] + fr["vessels/full"][:, :, 0]).sum(1)
    assert (total == topo.total_containers).all()
    assert ((fr["ports/acc_booking"] - fr["ports/acc_shortage"]) == fr["ports/acc_fulfillment"]).all()
    m = met.cpu().numpy()
    assert (m[:, 0] == 2000 * 1000).all()  # order_requirements is exogenous for l0.0
    for i in sample:
        o = CimOracle(topo, max_snapshots=64)
        st, od, om = o.step(None)
        for k, (gd, ga) in enumerate(tapes[i]):
            assert od[:7].tolist() == gd[:7].tolist(), (i, k)
            assert policy_random(od, 5, i, k).tolist() == ga.tolist()
            st, od, om = o.step(ga.reshape(1, 4))
        assert st == 1 and om.tolist() == m[i].tolist()
        assert np.array_equal(env.read_frame(i), o.frame())
        for f in (999, 980, 936):
            assert np.array_equal(env.snapshot_row(f, i), o.snapshot(f))
    env.close()


def test_cuda_query_device_matches_host_query():
    """snapshot_list gather left in HBM: the RL state of examples/cim/rl/config.py:10-36
    (7 look-back ticks x ports x 7 attrs) for every replica at once equals the host query, zero padding included."""
    import torch

    from maro_b200.scenarios.cim.topology import build_topology

    topo = build_topology("toy.4p_ssdd_l0.0", 120)
    env = _batch(topo, 64, max_snapshots=16)
    dec, met = env.step(None)
    for _ in range(40):
        dec, met = env.step(None)
    tick = int(dec[0, 0])
    frames = [max(tick - k, -1) if tick - k >= 0 else 10 ** 6 for k in (16, 8, 4, 2, 1, 0, 40)]  # includes evicted frames
    attrs = ["empty", "full", "on_shipper", "on_consignee", "booking", "shortage", "fulfillment"]
    host = env.query("ports", frames, [0, 1, 2, 3], attrs)
    dev = env.query_device("ports", frames, [0, 1, 2, 3], attrs)
    assert dev.is_cuda and dev.dtype == torch.float64
    assert np.array_equal(dev.cpu().numpy(), host)
    assert (host.reshape(64, 7, -1)[:, 6] == 0).all()  # frame beyond the episode -> zeros
    env.close()


def test_cuda_pinned_step_matches_copying_step():
    """maro_cim_step_pinned (inputs / outputs in the library's pinned staging buffers) == maro_cim_step."""
    from maro_b200.scenarios.cim.topology import build_topology
    from oracle.cim_oracle import policy_random

    topo = build_topology("toy.4p_ssdd_l0.8", 90)
    a_env, b_env = _batch(topo, 32), _batch(topo, 32)
    p_act, p_nact, p_active, p_dec, p_met = b_env.pinned()
    dec, met = a_env.step(None)
    b_env.step_pinned(use_actions=False)
    step = 0
    while (dec[:, 6] == 0).any():
        assert np.array_equal(dec, p_dec) and np.array_equal(met, p_met)
        acts = np.zeros((32, 1, 4), np.int32)
        for i in range(32):
            acts[i, 0] = policy_random(dec[i], 3, i, step)
        p_act[:] = acts
        dec, met = a_env.step(acts)
        b_env.step_pinned()
        step += 1
    assert np.array_equal(dec, p_dec) and np.array_equal(met, p_met)
    a_env.close()
    b_env.close()


def test_cuda_config4_full_size_distinct_seeds():
    """BASELINE config #4 at full size on one GPU's share: global_trade.22p_l0.8, 500 ticks, 1 024 replicas, 8 distinct
    topology seeds in ONE handle (replica r runs seed 4096 + r % 8: its own stop tables and MT19937 order / buffer streams),
    hashed random agent per replica.  (a) replica 3 must reproduce the 500-tick trace of the unmodified reference run with
    set_seed(4099) — decisions, metrics and the kept snapshots; (b) 12 sampled replicas across all seeds are replayed on the
    oracle step by step; (c) the fused resident rollout of the same batch ends in the same metrics / frames; (d) container
    conservation and booking = fulfillment + shortage for every replica."""
    import torch

    from maro_b200.scenarios.cim.topology import build_topology, load_config
    from oracle.cim_oracle import CimOracle, policy_random

    spec = CASES["gt22p_l08_500_rand_seed4099"]
    gold = load_golden("gt22p_l08_500_rand_seed4099")
    conf = load_config("global_trade.22p_l0.8")
    K, B, T = 8, 1024, 500
    topos = [build_topology(conf, T, seed=4096 + k) for k in range(K)]
    rt = (np.arange(B) % K).astype(np.int32)
    env = _batch(topos, B, replica_topology=rt)
    env.set_stream(torch.cuda.current_stream().cuda_stream)
    dec = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
    met = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
    act = torch.zeros((B, 1, 4), dtype=torch.int32, device="cuda")
    sample = sorted({3, 0, 1, 2, 4, 5, 6, 7, 515, 777, 1022, 1023})
    sidx = torch.tensor(sample, device="cuda")
    tapes = {i: [] for i in sample}
    mets = {i: [] for i in sample}
    final = torch.zeros_like(met)  # metrics returned with each replica's DONE row (a later step returns the zero FINISHED row)
    env.step_device(dec.data_ptr(), met.data_ptr())
    for step in range(4000):
        d = dec[sidx].cpu().numpy()
        final = torch.where((dec[:, 6] == 1).unsqueeze(1), met, final)
        if step % 16 == 0 and bool((dec[:, 6] != 0).all().item()):
            break
        env.random_policy_device(dec.data_ptr(), act.data_ptr(), 0)
        a = act[sidx].cpu().numpy()
        m = met[sidx].cpu().numpy()
        for k, i in enumerate(sample):
            if d[k, 6] == 0:
                tapes[i].append((d[k].copy(), a[k, 0].copy()))
                mets[i].append(m[k].copy())
        env.step_device(dec.data_ptr(), met.data_ptr(), act.data_ptr())
    torch.cuda.synchronize()
    dn, mn = dec.cpu().numpy(), final.cpu().numpy()
    assert (dn[:, 6] != 0).all() and (env.ticks() == T - 1).all()
    # (a) the reference trace (seed 4099 = replica 3, policy tape (pseed 0, replica 3))
    rows = np.asarray([list(d[:6]) + list(m) for (d, _), m in zip(tapes[3], mets[3])], np.int64)
    assert rows.shape == gold["steps"].shape and np.array_equal(rows, gold["steps"])
    assert mn[3].tolist() == gold["final_metrics"].tolist()
    assert_snapshots_equal(lambda f: env.snapshot_row(f, 3), gold, topos[3])
    assert len({tuple(mn[k]) for k in range(K)}) == K  # the seeds really differ
    # (b) oracle replay
    for i in sample:
        o = CimOracle(topos[rt[i]])
        st, od, om = o.step(None)
        for k, (gd, ga) in enumerate(tapes[i]):
            assert od[:7].tolist() == gd[:7].tolist(), (i, k, od, gd)
            assert om.tolist() == mets[i][k].tolist()
            assert policy_random(od, 0, i, k).tolist() == ga.tolist()
            st, od, om = o.step(ga.reshape(1, 4))
        assert st == 1 and om.tolist() == mn[i].tolist()
        assert np.array_equal(env.read_frame(i), o.frame())
        assert env.counters()[i].tolist() == o.counters().tolist()
        for f in (T - 1, T - 37, 123):
            assert np.array_equal(env.snapshot_row(f, i), o.snapshot(f))
    # (d) invariants over every replica
    frames = np.stack([env.read_frame(i) for i in range(0, B, 9)])
    fr = named_frames(frames, topos[0])
    total = (fr["ports/empty"] + fr["ports/full"] + fr["ports/on_shipper"] + fr["ports/on_consignee"]).sum(1) + \
        (fr["vessels/empty"][:, :, 0] + fr["vessels/full"][:, :, 0]).sum(1)
    assert (total == topos[0].total_containers).all()
    assert ((fr["ports/acc_booking"] - fr["ports/acc_shortage"]) == fr["ports/acc_fulfillment"]).all()
    final_frames = {i: env.read_frame(i) for i in sample}
    # (c) the same episode as fused resident rollouts (agent as a device callback)
    env.reset()
    dec.zero_()
    final2 = torch.zeros_like(met)
    for _ in range(64):
        env.rollout_device(dec.data_ptr(), met.data_ptr(), 64, 1, 0, 0)
        final2 = torch.where((dec[:, 6] == 1).unsqueeze(1), met, final2)  # a fused rollout stops at its replica's DONE row
        if bool((dec[:, 6] != 0).all().item()):
            break
    torch.cuda.synchronize()
    assert np.array_equal(final2.cpu().numpy(), mn)
    for i in sample:
        assert np.array_equal(env.read_frame(i), final_frames[i])
    assert_snapshots_equal(lambda f: env.snapshot_row(f, 3), gold, topos[3])
    env.close()


@pytest.mark.parametrize("name", ["toy4p_l00_160", "toy5p_l03_140_res4", "gt22p_l08_70"])
def test_cuda_joint_mode_matches_reference_trace(name):
    """DecisionMode.Joint (core.py:354-366) on the device: a step returns every decision event of the tick, the answers are
    applied in list order; 3 replicas through the C ABI against traces of the reference run in Joint mode, then the Env façade."""
    import test_cim_joint as tj
    from maro_b200.scenarios.cim.topology import build_topology

    spec = tj.gen.CASES[name]
    topo = build_topology(spec["topology"], spec["durations"])
    B = 3
    env = _batch(topo, B, 0, spec.get("snapshot_resolution", 1), spec.get("max_snapshots"), decision_mode=1)
    A = env.max_actions

    def step_fn(answers):
        if answers is None:
            dec, met = env.step(None)
        else:
            a = np.zeros((B, A, 4), np.int32)
            for k, x in enumerate(answers):
                a[:, k] = (0, 0, 0, 2) if x is None else x
            dec, met = env.step(a, np.full(B, len(answers), np.int32))
        assert (dec == dec[0]).all() and (met == met[0]).all()
        rows = dec[0].reshape(-1, 8)
        st, n = int(rows[0, 6]), 0
        while st == 0 and n < len(rows) and rows[n, 6] == 0:
            n += 1
        return st, rows[:n].copy() if st == 0 else rows[:1].copy(), met[0].copy()

    rows, mets, final, st = tj.drive_joint(step_fn, spec)
    assert st == 1
    gold = tj.check_against_gold(name, rows, mets, final, None)
    frames = gold["frames"].tolist()
    got = env.query("ports", frames, list(range(topo.n_ports)), ["empty", "full", "shortage"], [B - 1])[0].reshape(len(frames), -1, 3)
    assert np.array_equal(got[:, :, 0], gold["ports/empty"]) and np.array_equal(got[:, :, 1], gold["ports/full"])
    assert np.array_equal(got[:, :, 2], gold["ports/shortage"])
    env.close()
    if name == "toy5p_l03_140_res4":
        tj.env_joint_case(name)
    if name == "toy4p_l00_160":
        tj.vector_env_joint_case(name)
