"""GPU (-m gpu): the resident kernel (replica blocks stay in shared memory across env-steps) against the per-step
kernel and the oracle.  Two modes: fused K-step rollouts with a device agent (maro_cim_rollout_device) and the
host-driven session behind maro_cim_step / maro_cim_step_pinned (command rows in mapped pinned memory)."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _topo(name, ticks):
    from maro_b200.scenarios.cim.topology import build_topology

    return build_topology(name, ticks)


def _batch(*a, **k):
    from maro_b200.batch import CimBatch

    return CimBatch(*a, **k)


def _ring_rows(env, replica, topo):
    """every snapshot row the ring holds for one replica, as raw frame words"""
    return {int(f): env.snapshot_row(int(f), replica) for f in env.snapshot_frames(replica)}


@pytest.mark.parametrize("slice_steps", [None, 3])            # 3: sliced launches forced (lane groups pull (slice, replica) tickets)
@pytest.mark.parametrize("topology,ticks,B,chunks", [
    ("toy.4p_ssdd_l0.0", 200, 256, [1, 1, 7, 50, 3, 1000]),
    ("toy.4p_ssdd_l0.8", 120, 96, [5, 40, 1000]),          # general kernel: MT19937 order / buffer noise
    ("global_trade.22p_l0.8", 40, 24, [3, 1000]),           # one warp = one replica, 30 KB blocks
])
def test_rollout_matches_step_loop_and_oracle(topology, ticks, B, chunks, slice_steps, monkeypatch):
    import torch

    from oracle.cim_oracle import CimOracle, policy_random

    topo = _topo(topology, ticks)
    seed, base = 11, 5
    # ---- reference run on the device: policy kernel + per-step kernel
    a = _batch(topo, B)
    s = torch.cuda.current_stream().cuda_stream
    a.set_stream(s)
    dec = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
    met = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
    act = torch.zeros((B, 1, 4), dtype=torch.int32, device="cuda")
    rows, mets = [], []
    for _ in range(100000):
        a.random_policy_device(dec.data_ptr(), act.data_ptr(), seed, base)
        a.step_device(dec.data_ptr(), met.data_ptr(), act.data_ptr())
        rows.append(dec.cpu().numpy().copy())
        mets.append(met.cpu().numpy().copy())
        if (rows[-1][:, 6] == 2).all():
            break
    want = np.stack(rows)
    done_at = np.argmax(want[:, :, 6] == 1, axis=0)  # per replica: the step that returned its DONE row
    want_met = np.stack([mets[done_at[i]][i] for i in range(B)])
    # ---- fused rollouts on a second handle, uneven chunk sizes
    if slice_steps:
        monkeypatch.setenv("MARO_B200_RES_SLICE_STEPS", str(slice_steps))
    b = _batch(topo, B)
    b.set_stream(s)
    dec2 = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
    met2 = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
    got, done = [], 0
    got_met = np.zeros((B, 3), np.int64)
    for n in chunks:
        n = min(n, len(want) - done)
        if n <= 0:
            break
        trace = torch.full((n, B, 8), -7, dtype=torch.int32, device="cuda")
        b.rollout_device(dec2.data_ptr(), met2.data_ptr(), n, 1, seed, base, trace.data_ptr())
        got.append(trace.cpu().numpy())
        done += n
        d2, m2 = dec2.cpu().numpy(), met2.cpu().numpy()
        got_met[d2[:, 6] == 1] = m2[d2[:, 6] == 1]  # a fused rollout stops at the replica's DONE row (final metrics kept)
    got = np.concatenate(got)
    assert done == len(want)
    for i in range(B):  # identical up to and including the DONE row; afterwards the launch repeats it / a new launch answers FINISHED
        k = done_at[i] + 1
        if not np.array_equal(got[:k, i], want[:k, i]):
            bad = np.argwhere(got[:k, i] != want[:k, i])[0]
            raise AssertionError(f"step {bad[0]} replica {i}: got {got[bad[0], i]} want {want[bad[0], i]}")
        assert set(got[k:, i, 6].tolist()) <= {1, 2}
    assert (dec2.cpu().numpy()[:, 6] != 0).all()
    assert np.array_equal(got_met, want_met)
    assert np.array_equal(a.counters(), b.counters())
    assert np.array_equal(a.ticks(), b.ticks())
    for i in sorted({0, 1, B // 2, B - 1}):
        assert np.array_equal(a.read_frame(i), b.read_frame(i))
        ra, rb = _ring_rows(a, i, topo), _ring_rows(b, i, topo)
        assert ra.keys() == rb.keys()
        for f in ra:
            assert np.array_equal(ra[f], rb[f]), (i, f)
    # ---- and the oracle on sampled replicas, replaying the traced decisions
    for i in sorted({0, B // 3, B - 1}):
        o = CimOracle(topo)
        st, d, m = o.step(None)
        k = 0
        while st == 0:
            assert d[:7].tolist() == got[k, i, :7].tolist(), (i, k, d, got[k, i])
            st, d, m = o.step(np.asarray([policy_random(got[k, i], seed, i + base, int(got[k, i, 7]))], np.int32))
            k += 1
        assert st == 1 and got[k, i, 6] == 1 and m.tolist() == mets[k][i].tolist()
        assert np.array_equal(b.read_frame(i), o.frame())
    a.close()
    b.close()


def test_sliced_rollouts_of_a_grid_larger_than_the_gpu_equal_whole_rollouts(monkeypatch):
    """BASELINE config #4 at 1 024 replicas: 171 CTAs of six 37 KB blocks on 148 SMs.  The launch slices itself (resident lane
    groups pull (slice, replica) tickets, a replica's slices hand the block over through global memory, possibly across SMs);
    results must equal the unsliced launch bit for bit — decisions, metrics, counters, frames, snapshot rings."""
    import torch

    topo, B, seed = _topo("global_trade.22p_l0.8", 60), 1024, 3
    s = torch.cuda.current_stream().cuda_stream
    out = []
    for forced in ("0", None):
        if forced is None:
            monkeypatch.delenv("MARO_B200_RES_SLICE_STEPS", raising=False)
        else:
            monkeypatch.setenv("MARO_B200_RES_SLICE_STEPS", forced)
        env = _batch(topo, B)
        env.set_stream(s)
        dec = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
        met = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
        traces = []
        for _ in range(40):
            trace = torch.full((32, B, 8), -7, dtype=torch.int32, device="cuda")
            env.rollout_device(dec.data_ptr(), met.data_ptr(), 32, 1, seed, 0, trace.data_ptr())
            traces.append(trace.cpu().numpy())
            if (dec.cpu().numpy()[:, 6] != 0).all():
                break
        assert (dec.cpu().numpy()[:, 6] != 0).all()
        out.append((np.concatenate(traces), dec.cpu().numpy(), met.cpu().numpy(), env.counters(), env.ticks(),
                    [env.read_frame(i) for i in range(0, B, 37)], [_ring_rows(env, i, topo) for i in (0, 500, B - 1)]))
        env.close()
    whole, sliced = out
    for k in range(5):
        assert np.array_equal(whole[k], sliced[k]), k
    for fa, fb in zip(whole[5], sliced[5]):
        assert np.array_equal(fa, fb)
    for ra, rb in zip(whole[6], sliced[6]):
        assert ra.keys() == rb.keys() and all(np.array_equal(ra[f], rb[f]) for f in ra)


def _host_policy(dec, seed, base, step):
    from oracle.cim_oracle import policy_random

    B = dec.shape[0]
    acts = np.zeros((B, 2, 4), np.int32)
    for i in range(B):
        acts[i, 0] = policy_random(dec[i], seed, i + base, step)
    return acts


@pytest.mark.parametrize("topology,ticks,B", [("toy.4p_ssdd_l0.0", 150, 64), ("toy.4p_ssdd_l0.8", 90, 40),
                                               ("global_trade.22p_l0.8", 30, 12)])
def test_session_matches_launch_per_step(topology, ticks, B, monkeypatch):
    """maro_cim_step through the resident session == through one launch per call: decisions, metrics, frames, rings;
    with action lists, None actions, subset (dict) stepping and interleaved inspection calls (which end the session)."""
    topo = _topo(topology, ticks)
    monkeypatch.setenv("MARO_B200_SESSION", "0")
    ref = _batch(topo, B, max_actions=2)
    monkeypatch.setenv("MARO_B200_SESSION", "1")
    ses = _batch(topo, B, max_actions=2)
    rng = np.random.default_rng(3)
    d0, m0 = (x.copy() for x in ref.step(None))
    d1, m1 = (x.copy() for x in ses.step(None))
    step = 0
    while True:
        assert np.array_equal(d0, d1) and np.array_equal(m0, m1), step
        if step % 4 != 2 and (d0[:, 6] == 2).all():  # (a step without a mask: every row is a real answer)
            break
        acts = _host_policy(d0, 4, 0, step)
        nact = np.ones(B, np.int32)
        if step % 5 == 2:  # action lists: LOAD part then a DISCHARGE of 0; every 5th step a few None actions
            acts[:, 1] = acts[:, 0]
            acts[:, 1, 2] = 0
            acts[:, 1, 3] = 1
            nact[:] = 2
            nact[::7] = 0
        active = None
        if step % 4 == 1:
            active = (rng.random(B) < 0.7).astype(np.uint8)
        d0, m0 = (x.copy() for x in ref.step(acts, nact, active))
        d1, m1 = (x.copy() for x in ses.step(acts, nact, active))
        if step % 9 == 4:  # inspection in the middle of a session: the kernel writes back, the next step relaunches
            assert np.array_equal(ref.ticks(), ses.ticks())
            assert np.array_equal(ref.read_frame(B - 1), ses.read_frame(B - 1))
        step += 1
        assert step < 100000
    assert np.array_equal(ref.counters(), ses.counters())
    for i in (0, B - 1):
        assert np.array_equal(ref.read_frame(i), ses.read_frame(i))
        assert np.array_equal(ref.snapshot_frames(i), ses.snapshot_frames(i))
        f = int(ref.snapshot_frames(i)[-1])
        assert np.array_equal(ref.snapshot_row(f, i), ses.snapshot_row(f, i))
    # a finished env keeps answering FINISHED through the session as well
    assert (ses.step(None)[0][:, 6] == 2).all()
    ref.close()
    ses.close()


def test_session_survives_idle_timeout_and_reset(monkeypatch):
    """The resident kernel leaves after MARO_B200_IDLE_US without a command; the next call relaunches it.  Reset in the
    middle of a session; bad actions are reported per replica."""
    from oracle.cim_oracle import CimOracle, policy_random

    monkeypatch.setenv("MARO_B200_IDLE_US", "50")
    topo = _topo("toy.4p_ssdd_l0.0", 80)
    B = 32
    env = _batch(topo, B)
    pa, pn, pact, pd, pm = env.pinned()
    for episode in range(2):
        o = CimOracle(topo)
        env.step_pinned(use_actions=False)
        st, d, m = o.step(None)
        step = 0
        while st == 0:
            assert d[:7].tolist() == pd[3, :7].tolist() and m.tolist() == pm[3].tolist(), (episode, step)
            for i in range(B):
                pa[i, 0] = policy_random(pd[i], 9, i, step)
            a3 = pa[3].copy()
            if step in (5, 6, 17):
                time.sleep(0.003)  # far beyond the idle limit: the kernel has left
            env.step_pinned()
            st, d, m = o.step(a3)
            step += 1
        assert st == 1 and pd[3, 6] == 1 and m.tolist() == pm[3].tolist()
        assert np.array_equal(env.read_frame(3), o.frame())
        env.reset()
    # bad action on one replica only
    env.step_pinned(use_actions=False)
    pa[:, 0] = 0
    pa[:, 0, 0] = pd[:, 2]
    pa[:, 0, 1] = pd[:, 1]
    pa[7, 0, 2] = 10 ** 7
    env.step_pinned()
    assert pd[7, 6] == -1 and (np.delete(pd[:, 6], 7) == 0).all()
    pa[:, 0, 2] = 0
    pn[:] = 1
    pn[5] = 2  # more actions than the handle's rows hold
    env.step_pinned(use_n_actions=True)
    assert pd[5, 6] == -1 and pd[7, 6] == 2
    env.close()


def test_too_many_actions_is_a_bad_action_without_session(monkeypatch):
    monkeypatch.setenv("MARO_B200_SESSION", "0")
    topo = _topo("toy.4p_ssdd_l0.0", 40)
    env = _batch(topo, 4)
    dec, _ = env.step(None)
    acts = np.zeros((4, 1, 4), np.int32)
    acts[:, 0, 0] = dec[:, 2]
    acts[:, 0, 1] = dec[:, 1]
    dec, _ = env.step(acts, np.asarray([1, 2, 1, 0], np.int32))
    assert dec[:, 6].tolist() == [0, -1, 0, 0]
    env.close()


def test_async_submit_wait_pipelines_sub_batches():
    """maro_cim_submit_pinned / maro_cim_wait_pinned: two halves of the batch advance independently (one is always a
    step ahead of the other); both follow the oracle."""
    from oracle.cim_oracle import CimOracle, policy_random

    topo = _topo("toy.4p_ssdd_l0.0", 70)
    B = 64
    env = _batch(topo, B)
    g = env.pinned_granularity()
    assert g > 0 and B % g == 0
    half = (B // g // 2) * g
    ranges = [(0, half), (half, B - half)]
    pa, pn, pact, pd, pm = env.pinned()
    probes = [1, half + 2]
    oracles = [CimOracle(topo) for _ in probes]
    outs = [o.step(None) for o in oracles]
    steps = [0, 0]
    for f, c in ranges:
        env.submit_pinned(f, c, use_actions=False)
    env.wait_pinned(*ranges[0])
    live = [True, True]
    while any(live):
        for k, (f, c) in enumerate(ranges):
            if not live[k]:
                continue
            env.wait_pinned(f, c)
            st, d, m = outs[k]
            i = probes[k]
            assert d[:7].tolist() == pd[i, :7].tolist() and m.tolist() == pm[i].tolist(), (k, steps[k])
            if st != 0:
                assert pd[i, 6] == 1
                live[k] = False
                continue
            for r in range(f, f + c):
                pa[r, 0] = policy_random(pd[r], 3, r, steps[k])
            outs[k] = oracles[k].step(pa[i].copy())
            env.submit_pinned(f, c)
            steps[k] += 1
    for k, i in enumerate(probes):
        assert np.array_equal(env.read_frame(i), oracles[k].frame())
    env.close()


def test_host_threads_drive_disjoint_sub_batches_with_in_session_resets():
    """maro_cim_submit_pinned / _wait_pinned / maro_cim_reset from several host threads on disjoint replica ranges (the
    contract of include/maro_b200.h): every thread runs its own wait -> agent -> submit pipeline over two sub-batches for
    two and a half episodes; Env.reset of a finished sub-batch rides on the next command row (the resident kernel resets the
    replica in shared memory, no write-back / relaunch).  Every replica's decisions and metrics are checked against oracles."""
    import threading

    from oracle.cim_oracle import CimOracle, policy_random

    topo = _topo("toy.4p_ssdd_l0.8", 60)  # noisy orders / buffers: the reset also has to restore the MT19937 streams
    B = 64
    env = _batch(topo, B)
    g = env.pinned_granularity()
    assert g > 0 and B % (4 * g) == 0
    q = B // 4
    ranges = [(k * q, q) for k in range(4)]
    pa, pn, pact, pd, pm = env.pinned()
    oracles = [CimOracle(topo) for _ in range(B)]
    errors = []

    def drive(my_ranges):
        try:
            outs = {}
            steps = {}
            episodes = {f: 0 for f, _ in my_ranges}
            for f, c in my_ranges:
                for r in range(f, f + c):
                    oracles[r].reset()
                    outs[r] = oracles[r].step(None)
                steps[f] = 0
                env.submit_pinned(f, c, use_actions=False)
            budget = 400
            live = {f: True for f, _ in my_ranges}
            while any(live.values()) and budget > 0:
                for f, c in my_ranges:
                    if not live[f]:
                        continue
                    budget -= 1
                    env.wait_pinned(f, c)
                    for r in range(f, f + c):
                        st, d, m = outs[r]
                        assert d[:7].tolist() == pd[r, :7].tolist() and m.tolist() == pm[r].tolist(), (r, steps[f], d, pd[r])
                    if all(outs[r][0] != 0 for r in range(f, f + c)):  # the sub-batch finished its episode
                        episodes[f] += 1
                        if episodes[f] == 3:
                            live[f] = False
                            continue
                        mask = np.zeros(B, np.uint8)
                        mask[f:f + c] = 1
                        env.reset(mask)
                        for r in range(f, f + c):
                            oracles[r].reset()
                            outs[r] = oracles[r].step(None)
                        steps[f] = 0
                        env.submit_pinned(f, c, use_actions=False)
                        continue
                    for r in range(f, f + c):
                        pa[r, 0] = policy_random(pd[r], 9, r, steps[f])
                        if outs[r][0] == 0:
                            outs[r] = oracles[r].step(pa[r].copy())
                        elif outs[r][0] == 1:
                            outs[r] = oracles[r].step(None)  # DONE -> FINISHED row
                    env.submit_pinned(f, c)
                    steps[f] += 1
            assert not any(live.values())
        except BaseException as ex:  # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=drive, args=(ranges[:2],)), threading.Thread(target=drive, args=(ranges[2:],))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    for r in (0, q, 2 * q + 1, B - 1):
        assert np.array_equal(env.read_frame(r), oracles[r].frame())
    # a reset that never rode on a command row is applied when the session ends
    env.submit_pinned(0, B, use_actions=False)
    env.wait_pinned(0, B)
    env.reset()
    assert (env.ticks() == 0).all() and env.read_frame(5)[0] == 0
    env.close()
