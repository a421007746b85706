"""GPU (-m gpu): device-state checkpoint (maro_*_save / maro_*_load; Env.dump(path) / Env.restore(path)) — the "dump
environment for restore" the reference declares and leaves unimplemented (maro/simulator/core.py:135-141).  An episode
interrupted by save -> (other work) -> load must continue bit for bit: decisions, metrics, live frames, snapshot ring."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(env, policy, dec, met, n_steps, step0):
    rows = []
    for k in range(n_steps):
        if not (dec[:, 6] == 0).any():
            break
        acts = np.zeros((env.n_replicas, 1, 4), np.int32)
        for i in range(env.n_replicas):
            acts[i, 0] = policy(dec[i], 21, i, step0 + k)
        dec, met = env.step(acts)
        rows.append((dec.copy(), met.copy()))
    return rows, dec, met


@pytest.mark.parametrize("topology,ticks,with_snapshots", [("toy.4p_ssdd_l0.8", 80, True), ("global_trade.22p_l0.8", 30, False)])
def test_cim_checkpoint_continues_bit_for_bit(tmp_path, topology, ticks, with_snapshots):
    from maro_b200.batch import CimBatch
    from maro_b200.scenarios.cim.topology import build_topology
    from oracle.cim_oracle import policy_random

    topo = build_topology(topology, ticks)
    B = 12
    env = CimBatch(topo, B, max_snapshots=16)
    dec, met = env.step(None)
    dec, met = dec.copy(), met.copy()
    _, dec, met = _run(env, policy_random, dec, met, 17, 0)
    dec0, met0 = dec.copy(), met.copy()
    path = str(tmp_path / "cim.ckpt")
    env.save(path, with_snapshots)
    rest_a, dec_a, met_a = _run(env, policy_random, dec0.copy(), met0.copy(), 10 ** 6, 17)
    frames_a = [env.read_frame(i) for i in range(B)]
    ring_a = {int(f): env.snapshot_row(int(f), 3) for f in env.snapshot_frames(3)}
    # same handle: rewind
    env.load(path)
    rest_b, dec_b, met_b = _run(env, policy_random, dec0.copy(), met0.copy(), 10 ** 6, 17)
    assert len(rest_a) == len(rest_b) and len(rest_a) > 5
    for (d1, m1), (d2, m2) in zip(rest_a, rest_b):
        assert np.array_equal(d1, d2) and np.array_equal(m1, m2)
    for i in range(B):
        assert np.array_equal(env.read_frame(i), frames_a[i])
    if with_snapshots:
        assert {int(f) for f in env.snapshot_frames(3)} == set(ring_a)
        for f, row in ring_a.items():
            assert np.array_equal(env.snapshot_row(f, 3), row)
    # a fresh handle of the same configuration picks the episode up as well
    env2 = CimBatch(topo, B, max_snapshots=16)
    env2.load(path)
    rest_c, _, _ = _run(env2, policy_random, dec0.copy(), met0.copy(), 10 ** 6, 17)
    for (d1, m1), (d2, m2) in zip(rest_a, rest_c):
        assert np.array_equal(d1, d2) and np.array_equal(m1, m2)
    # a handle of another shape refuses the file
    env3 = CimBatch(topo, B + 1, max_snapshots=16)
    with pytest.raises(RuntimeError):
        env3.load(path)
    for e in (env, env2, env3):
        e.close()


def test_env_dump_and_restore(tmp_path):
    from maro_b200.simulator import Env

    env = Env("cim", "toy.4p_ssdd_l0.0", durations=60)
    m, ev, done = env.step(None)
    for _ in range(9):
        m, ev, done = env.step(None)
    path = str(tmp_path / "env.ckpt")
    env.dump(path)
    tick = env.tick
    tail_a = []
    while not done:
        m, ev, done = env.step(None)
        tail_a.append((ev.tick, ev.port_idx, ev.vessel_idx, ev.action_scope.load, dict(m)) if ev else dict(m))
    env.restore(path)
    assert env.tick == tick
    tail_b, done = [], False
    while not done:
        m, ev, done = env.step(None)
        tail_b.append((ev.tick, ev.port_idx, ev.vessel_idx, ev.action_scope.load, dict(m)) if ev else dict(m))
    assert tail_a == tail_b and len(tail_a) > 10
    env.dump()  # no path: the reference's no-op
    env.close()


def test_bike_and_vm_checkpoints(tmp_path):
    from bike_helpers import BIKE_CASES, bike_topology, greedy_py
    from maro_b200.batch import BikeBatch, VmBatch
    from vm_helpers import VM_CASES, vm_topology

    spec = BIKE_CASES["toy_600_greedy_res1"]
    env = BikeBatch(bike_topology(spec), 4, spec["snapshot_resolution"], spec.get("max_snapshots"))

    def bike_run(n, dec):
        out = []
        for _ in range(n):
            if not (dec[:, 6] == 0).any():
                break
            a = np.zeros((4, 1, 4), np.int32)
            for i in range(4):
                a[i, 0] = greedy_py(dec[i])
            dec, met = env.step(a)
            out.append((dec.copy(), met.copy()))
        return out, dec

    dec, met = env.step(None)
    _, dec = bike_run(25, dec.copy())
    d0 = dec.copy()
    p = str(tmp_path / "bike.ckpt")
    env.save(p)
    a, _ = bike_run(10 ** 6, d0.copy())
    env.load(p)
    b, _ = bike_run(10 ** 6, d0.copy())
    assert len(a) == len(b) and len(a) > 5
    for (d1, m1), (d2, m2) in zip(a, b):
        assert np.array_equal(d1, d2) and np.array_equal(m1, m2)
    env.close()

    vspec = VM_CASES["synth_160_bestfit"]
    topo = vm_topology(vspec)
    venv = VmBatch(topo, 3, 1, 8)

    def vm_run(n, dec):
        out = []
        for _ in range(n):
            if not (dec[:, 6] == 0).any():
                break
            a = np.zeros((3, 1, 4), np.int32)
            for i in range(3):  # first valid PM (decision row: vm id at [1], n_valid at [10], valid PM ids from [12])
                a[i, 0] = [dec[i, 1], 0, dec[i, 12], 0] if dec[i, 6] == 0 and dec[i, 10] > 0 else [-1, -1, 0, 0]
            dec, met = venv.step(a)
            out.append((dec.copy(), met.copy()))
        return out, dec

    dec, met = venv.step(None)
    _, dec = vm_run(30, dec.copy())
    d0 = dec.copy()
    p = str(tmp_path / "vm.ckpt")
    venv.save(p)
    a, _ = vm_run(10 ** 6, d0.copy())
    venv.load(p)
    b, _ = vm_run(10 ** 6, d0.copy())
    assert len(a) == len(b) and len(a) > 5
    for (d1, m1), (d2, m2) in zip(a, b):
        assert np.array_equal(d1, d2) and np.array_equal(m1, m2)
    venv.close()
