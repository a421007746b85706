"""CPU: the numpy restatement of the reference's RL shaping, evaluated on the C oracle's snapshots, against vectors
recorded from the unmodified reference (tests/golden/gen_cim_rl_golden.py).  Pins the checker used by the GPU test."""
import numpy as np
import pytest

from maro_b200.scenarios.cim.topology import build_topology
from oracle.cim_oracle import CimOracle
from rl_helpers import RL_CASES, SnapshotView, action_numpy, load_rl_golden, reward_numpy, state_numpy


class _EmulAsOracle:
    """the device logic under the host emulator behind the oracle's step / snapshot interface"""

    def __init__(self, topo, max_snapshots=None):
        from emul import EmulEnv

        self._e = EmulEnv(topo, 1, 0, 1, max_snapshots)
        self.snapshot = self._e.snapshot

    def step(self, actions):
        return self._e.step1(actions)


@pytest.mark.parametrize("backend", ["oracle", "emulated-kernel"])
@pytest.mark.parametrize("name", sorted(RL_CASES))
def test_rl_shaping_restatement_matches_reference(name, backend):
    spec, gold = RL_CASES[name], load_rl_golden(name)
    topo = build_topology(spec["topology"], spec["durations"])
    o = (CimOracle if backend == "oracle" else _EmulAsOracle)(topo, max_snapshots=spec.get("max_snapshots"))
    view = SnapshotView(o.snapshot, topo)
    st, dec, _ = o.step(None)
    k = 0
    while st == 0:
        assert [int(dec[0]), int(dec[1]), int(dec[2])] == gold["steps"][k].tolist()
        s = state_numpy(view, int(dec[0]), int(dec[1]), int(dec[2]))
        assert s.shape == gold["states"][k].shape and np.array_equal(s, gold["states"][k]), k
        assert action_numpy(view, dec, int(gold["model_actions"][k])) == gold["actions"][k].tolist(), k
        st, dec, _ = o.step(gold["actions"][k].reshape(1, 4))
        k += 1
    assert k == len(gold["steps"])
    r = np.asarray([reward_numpy(view, int(p), int(t)) for t, p, _ in gold["steps"]], np.float32)
    assert np.allclose(r, gold["rewards"], rtol=1e-6, atol=1e-3)
