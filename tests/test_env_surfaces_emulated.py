"""CPU: the Python façades (maro_b200.simulator.Env, maro_b200.vector_env.VectorEnv, SnapshotList) driven exactly like
the GPU suite drives them (tests/test_gpu_env_surfaces.py — reference-style tests), with the CUDA batch replaced by the
host emulation of the same device code (tests/emul_batch.py).  Covers the host logic of the drop-in surfaces: action
encoding, DecisionEvent / metrics decoding, status handling, snapshot_list slicing, reset / set_seed / keep_seed
bookkeeping, list / dict / broadcast stepping."""
import pytest

import test_gpu_bike as bike_surfaces
import test_gpu_env_surfaces as surfaces
import test_gpu_vm as vm_surfaces
from emul_batch import EmulBikeBatch, EmulCimBatch, EmulVmBatch


@pytest.fixture(autouse=True)
def emulated_batch(monkeypatch):
    import maro_b200.simulator.env as env_mod
    import maro_b200.vector_env.vector_env as venv_mod

    for mod in (env_mod, venv_mod):
        monkeypatch.setattr(mod, "CimBatch", EmulCimBatch)
        monkeypatch.setattr(mod, "BikeBatch", EmulBikeBatch)
        monkeypatch.setattr(mod, "VmBatch", EmulVmBatch)
    yield


def test_env_hello_world_null_policy_emulated():
    surfaces.test_env_hello_world_null_policy()


def test_env_actions_pickle_reset_and_seed_emulated():
    surfaces.test_env_actions_pickle_reset_and_seed()


def test_vector_env_like_reference_test_emulated():
    surfaces.test_vector_env_like_reference_test()


def test_env_reset_new_seed_matches_reference_emulated():
    surfaces.test_env_reset_new_seed_matches_reference()


def test_bike_env_surface_like_reference_tests_emulated(tmp_path):
    bike_surfaces.test_bike_env_surface_like_reference_tests(tmp_path)


def test_vm_env_surface_replays_reference_trace_emulated():
    vm_surfaces.test_vm_env_surface_replays_reference_trace()
