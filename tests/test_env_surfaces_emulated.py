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


def test_vm_float_queries_are_lifted_to_the_reference_float64_values():
    """VmBatch.query: cpu_utilization / energy_consumption come back as the float64 values the reference's static backend
    holds (k / 100 and the energy model at k), not as the float32 words of the ring; unknown frames stay zero."""
    import numpy as np

    from vm_helpers import VM_CASES, load_vm_golden, vm_topology

    spec, gold = VM_CASES["synth_160_bestfit"], load_vm_golden("synth_160_bestfit")
    topo = vm_topology(spec)
    b = EmulVmBatch(topo, 1)
    b.step(None)
    for k in range(40):
        a = np.zeros((1, 1, 4), np.int32)
        a[0, 0] = gold["actions"][k]
        b.step(a, np.ones(1, np.int32))
    frames = b.snapshot_frames(0).tolist()[-3:]
    n = topo.n_pm
    q = b.query("pms", frames + [10 ** 6], np.arange(n), ["cpu_utilization", "pm_type", "energy_consumption"])[0].reshape(4, n, 3)
    assert (q[3] == 0).all()                                   # unknown frame -> zeros
    u = q[:3, :, 0]
    assert np.array_equal(u, np.rint(u * 100) / 100.0) and (u > 0).any()   # exact multiples of 0.01 in float64
    for f in range(3):
        for p in range(n):
            calib, busy, idle = topo.pmtype_power[int(q[f, p, 1])]
            x = min(1, u[f, p] / 100)
            assert q[f, p, 2] == ((idle + (busy - idle) * (2 * x - pow(x, calib))) / topo.ticks_per_hour) / 1000
    only_e = b.query("pms", frames, np.arange(n), ["energy_consumption"])[0].reshape(3, n)
    assert np.array_equal(only_e, q[:3, :, 2])                 # energy alone: utilisation fetched behind the scenes


def test_reference_examples_run_unchanged_on_the_shim_emulated():
    """maro import shim: hello_world/cim/hello.py, vector_env/hello.py and the maro.rl CIM sampler + DQN train step, unmodified"""
    import test_gpu_shim as shim_cases

    if not shim_cases.HAVE_REF:
        pytest.skip("oracle/_ref/examples not built")
    shim_cases.test_hello_world_cim_unchanged(emulate=True)
    shim_cases.test_vector_env_hello_unchanged(emulate=True)
    shim_cases.test_rl_toolkit_sampler_and_train_step_unchanged(emulate=True)


def test_vector_env_reset_reseeds_like_the_reference_processes_emulated():
    surfaces.test_vector_env_reset_reseeds_like_the_reference_processes()


def test_dynamic_backend_query_layout_matches_the_reference_emulated(monkeypatch):
    surfaces.test_dynamic_backend_query_layout_matches_the_reference(monkeypatch)


def test_frame_dump_is_byte_identical_to_the_reference_emulated(tmp_path):
    surfaces.test_frame_dump_is_byte_identical_to_the_reference(tmp_path)
