"""Randomised differential check of the vm_scheduling path: unmodified reference (oracle/_ref, one process per case)
vs the C oracle vs the device code under the host emulator, on random configurations over the synthetic trace.
Build-container tool (needs oracle/_ref); the committed golden traces are the fixed subset the test-suite uses.

    python tools/fuzz_vm_parity.py [n_cases] [first_seed]
"""
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def random_spec(seed):
    import gen_vm_golden as g

    rng = np.random.default_rng(seed)
    n_types = int(rng.integers(1, 3))
    pms = [(int(rng.choice([16, 24, 32])), int(rng.choice([64, 96, 128])), int(rng.integers(150, 200)), int(rng.integers(60, 130)))
           for _ in range(n_types)]
    over = dict(BUFFER_TIME_BUDGET=int(rng.integers(0, 7)), DELAY_DURATION=int(rng.integers(1, 4)),
                KILL_ALL_VMS_IF_OVERLOAD=bool(rng.integers(0, 2)),
                MAX_CPU_OVERSUBSCRIPTION_RATE=float(rng.choice([1.0, 1.15, 1.5, 2.5])),
                MAX_MEM_OVERSUBSCRIPTION_RATE=float(rng.choice([1.0, 1.2])),
                MAX_UTILIZATION_RATE=float(rng.choice([0.8, 1.0, 1.5, 3.0])), TICKS_PER_HOUR=int(rng.choice([12, 6])))
    conf = g.config_multi("vm_synth", **over) if rng.random() < 0.3 else g.config("vm_synth", pms, int(rng.integers(1, 4)),
                                                                                   int(rng.integers(1, 4)), **over)
    start = int(rng.choice([0, 0, 30, 85]))
    return dict(conf=conf, start_tick=start, durations=int(rng.integers(40, 160 - start // 2)),
                agent=str(rng.choice(["first", "best", "mixed"])), snapshot_resolution=int(rng.choice([1, 1, 2, 5])),
                max_snapshots=(None if rng.random() < 0.5 else int(rng.integers(3, 40))))


def main():
    import gen_vm_golden as g
    from emul import VmEmulEnv
    from oracle.vm_oracle import VmOracle
    from vm_helpers import assert_metrics_close, assert_vm_snapshots_equal, drive_vm, vm_topology

    n, first = (int(sys.argv[1]) if len(sys.argv) > 1 else 20), (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    out = tempfile.mkdtemp()
    ctx = mp.get_context("spawn")
    bad = 0
    for seed in range(first, first + n):
        spec = random_spec(seed)
        name = f"fuzz{seed}"
        p = ctx.Process(target=g.run_case, args=(name, spec, out))
        p.start(); p.join()
        if p.exitcode != 0:
            print(seed, "reference failed (skipped)")
            continue
        gold = np.load(os.path.join(out, f"vm_{name}.npz"))
        topo = vm_topology(spec)
        try:
            for make in (lambda: VmOracle(topo, spec["snapshot_resolution"], spec["max_snapshots"]),
                         lambda: VmEmulEnv(topo, spec["snapshot_resolution"], spec["max_snapshots"])):
                e = make()
                rows, valid, mets, final, st, dec = drive_vm(lambda a: e.step(a), gold, topo.n_pm)
                assert np.array_equal(rows, gold["steps"]) and np.array_equal(valid, gold["valid"])
                assert_metrics_close(mets, gold["metrics"], "per-step")
                assert_metrics_close(final, gold["final_metrics"], "final")
                assert_vm_snapshots_equal(e.snapshot, gold, topo)
        except AssertionError as ex:
            bad += 1
            print(seed, "MISMATCH", str(ex)[:300], {k: v for k, v in spec.items() if k != "conf"})
            continue
        print(seed, "ok", len(gold["steps"]), "steps", {k: v for k, v in spec.items() if k != "conf"}, flush=True)
    print("mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
