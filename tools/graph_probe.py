"""GPU probe: which part of the BatchedCimEnvSampler loop body can be captured in a CUDA graph."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from maro_b200.batch import CimBatch  # noqa: E402
from maro_b200.rl_shaping import CimShaper  # noqa: E402
from maro_b200.scenarios.cim.topology import build_topology  # noqa: E402

B = 256
env = CimBatch(build_topology("toy.4p_ssdd_l0.0", 200), B)
sh = CimShaper(env)
dec = torch.zeros((B, 8), dtype=torch.int32, device="cuda")
met = torch.zeros((B, 3), dtype=torch.int64, device="cuda")
env.step_device(dec.data_ptr(), met.data_ptr())
for _ in range(5):
    env.step_device(dec.data_ptr(), met.data_ptr())
m = torch.zeros(B, dtype=torch.int32, device="cuda")
mlp = torch.nn.Sequential(torch.nn.Linear(sh.state_dim, 64), torch.nn.ReLU(), torch.nn.Linear(64, 21)).cuda()
torch.cuda.synchronize()
pieces = {
    "states": lambda: sh.states(dec),
    "states+cast": lambda: sh.states(dec).to(torch.float32),
    "actions": lambda: sh.env_actions(dec, m),
    "step": lambda: env.step_device(dec.data_ptr(), met.data_ptr(), sh._actions.data_ptr()),
    "policy": lambda: mlp(sh._state.to(torch.float32) * 1e-4).argmax(1).to(torch.int32),
    "rewards": lambda: sh.rewards(dec[:, 0].contiguous(), dec[:, 1].contiguous()),
}
for name, fn in pieces.items():
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    try:
        with torch.cuda.stream(side), torch.no_grad():
            env.set_stream(side.cuda_stream)
            fn()  # warm-up on the side stream
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                fn()
        g.replay()
        torch.cuda.synchronize()
        print(name, "captured ok", flush=True)
    except Exception as ex:
        print(name, "FAILED:", repr(ex)[:300], flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as ex2:
            print("  sync after failure:", repr(ex2)[:200])
    finally:
        env.set_stream(torch.cuda.current_stream().cuda_stream)
