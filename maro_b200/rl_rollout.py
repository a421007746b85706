"""Device-resident rollout loop for the CIM scenario.

The reference collects experience one environment and one Python object at a time
(``maro/rl/rollout/env_sampler.py:391-424``: ``get_state`` -> policy -> ``translate_to_env_action`` -> ``env.step`` ->
``get_reward`` once ``reward_eval_delay`` ticks have passed).  ``CimDeviceRollout`` runs the same loop for all replicas
of a ``CimBatch`` at once with every tensor in HBM: shaping kernels (``rl_shaping.CimShaper``), the user's policy (any
callable on CUDA tensors, e.g. a torch module), the step kernel — and the rewards after the episode, when all the frames
they look at exist.  No per-replica Python objects, no PCIe traffic inside the loop.
"""
from typing import Callable, Optional

from .batch import CimBatch
from .rl_shaping import CimShaper


class CimDeviceRollout:
    def __init__(self, batch: CimBatch, policy: Callable, shaper: Optional[CimShaper] = None, store_states: bool = True):
        """policy(states float32 [B][state_dim]) -> integer tensor [B] of indices into the shaper's action space."""
        import torch

        self._torch = torch
        self.batch, self.policy = batch, policy
        self.shaper = shaper or CimShaper(batch)
        self.store_states = store_states
        B = batch.n_replicas
        dev = torch.device("cuda", batch.device)
        batch.set_stream(torch.cuda.current_stream(dev).cuda_stream)  # step kernels and torch ops share one stream
        self._dec = torch.zeros((B, 8), dtype=torch.int32, device=dev)
        self._met = torch.zeros((B, 3), dtype=torch.int64, device=dev)

    def run_episode(self, max_steps: int = 1 << 30) -> dict:
        """Reset, run until every replica is done (or max_steps).  Returns CUDA tensors, time-major:
        ``valid`` bool [T][B] (replica had a decision at step t), ``ticks`` / ``ports`` / ``vessels`` / ``model_actions``
        int32 [T][B], ``states`` float32 [T][B][D] (if store_states), ``rewards`` float32 [T][B] (0 where not valid),
        ``metrics`` int64 [B][3] (order_requirements, container_shortage, operation_number)."""
        torch, env, sh = self._torch, self.batch, self.shaper
        dec, met = self._dec, self._met
        env.reset()
        env.step_device(dec.data_ptr(), met.data_ptr())
        valid, ticks, ports, vessels, models, states = [], [], [], [], [], []
        for _ in range(max_steps):
            live = dec[:, 6] == 0
            if not bool(live.any()):
                break
            s = sh.states(dec).to(torch.float32)
            with torch.no_grad():
                m = self.policy(s).to(torch.int32).contiguous()
            actions = sh.env_actions(dec, m)
            valid.append(live)
            ticks.append(dec[:, 0].clone()); ports.append(dec[:, 1].clone()); vessels.append(dec[:, 2].clone())
            models.append(m)
            if self.store_states:
                states.append(s)
            env.step_device(dec.data_ptr(), met.data_ptr(), actions.data_ptr())
        out = {"valid": torch.stack(valid), "ticks": torch.stack(ticks), "ports": torch.stack(ports),
               "vessels": torch.stack(vessels), "model_actions": torch.stack(models), "metrics": met.clone()}
        if self.store_states:
            out["states"] = torch.stack(states)
        # rewards look `time_window` frames past the action tick: evaluated once the episode's frames all exist
        rewards = torch.zeros(out["ticks"].shape, dtype=torch.float32, device=dec.device)
        for t in range(out["ticks"].shape[0]):
            r = sh.rewards(out["ticks"][t].contiguous(), out["ports"][t].contiguous())
            rewards[t] = torch.where(out["valid"][t], r, torch.zeros_like(r))
        out["rewards"] = rewards
        return out
