"""Batched, device-resident state / reward shaping for the CIM scenario.

The reference's RL example shapes one decision at a time from Python: a ``snapshot_list`` query for the look-back
window of the acting port and the vessel's next stops, and later a 99-tick window query + a decayed dot product for
the reward (``examples/cim/rl/env_sampler.py:15-36, 66-80``, constants in ``examples/cim/rl/config.py``).  Here both run
as one kernel launch over all replicas on the snapshot ring in HBM and return torch CUDA tensors — nothing crosses PCIe.
"""
from typing import Sequence

import numpy as np

from .batch import CimBatch

# examples/cim/rl/config.py:10-36
PORT_ATTRIBUTES = ("empty", "full", "on_shipper", "on_consignee", "booking", "shortage", "fulfillment")
VESSEL_ATTRIBUTES = ("empty", "full", "remaining_space")
ACTION_SPACE = tuple((i - 10) / 10 for i in range(21))


class CimShaper:
    def __init__(self, batch: CimBatch, look_back: int = 7, port_attributes: Sequence[str] = PORT_ATTRIBUTES,
                 vessel_attributes: Sequence[str] = VESSEL_ATTRIBUTES, time_window: int = 99, time_decay: float = 0.97,
                 fulfillment_factor: float = 1.0, shortage_factor: float = 1.0, action_space: Sequence[float] = ACTION_SPACE,
                 finite_vessel_space: bool = True, has_early_discharge: bool = True):
        import torch

        self._torch = torch
        self.batch = batch
        self.look_back = int(look_back)
        self._pa = [batch.attr_id("ports", a) for a in port_attributes]
        self._va = [batch.attr_id("vessels", a) for a in vessel_attributes]
        self.state_dim = batch.rl_state_dim(self.look_back, len(self._pa), len(self._va))
        self.time_window = int(time_window)
        self.fulfillment_factor, self.shortage_factor = float(fulfillment_factor), float(shortage_factor)
        dev = torch.device("cuda", batch.device)
        # the shaping kernels and the caller's torch ops touch the same tensors: run the library on torch's current stream
        # of the batch's device (its own non-blocking stream would race with them)
        batch.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        # decay_list = [time_decay ** i ...] evaluated on the host exactly like the example (env_sampler.py:75)
        self._decay = torch.tensor(np.asarray([time_decay ** i for i in range(self.time_window)], np.float64), device=dev)
        B = batch.n_replicas
        self._state = torch.zeros((B, self.state_dim), dtype=torch.float64, device=dev)
        self._reward = torch.zeros(B, dtype=torch.float32, device=dev)
        self._space = torch.tensor(np.asarray(action_space, np.float64), device=dev)
        self.finite_vessel_space, self.has_early_discharge = bool(finite_vessel_space), bool(has_early_discharge)
        self._actions = torch.zeros((B, batch.max_actions, 4), dtype=torch.int32, device=dev)

    def states(self, decisions, out=None):
        """decisions: int32 CUDA tensor [B][8] (MARO_DEC_* rows of the last step) -> float64 CUDA tensor [B][state_dim]; with
        ``out`` (contiguous float32 CUDA tensor [B][state_dim]) the float32-rounded state is written there instead."""
        torch = self._torch
        assert decisions.is_cuda and decisions.dtype == torch.int32 and decisions.is_contiguous()
        if out is not None:
            assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == tuple(self._state.shape)
            self.batch.rl_state_device(decisions.data_ptr(), self.look_back, self._pa, self._va, out.data_ptr(), f32=True)
            return out
        self.batch.rl_state_device(decisions.data_ptr(), self.look_back, self._pa, self._va, self._state.data_ptr())
        return self._state

    def rewards(self, ticks, ports):
        """ticks, ports: int32 CUDA tensors [B] (tick at which replica i acted, acting port) -> float32 CUDA tensor [B]."""
        for t in (ticks, ports):
            assert t.is_cuda and t.dtype == self._torch.int32 and t.is_contiguous()
        self.batch.rl_reward_device(ticks.data_ptr(), ports.data_ptr(), self._decay.data_ptr(), self.time_window,
                                    self.fulfillment_factor, self.shortage_factor, self._reward.data_ptr())
        return self._reward

    def rewards_batch(self, ticks, ports):
        """ticks, ports: int32 CUDA tensors [T][B] (a negative tick = no decision there -> reward 0) -> float32 [T][B], one launch"""
        torch = self._torch
        for t in (ticks, ports):
            assert t.is_cuda and t.dtype == torch.int32 and t.is_contiguous() and t.dim() == 2
        out = torch.empty(ticks.shape, dtype=torch.float32, device=ticks.device)
        self.batch.rl_reward_batch_device(ticks.data_ptr(), ports.data_ptr(), ticks.shape[0], self._decay.data_ptr(), self.time_window,
                                          self.fulfillment_factor, self.shortage_factor, out.data_ptr())
        return out

    def env_actions(self, decisions, model_actions, record=None, metrics=None, final_metrics=None):
        """decisions int32 [B][8], model_actions int32 or int64 [B] (indices into the action space) -> int32 CUDA tensor
        [B][max_actions][4], the `actions` argument of ``CimBatch.step_device`` (env_sampler.py:38-64).  Optional bookkeeping of a
        collection loop, done by the same launch: ``record`` int32 [B] receives the indices; ``metrics`` int64 [B][3] (the
        previous step's) are folded into the running maximum ``final_metrics`` int64 [B][3]."""
        torch = self._torch
        assert model_actions.is_cuda and model_actions.is_contiguous() and model_actions.dtype in (torch.int32, torch.int64)
        if record is None and metrics is None and model_actions.dtype == torch.int32:
            self.batch.rl_action_device(decisions.data_ptr(), model_actions.data_ptr(), self._space.data_ptr(), self._space.numel(),
                                        self.finite_vessel_space, self.has_early_discharge, self._actions.data_ptr())
            return self._actions
        for t, dt in ((record, torch.int32), (metrics, torch.int64), (final_metrics, torch.int64)):
            assert t is None or (t.is_cuda and t.is_contiguous() and t.dtype == dt)
        assert (metrics is None) == (final_metrics is None)
        self.batch.rl_action_ex_device(decisions.data_ptr(), model_actions.data_ptr(), model_actions.dtype == torch.int64,
                                       0 if record is None else record.data_ptr(), 0 if metrics is None else metrics.data_ptr(),
                                       0 if final_metrics is None else final_metrics.data_ptr(), self._space.data_ptr(),
                                       self._space.numel(), self.finite_vessel_space, self.has_early_discharge, self._actions.data_ptr())
        return self._actions
