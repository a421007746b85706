"""Batched, device-resident experience collection for the CIM scenario.

The reference collects experience one environment and one Python object at a time
(``AbsEnvSampler.sample``, maro/rl/rollout/env_sampler.py:438-520: ``get_state`` -> policy -> ``translate_to_env_action`` ->
``env.step`` -> ``get_reward`` once ``reward_eval_delay`` ticks have passed -> ``ExpElement`` per transition).
``BatchedCimEnvSampler`` runs the same loop for every replica of a ``CimBatch`` at once with all tensors in HBM:

* ``collect()``: shaping kernels (``rl_shaping.CimShaper``) -> the caller's policy (any callable on CUDA tensors, e.g. a torch
  module) -> action translation -> step kernel, for a fixed number of steps that is known up front (an episode's decision
  count is a property of the stop tables) — no host synchronisation inside the loop, finished replicas answer with
  no-op rows; the loop body is replayed from CUDA graphs in chunks when the policy can be captured; the rewards of the
  whole trajectory come from ONE launch over ``[T, B]``.  Output: columnar, time-major CUDA tensors.
* ``sample()``: the reference sampler's return value — ``{"experiences": [[ExpElement, ...] per env], "info": [...]}`` —
  materialised lazily from the columns for the requested environments: per-agent (= per-port) next states, terminal flags
  and the ``reward_eval_delay`` cut-off follow ``env_sampler.py:402-424, 500-520``.  With ``maro.rl`` importable its own
  ``ExpElement`` class is used, so ``TrainingManager.record_experiences`` takes the lists as they are.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from .batch import CimBatch
from .rl_shaping import CimShaper


@dataclass
class ExpElement:
    """Field-compatible stand-in for ``maro.rl.rollout.ExpElement`` (env_sampler.py:140-215) when maro.rl is not importable."""

    tick: int
    state: np.ndarray
    agent_state_dict: Dict[Any, np.ndarray]
    action_dict: Dict[Any, np.ndarray]
    reward_dict: Dict[Any, float]
    terminal_dict: Dict[Any, bool]
    next_state: Optional[np.ndarray]
    next_agent_state_dict: Dict[Any, np.ndarray]
    truncated: bool

    @property
    def agent_names(self) -> list:
        return sorted(self.agent_state_dict.keys())

    @property
    def num_agents(self) -> int:
        return len(self.agent_state_dict)


def _exp_element_cls():
    try:
        from maro.rl.rollout import ExpElement as RefExpElement  # the real maro.rl, if the user has it (import shim)

        return RefExpElement
    except Exception:
        return ExpElement


def episode_step_bound(topologies, start_tick: int = 0) -> int:
    """Upper bound of ``Env.step`` calls per episode: one per vessel arrival inside the horizon (a decision event each,
    cim/business_engine.py:150-199) + the first call + the final one.  Arrivals are rows of the unrolled stop tables."""
    bound = 0
    for t in topologies:
        n = 0
        for v in range(t.n_vessels):
            stops = t.stops_of(v)
            n += sum(1 for k, (arrival, _, _) in enumerate(stops) if k >= 1 and start_tick <= arrival < t.max_tick)
        bound = max(bound, n)
    return bound + 2


class BatchedCimEnvSampler:
    def __init__(self, batch: CimBatch, policy: Callable, shaper: Optional[CimShaper] = None, reward_eval_delay: Optional[int] = None,
                 store_states: bool = True, graph_chunk: int = 16, use_graph: bool = True, policy_takes_decisions: bool = False):
        """``policy(states float32 [B][state_dim]) -> integer tensor [B]`` (indices into the shaper's action space); with
        ``policy_takes_decisions`` it is called as ``policy(states, decisions int32 [B][8])`` (column 1 = acting port, the
        reference's agent id; column 6 = status, 0 where the replica has a decision) — e.g. to route rows to per-port policies."""
        import torch

        self._torch = torch
        self.batch, self.policy = batch, policy
        self.shaper = shaper or CimShaper(batch)
        self.reward_eval_delay = self.shaper.time_window if reward_eval_delay is None else int(reward_eval_delay)
        self.store_states = store_states
        self.graph_chunk, self.use_graph = max(1, int(graph_chunk)), bool(use_graph)
        self.policy_takes_decisions = bool(policy_takes_decisions)
        self.device = torch.device("cuda", batch.device)
        batch.set_stream(torch.cuda.current_stream(self.device).cuda_stream)  # step kernels and torch ops share one stream
        B, D = batch.n_replicas, self.shaper.state_dim
        self._met = torch.zeros((B, 3), dtype=torch.int64, device=self.device)
        self.n_steps = episode_step_bound(batch.topologies, batch.start_tick)
        K = self.graph_chunk
        # one chunk of recorded columns (static addresses: the loop body may live in a CUDA graph).  Decision rows: step k reads row
        # k and the step kernel writes row k + 1 directly, so recording them costs no launch; row 0 is the chunk's input.
        self._c_dec = torch.zeros((K + 1, B, 8), dtype=torch.int32, device=self.device)
        self._c_act = torch.zeros((K, B), dtype=torch.int32, device=self.device)
        self._c_state = torch.zeros((K if store_states else 1, B, D), dtype=torch.float32, device=self.device)
        self._final_met = torch.zeros((B, 3), dtype=torch.int64, device=self.device)
        self._graph = None
        self._warmed_up = False
        self.graph_error = None
        self.last = None

    # ------------------------------------------------------------------------------------------------ device loop
    def _body(self, k: int):
        """one env-step of every replica: state -> policy -> action -> step; slot k records the decision it answered.
        Launches per step besides the policy's own: state kernel (float32, straight into the record), action kernel (takes the
        policy's int64 / int32 output as it is, records it as int32 and folds the previous step's metrics into the running
        maximum), step kernel (writes the next decision row of the record)."""
        torch, sh, env = self._torch, self.shaper, self.batch
        dec = self._c_dec[k]
        s = sh.states(dec, out=self._c_state[k if self.store_states else 0])
        m = self.policy(s, dec) if self.policy_takes_decisions else self.policy(s)
        if m.dtype not in (torch.int32, torch.int64) or not m.is_contiguous():
            m = m.to(torch.int32).contiguous()
        # the episode's metrics come with the DONE row; a replica stepped past it answers all-zero FINISHED rows.  The three
        # metrics are non-negative running totals, so the value at DONE is the maximum over the episode (the last step's
        # metrics are folded in by collect()).
        actions = sh.env_actions(dec, m, record=self._c_act[k], metrics=self._met, final_metrics=self._final_met)
        env.step_device(self._c_dec[k + 1].data_ptr(), self._met.data_ptr(), actions.data_ptr())

    def _run_chunk(self, n: int):
        torch = self._torch
        # the first full chunk always runs eagerly: libraries the policy uses (cuBLAS handles, workspaces, lazily compiled
        # kernels) must be initialised outside of a stream capture
        if not self.use_graph or n != self.graph_chunk or not self._warmed_up:
            with torch.no_grad():
                for k in range(n):
                    self._body(k)
            self._warmed_up = self._warmed_up or n == self.graph_chunk
            return
        if self._graph is None:
            side = torch.cuda.Stream(self.device)  # capture on a side stream (torch's rule), the batch follows it there
            side.wait_stream(torch.cuda.current_stream(self.device))
            try:
                with torch.cuda.stream(side), torch.no_grad():
                    self.batch.set_stream(side.cuda_stream)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        for k in range(n):
                            self._body(k)
                self._graph = g
            except Exception as ex:  # a policy that cannot be captured: fall back to eager launches for good
                self.use_graph = False
                self._graph = None
                self.graph_error = repr(ex)[:300]
            finally:
                self.batch.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
                torch.cuda.current_stream(self.device).wait_stream(side)
            if self._graph is None:
                return self._run_chunk(n)
            # (capturing does not execute: fall through to the replay)
        self._graph.replay()

    def collect(self, max_steps: Optional[int] = None) -> dict:
        """One episode of every replica (or ``max_steps`` env-steps).  Returns time-major CUDA tensors: ``valid`` bool [T][B]
        (replica had a decision at step t), ``ticks`` / ``ports`` / ``vessels`` / ``model_actions`` int32 [T][B], ``states``
        float32 [T][B][D] (if ``store_states``), ``rewards`` float32 [T][B] (0 where not valid), ``metrics`` int64 [B][3]
        (order_requirements, container_shortage, operation_number of the finished episode), ``last_tick`` int32 [B]."""
        torch, env = self._torch, self.batch
        B = env.n_replicas
        T = self.n_steps if max_steps is None else min(int(max_steps), self.n_steps)
        K = self.graph_chunk
        env.reset()
        self._final_met.zero_()
        env.step_device(self._c_dec[0].data_ptr(), self._met.data_ptr())  # generator start: first decisions
        decs = torch.empty((T, B, 8), dtype=torch.int32, device=self.device)
        acts = torch.empty((T, B), dtype=torch.int32, device=self.device)
        states = torch.empty((T, B, self.shaper.state_dim), dtype=torch.float32, device=self.device) if self.store_states else None
        t = 0
        while t < T:
            n = min(K, T - t)
            self._run_chunk(n)
            decs[t:t + n].copy_(self._c_dec[:n])
            acts[t:t + n].copy_(self._c_act[:n])
            if states is not None:
                states[t:t + n].copy_(self._c_state[:n])
            self._c_dec[0].copy_(self._c_dec[n])  # the next chunk's input
            t += n
        torch.maximum(self._final_met, self._met, out=self._final_met)  # (the last step's metrics)
        valid = decs[:, :, 6] == 0
        ticks = torch.where(valid, decs[:, :, 0], torch.full_like(decs[:, :, 0], -1)).contiguous()
        ports = decs[:, :, 1].contiguous()
        rewards = self.shaper.rewards_batch(ticks, ports)  # one launch over [T, B]
        rewards = torch.where(valid, rewards, torch.zeros_like(rewards))
        out = {"valid": valid, "ticks": decs[:, :, 0].contiguous(), "ports": ports, "vessels": decs[:, :, 2].contiguous(),
               "model_actions": acts, "rewards": rewards, "metrics": self._final_met.clone(), "decisions": decs,
               "last_tick": torch.as_tensor(env.ticks(), device=self.device)}
        if states is not None:
            out["states"] = states
        self.last = out
        return out

    # ------------------------------------------------------------------------------------------------ reference surface
    def sample(self, env_indices: Optional[Sequence[int]] = None, collect: bool = True) -> dict:
        """The reference sampler's result for the environments ``env_indices`` (default: all) of one freshly collected episode
        (``collect=False``: of the last ``collect()``): ``{"experiences": [[ExpElement ...] per env], "info": [{"env_metric":
        ...} per env]}`` — what ``TrainingManager.record_experiences`` consumes (maro/rl/training/training_manager.py:105-117)."""
        traj = self.collect() if collect or self.last is None else self.last
        if "states" not in traj:
            raise RuntimeError("sample() needs the states: build the sampler with store_states=True")
        Exp = _exp_element_cls()
        idx = list(range(self.batch.n_replicas)) if env_indices is None else [int(i) for i in env_indices]
        sel = self._torch.as_tensor(idx, device=self.device)
        valid = traj["valid"][:, sel].cpu().numpy()
        ticks = traj["ticks"][:, sel].cpu().numpy()
        ports = traj["ports"][:, sel].cpu().numpy()
        acts = traj["model_actions"][:, sel].cpu().numpy()
        rewards = traj["rewards"][:, sel].cpu().numpy()
        states = traj["states"][:, sel].cpu().numpy().astype(np.float64)  # (the reference's states are float64 query results)
        metrics = traj["metrics"][sel].cpu().numpy()
        last_tick = traj["last_tick"][sel].cpu().numpy()
        done = traj["decisions"][-1][sel][:, 6].cpu().numpy() != 0
        experiences, info = [], []
        for j in range(len(idx)):
            steps = np.flatnonzero(valid[:, j])
            # per agent (= port): the next decision of the same port provides next_agent_state; the agent's last cached
            # transition is terminal iff the episode ended (env_sampler.py:402-424)
            nxt: Dict[int, int] = {}
            next_of = np.full(len(steps), -1, np.int64)
            for pos in range(len(steps) - 1, -1, -1):
                p = int(ports[steps[pos], j])
                next_of[pos] = nxt.get(p, -1)
                nxt[p] = pos
            bound = int(last_tick[j]) - self.reward_eval_delay  # transitions whose reward window is complete (:500-506)
            elems = []
            for pos, t in enumerate(steps):
                if ticks[t, j] > bound:
                    break
                p = int(ports[t, j])
                st = states[t, j]
                last_of_agent = next_of[pos] < 0
                nst = st if last_of_agent else states[steps[next_of[pos]], j]
                elems.append(Exp(tick=int(ticks[t, j]), state=st, agent_state_dict={p: st}, action_dict={p: np.asarray([acts[t, j]])},
                                 reward_dict={p: np.float32(rewards[t, j])}, terminal_dict={p: bool(done[j]) if last_of_agent else False},
                                 next_state=states[steps[pos + 1], j] if pos + 1 < len(steps) else None,
                                 next_agent_state_dict={p: nst}, truncated=False))
            experiences.append(elems)
            info.append({"env_metric": {"order_requirements": int(metrics[j, 0]), "container_shortage": int(metrics[j, 1]),
                                        "operation_number": int(metrics[j, 2])}})
        return {"experiences": experiences, "info": info}


class CimDeviceRollout(BatchedCimEnvSampler):
    """Round-1 name of the device rollout loop; ``run_episode`` = ``collect``."""

    def __init__(self, batch: CimBatch, policy: Callable, shaper: Optional[CimShaper] = None, store_states: bool = True, **kw):
        super().__init__(batch, policy, shaper, store_states=store_states, **kw)

    def run_episode(self, max_steps: int = 1 << 30) -> dict:
        return self.collect(None if max_steps >= (1 << 29) else max_steps)
