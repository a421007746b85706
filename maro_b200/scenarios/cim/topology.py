"""Host-side CIM topology loader: config (YAML folder or built-in JSON) -> flat static tables for the device.

This is init-time work (once per distinct seed), not the hot path.  It restates, from scratch, what the
reference does in ``maro/data_lib/cim/cim_data_generator.py:18-205`` (route unrolling into Stop tables),
``maro/data_lib/cim/parsers.py:14-211`` (ports / vessels / routes / global order proportion) and the seed
bookkeeping of ``maro/simulator/utils/sim_random.py:48-63`` (named MT19937 streams, stream *i* in creation
order gets ``seed + i``).  CPython's ``random.Random`` *is* MT19937 with the exact ``uniform`` the reference
calls, so stop tables are bit-identical for noisy topologies as well.

Everything the per-event wrappers of the reference compute lazily (future-stop prediction, sailing plan,
reachable stops, past stops; ``maro/data_lib/cim/vessel_*_wrapper.py``) is a pure function of these tables and
is evaluated inside the kernel from ``stop_*`` + ``vessel_leg`` + ``route_port``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from math import ceil, floor
from random import Random
from typing import Dict, List, Optional

import numpy as np

_TOPOLOGY_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "topologies")

#: creation order of the reference's named random streams in a fresh process
STREAMS_NOISY = ("order_init", "route_init", "order_number", "buffer_time")
STREAMS_PLAIN = ("route_init", "order_number", "buffer_time")

DATA_CONTAINER_INIT_SEED_LIMIT = 4096  # maro/data_lib/cim/utils.py:12


def builtin_topologies() -> List[str]:
    return sorted(f[:-5] for f in os.listdir(_TOPOLOGY_DIR) if f.endswith(".json"))


def load_config(topology: str) -> dict:
    """Resolve a topology the way ``AbsBusinessEngine`` does (abs_business_engine.py:136-164): an existing
    folder containing ``config.yml`` wins, otherwise a built-in topology name."""
    if os.path.isdir(topology):
        cfg = os.path.join(topology, "config.yml")
        if not os.path.isfile(cfg):
            raise FileNotFoundError(cfg)
        import yaml

        with open(cfg) as fp:
            return yaml.safe_load(fp)
    if os.path.isfile(topology):
        with open(topology) as fp:
            if topology.endswith(".json"):
                return json.load(fp)
            import yaml

            return yaml.safe_load(fp)
    path = os.path.join(_TOPOLOGY_DIR, topology + ".json")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"unknown CIM topology {topology!r}")
    with open(path) as fp:
        return json.load(fp)


def _apply_noise(value, noise, rand: Random) -> float:
    # maro/data_lib/cim/utils.py:30-42
    return value + rand.uniform(-noise, noise)


@dataclass
class CimTopology:
    """Flat static tables of one (config, max_tick, seed) triple."""

    config: dict
    seed: int
    max_tick: int
    n_ports: int
    n_vessels: int
    n_routes: int
    past_stop_number: int
    future_stop_number: int
    order_mode: int  # 0 fixed / 1 unfixed
    total_containers: int
    container_volume: float
    load_cost_factor: float
    dsch_cost_factor: float
    port_names: List[str]
    vessel_names: List[str]
    route_names: List[str]
    # ports
    port_capacity: np.ndarray
    port_init_empty: np.ndarray
    full_return_base: np.ndarray
    full_return_noise: np.ndarray
    empty_return_base: np.ndarray
    empty_return_noise: np.ndarray
    source_base: np.ndarray
    source_noise: np.ndarray
    target_offset: np.ndarray
    target_port: np.ndarray
    target_base: np.ndarray
    target_noise: np.ndarray
    # vessels
    vessel_capacity: np.ndarray
    vessel_init_empty: np.ndarray
    vessel_route: np.ndarray
    vessel_period: np.ndarray
    vessel_route_start: np.ndarray  # offset of the start port inside the route
    vessel_leg: np.ndarray  # [sum over vessels of route_len] no-noise duration + ceil(dist/speed) per route position
    vessel_leg_offset: np.ndarray
    stop_offset: np.ndarray
    stop_arrival: np.ndarray
    stop_leave: np.ndarray
    stop_port: np.ndarray
    # routes
    route_offset: np.ndarray
    route_port: np.ndarray
    route_distance: np.ndarray
    order_proportion: np.ndarray
    # random stream seeds (fresh-process creation order)
    stream_seeds: Dict[str, int] = field(default_factory=dict)
    #: state of the route_init stream after route unrolling (needed for reset(keep_seed=False))
    route_init_state: Optional[tuple] = None

    @property
    def has_order_noise(self) -> bool:
        return bool(np.any(self.source_noise != 0) or np.any(self.target_noise != 0))

    @property
    def has_buffer_noise(self) -> bool:
        return bool(np.any(self.full_return_noise != 0) or np.any(self.empty_return_noise != 0))

    @property
    def max_route_len(self) -> int:
        return int(np.max(np.diff(self.route_offset)))

    def max_event_delay(self) -> int:
        """Largest (target tick - current tick) of any dynamic event: DISCHARGE_FULL to a reachable stop
        (business_engine.py:583-587) or a buffered RETURN_FULL / RETURN_EMPTY (:480-497, :683-693)."""
        d = 1
        for v in range(self.n_vessels):
            lo, hi = int(self.stop_offset[v]), int(self.stop_offset[v + 1])
            arr = self.stop_arrival[lo:hi]
            rl = int(self.route_offset[self.vessel_route[v] + 1] - self.route_offset[self.vessel_route[v]])
            for k in range(1, rl + 1):
                if len(arr) > k:
                    d = max(d, int(np.max(arr[k:] - arr[:-k])))
        for base, noise in ((self.full_return_base, self.full_return_noise),
                            (self.empty_return_base, self.empty_return_noise)):
            if len(base):
                d = max(d, int(ceil(float(np.max(base + np.abs(noise))))) + 1)
        return d

    def stops_of(self, vessel: int):
        lo, hi = int(self.stop_offset[vessel]), int(self.stop_offset[vessel + 1])
        return list(zip(self.stop_arrival[lo:hi].tolist(), self.stop_leave[lo:hi].tolist(),
                        self.stop_port[lo:hi].tolist()))


def build_topology(topology, max_tick: int, seed: Optional[int] = None, start_tick: int = 0) -> CimTopology:
    """Generate the static tables for ``topology`` (name, folder, file or already-loaded dict)."""
    conf = topology if isinstance(topology, dict) else load_config(topology)
    if seed is None:
        seed = int(conf["seed"])

    usage = conf["container_usage_proportion"]
    sample_noise = usage["sample_noise"]
    # which named streams exist, in creation order, decides each stream's seed offset
    # (parsers.py:98-100: "order_init" is only touched when sample_noise != 0 and a ratio is non-zero)
    period = int(usage["period"])
    sample_nodes = [(x, y) for x, y in usage["sample_nodes"]]
    if sample_nodes[0][0] != 0:
        sample_nodes.insert(0, (0, 0))
    if sample_nodes[-1][0] != period - 1:
        sample_nodes.append((period - 1, 0))
    xp = [n[0] for n in sample_nodes]
    yp = [n[1] for n in sample_nodes]
    dist = np.interp(list(range(period)), xp, yp)

    uses_order_init = sample_noise != 0 and any(dist[t % period] != 0 for t in range(start_tick, max_tick))
    names = STREAMS_NOISY if uses_order_init else STREAMS_PLAIN
    stream_seeds = {name: seed + i for i, name in enumerate(names)}

    total_containers = conf["total_containers"]
    past_n, future_n = conf["stop_number"]

    # ---- global order proportion (parsers.py:63-107)
    order_proportion = np.zeros(max_tick - start_tick, dtype=np.int32)
    order_init = Random(stream_seeds["order_init"]) if uses_order_init else None
    for t in range(start_tick, max_tick):
        orders = dist[t % period]
        if orders != 0:
            if sample_noise != 0:
                orders = _apply_noise(orders, sample_noise, order_init)
            orders = floor(max(0, min(1, orders)) * total_containers)
        order_proportion[t - start_tick] = orders

    # ---- ports (parsers.py:133-211)
    ports_conf = conf["ports"]
    port_names = list(ports_conf.keys())
    port_idx = {n: i for i, n in enumerate(port_names)}
    total_ratio = sum(p["initial_container_proportion"] for p in ports_conf.values())
    assert round(total_ratio, 7) == 1
    P = len(port_names)
    port_capacity = np.zeros(P, np.int32)
    port_init_empty = np.zeros(P, np.int32)
    frb, frn, erb, ern, sb, sn = (np.zeros(P, np.float64) for _ in range(6))
    target_offset = np.zeros(P + 1, np.int32)
    t_port, t_base, t_noise = [], [], []
    for i, (name, info) in enumerate(ports_conf.items()):
        port_capacity[i] = info["capacity"]
        port_init_empty[i] = int(info["initial_container_proportion"] * total_containers)
        erb[i], ern[i] = info["empty_return"]["buffer_ticks"], info["empty_return"]["noise"]
        frb[i], frn[i] = info["full_return"]["buffer_ticks"], info["full_return"]["noise"]
        dconf = info["order_distribution"]
        sb[i], sn[i] = dconf["source"]["proportion"], dconf["source"]["noise"]
        for tname, tconf in (dconf.get("targets") or {}).items():
            t_port.append(port_idx[tname])
            t_base.append(tconf["proportion"])
            t_noise.append(tconf["noise"])
        target_offset[i + 1] = len(t_port)

    # ---- routes (parsers.py:110-130)
    routes_conf = conf["routes"]
    route_names = list(routes_conf.keys())
    route_idx = {n: i for i, n in enumerate(route_names)}
    route_offset = np.zeros(len(route_names) + 1, np.int32)
    r_port, r_dist, r_port_name = [], [], []
    for i, (name, pts) in enumerate(routes_conf.items()):
        for pt in pts:
            r_port.append(port_idx[pt["port_name"]])
            r_port_name.append(pt["port_name"])
            r_dist.append(pt["distance_to_next_port"])
        route_offset[i + 1] = len(r_port)

    # ---- vessels + route unrolling (parsers.py:14-60, cim_data_generator.py:18-115)
    vessels_conf = conf["vessels"]
    vessel_names = list(vessels_conf.keys())
    V = len(vessel_names)
    vessel_capacity = np.zeros(V, np.int32)
    vessel_init_empty = np.zeros(V, np.int32)
    vessel_route = np.zeros(V, np.int32)
    vessel_period = np.zeros(V, np.int32)
    vessel_route_start = np.zeros(V, np.int32)
    vessel_leg_offset = np.zeros(V + 1, np.int32)
    stop_offset = np.zeros(V + 1, np.int32)
    legs: List[int] = []
    s_arr, s_leave, s_port = [], [], []
    route_init = Random(stream_seeds["route_init"])
    for vi, (vname, vnode) in enumerate(vessels_conf.items()):
        r = route_idx[vnode["route"]["route_name"]]
        lo, hi = int(route_offset[r]), int(route_offset[r + 1])
        rl = hi - lo
        vessel_capacity[vi] = vnode["capacity"]
        vessel_init_empty[vi] = vnode.get("empty", 0)
        vessel_route[vi] = r
        speed, speed_noise = vnode["sailing"]["speed"], vnode["sailing"]["noise"]
        duration, duration_noise = vnode["parking"]["duration"], vnode["parking"]["noise"]
        loc = 0
        while r_port_name[lo + loc] != vnode["route"]["initial_port_name"]:
            loc += 1
        vessel_route_start[vi] = loc
        for k in range(rl):
            legs.append(duration + ceil(r_dist[lo + k] / speed))
        vessel_leg_offset[vi + 1] = len(legs)

        tick = 0
        period_no_noise = 0
        extra = 0
        n_stops = 0
        while extra <= future_n:
            parking = ceil(_apply_noise(duration, duration_noise, route_init))
            assert parking > 0
            s_arr.append(tick)
            s_leave.append(tick + parking)
            s_port.append(r_port[lo + loc])
            n_stops += 1
            d = r_dist[lo + loc]
            noised_speed = _apply_noise(speed, speed_noise, route_init)
            tick += parking + ceil(d / noised_speed)
            period_no_noise += (duration + ceil(d / speed)) if n_stops <= rl else 0
            loc = (loc + 1) % rl
            extra += 1 if tick > max_tick else 0
        vessel_period[vi] = period_no_noise
        stop_offset[vi + 1] = len(s_arr)

    return CimTopology(
        config=conf,
        seed=seed,
        max_tick=max_tick,
        n_ports=P,
        n_vessels=V,
        n_routes=len(route_names),
        past_stop_number=int(past_n),
        future_stop_number=int(future_n),
        order_mode=0 if conf["order_generate_mode"] == "fixed" else 1,
        total_containers=int(total_containers),
        container_volume=float(conf["container_volumes"][0]),
        load_cost_factor=float(conf["load_cost_factor"]),
        dsch_cost_factor=float(conf["dsch_cost_factor"]),
        port_names=port_names,
        vessel_names=vessel_names,
        route_names=route_names,
        port_capacity=port_capacity,
        port_init_empty=port_init_empty,
        full_return_base=frb,
        full_return_noise=frn,
        empty_return_base=erb,
        empty_return_noise=ern,
        source_base=sb,
        source_noise=sn,
        target_offset=target_offset,
        target_port=np.asarray(t_port, np.int32),
        target_base=np.asarray(t_base, np.float64),
        target_noise=np.asarray(t_noise, np.float64),
        vessel_capacity=vessel_capacity,
        vessel_init_empty=vessel_init_empty,
        vessel_route=vessel_route,
        vessel_period=vessel_period,
        vessel_route_start=vessel_route_start,
        vessel_leg=np.asarray(legs, np.int32),
        vessel_leg_offset=vessel_leg_offset,
        stop_offset=stop_offset,
        stop_arrival=np.asarray(s_arr, np.int32),
        stop_leave=np.asarray(s_leave, np.int32),
        stop_port=np.asarray(s_port, np.int32),
        route_offset=route_offset,
        route_port=np.asarray(r_port, np.int32),
        route_distance=np.asarray(r_dist, np.int32),
        order_proportion=order_proportion,
        stream_seeds=stream_seeds,
        route_init_state=route_init.getstate(),
    )


def next_topology_seed(topo: CimTopology) -> int:
    """Seed drawn by ``reset(keep_seed=False)`` (cim_data_container_helpers.py:56-66): one
    ``randint(0, 4095)`` from the route_init stream as left behind by route unrolling."""
    r = Random()
    r.setstate(topo.route_init_state)
    return r.randint(0, DATA_CONTAINER_INIT_SEED_LIMIT - 1)
