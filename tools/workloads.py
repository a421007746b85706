"""Workload definitions shared by bench.py and the tools (no test imports): where the frozen citi_bike toy dataset
lives and the decision config it runs with (the reference's toy.3s_4t topology, citi_bike/topologies/toy.3s_4t)."""
import os

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIKE_TOY_DIR = os.path.join(ROOT, "tests", "golden", "bike_toy")  # data files only (trips.bin, weather, stations, distances)

BIKE_TOY_DECISION = """decision:
  extra_cost_mode: source
  resolution: 20
  effective_time_mean: 20
  effective_time_std: 5
  supply_water_mark_ratio: 0.8
  demand_water_mark_ratio: 0.2
  action_scope:
    low: 0
    high: 1
    filters:
      - type: "distance"
        num: 80
      - type: "requirements"
        num: 40
      - type: "trip_window"
        windows: 10
        num: 20
time_zone: "America/New_York"
"""

BIKE_TOY_FILES = dict(trip_data="trips.bin", weather_data="KNYC_daily.bin", stations_init_data="station_meta.csv",
                      distance_adj_data="distance_adj.csv")


def bike_toy_config() -> dict:
    """config dict for maro_b200.scenarios.citi_bike.data.build_bike_topology"""
    conf = yaml.safe_load(BIKE_TOY_DECISION)
    for k, v in BIKE_TOY_FILES.items():
        conf[k] = os.path.join(BIKE_TOY_DIR, v)
    return conf


def bike_toy_config_dir() -> str:
    """a temp folder with a config.yml of absolute paths (what the reference's Env(topology=<folder>) wants)"""
    import tempfile

    d = tempfile.mkdtemp()
    with open(os.path.join(d, "config.yml"), "w") as fp:
        fp.write(BIKE_TOY_DECISION)
        for k, v in BIKE_TOY_FILES.items():
            fp.write(f'{k}: "{os.path.join(BIKE_TOY_DIR, v)}"\n')
    return d


def hash_u32(x: int) -> int:
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def cim_policy_random(dec, seed: int, replica: int, step: int):
    """the hashed hello-world agent (examples/hello_world/cim/hello.py:24-32), same arithmetic as cim_policy_kernel:
    (vessel, port, quantity, action type) from a decision row [tick, port, vessel, scope.load, scope.discharge, ...]"""
    h1 = hash_u32(seed ^ hash_u32((replica * 0x9E3779B9 + step * 0x85EBCA6B + 0x1234567) & 0xFFFFFFFF))
    h2 = hash_u32((h1 + 0x68BC21EB) & 0xFFFFFFFF)
    load, dis = dec[3], dec[4]
    to_discharge = dis > 0 and (h1 & 1)
    scope = dis if to_discharge else load
    qty = h2 % (scope + 1) if scope > 0 else 0
    return dec[2], dec[1], int(qty), 1 if to_discharge else 0


def bike_greedy(dec_event):
    """examples/citi_bike/greedy/launcher.py:35-65 with top_k = 1 on a reference DecisionEvent -> (value, candidate)"""
    best = None
    for cand, v in dec_event.action_scope.items():
        if cand == dec_event.station_idx:
            continue
        if best is None or (v, cand) > best:
            best = (v, cand)
    return best
