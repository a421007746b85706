"""Synthetic citi_bike dataset with 26 stations (seeded; the reference ships only 3-5 station toys and a network download):
station table, distance matrix, one day of trips in the MARO .bin format.  Used by the action-scope-filter golden cases
(filters that DROP neighbours need more stations than a filter keeps).

    python tests/golden/bike_synth_gen.py        ->  tests/golden/bike_synth26/{stations.csv,distance_adj.csv,trips.bin,weathers.bin,decision.yml}

trips.bin is written by maro_b200.data_lib.BinaryConverter (byte-compatible with the reference's, tests/test_data_lib.py) from
a CSV in the reference's trips schema (tests/golden/data_lib/trips.meta.yml); the weather file is the reference fixture."""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
OUT = os.path.join(HERE, "bike_synth26")
S, TICKS = 26, 1440

DECISION = """decision:
  extra_cost_mode: target
  resolution: 20
  effective_time_mean: 12
  effective_time_std: 4
  supply_water_mark_ratio: 0.7
  demand_water_mark_ratio: 0.25
  action_scope:
    low: 0.1
    high: 0.9
    filters:
      - type: "distance"
        num: 14
      - type: "requirements"
        num: 9
      - type: "trip_window"
        windows: 5
        num: 4
time_zone: "America/New_York"
"""


def main():
    sys.path.insert(0, ROOT)
    from maro_b200.data_lib import BinaryConverter

    rng = np.random.RandomState(20260923)
    os.makedirs(OUT, exist_ok=True)
    cap = rng.randint(12, 40, S)
    init = (cap * rng.uniform(0.3, 0.7, S)).astype(int)
    xy = rng.uniform(0, 10, (S, 2))
    with open(os.path.join(OUT, "stations.csv"), "w") as fp:
        fp.write("station_index,capacity,init,station_id\n")
        for i in range(S):
            fp.write(f"{i},{cap[i]},{init[i]},{1000 + 7 * i}\n")
    d = np.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    with open(os.path.join(OUT, "distance_adj.csv"), "w") as fp:
        fp.write(",".join(str(i) for i in range(S)) + "\n")
        for i in range(S):
            fp.write(",".join("0" if i == j else repr(round(float(d[i, j]), 6)) for j in range(S)) + "\n")
    # trips: popular stations attract / emit more; a morning and an evening wave, so that stations run dry and fill up
    pop = rng.dirichlet(np.ones(S) * 0.6)
    csv_path = os.path.join(OUT, "_trips.csv")
    with open(csv_path, "w") as fp:
        fp.write("start_time,duration,start_station_index,end_station_index\n")
        for t in range(TICKS):
            wave = 1.0 + 2.5 * np.exp(-((t - 480) / 90.0) ** 2) + 2.5 * np.exp(-((t - 1080) / 90.0) ** 2)
            for _ in range(rng.poisson(1.6 * wave)):
                src = rng.choice(S, p=pop)
                dst = rng.choice(S, p=pop[::-1] if t < 780 else pop)
                if dst == src:
                    dst = (src + 1 + rng.randint(S - 1)) % S
                fp.write(f"2019-01-01 {t // 60:02d}:{t % 60:02d}:00,{int(rng.randint(3, 45))},{src},{dst}\n")
    conv = BinaryConverter(os.path.join(OUT, "trips.bin"), os.path.join(HERE, "data_lib", "trips.meta.yml"))
    conv.add_csv(csv_path)
    conv.close()
    os.remove(csv_path)
    shutil.copy(os.path.join(HERE, "bike_case_1", "weathers.bin"), os.path.join(OUT, "weathers.bin"))
    with open(os.path.join(OUT, "decision.yml"), "w") as fp:
        fp.write(DECISION)
    print(sorted(os.listdir(OUT)), os.path.getsize(os.path.join(OUT, "trips.bin")))


if __name__ == "__main__":
    main()
