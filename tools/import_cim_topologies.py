"""Convert CIM topology configs (YAML, reference layout) into the compact JSON this package ships.

Topology files are DATA (port/vessel/route parameters), not code. They are re-encoded as canonical JSON
(`json.dumps(..., sort_keys=False)` keeps the YAML key order, which defines port / vessel / route indices;
Python float repr round-trips every float exactly) so that built-in topology names such as
``toy.4p_ssdd_l0.0`` resolve without the reference being installed.

Usage (in the build container, where /root/reference exists):
    python tools/import_cim_topologies.py [/root/reference]
"""
import json
import os
import sys

import yaml


def main() -> None:
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src_root = os.path.join(ref, "maro/simulator/scenarios/cim/topologies")
    here = os.path.dirname(os.path.abspath(__file__))
    dst_root = os.path.join(here, "..", "maro_b200", "scenarios", "cim", "topologies")
    os.makedirs(dst_root, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src_root)):
        cfg = os.path.join(src_root, name, "config.yml")
        if not os.path.isfile(cfg):
            continue
        with open(cfg) as fp:
            conf = yaml.safe_load(fp)
        with open(os.path.join(dst_root, name + ".json"), "w") as fp:
            json.dump(conf, fp, separators=(",", ":"))
        n += 1
    # the 22-port noisy topology the reference's own CIM tests run on
    test_cfg = os.path.join(ref, "tests/data/cim/case_data/config_folder/config.yml")
    if os.path.isfile(test_cfg):
        with open(test_cfg) as fp:
            conf = yaml.safe_load(fp)
        out = os.path.join(here, "..", "tests", "golden", "cim_case_config.json")
        with open(out, "w") as fp:
            json.dump(conf, fp, separators=(",", ":"))
    print("converted", n, "topologies")


if __name__ == "__main__":
    main()
