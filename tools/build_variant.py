"""Build an experimental variant of the CUDA library next to the product one:
    python tools/build_variant.py <tag> [-DMACRO ...]   ->  maro_b200/libmaro_b200_<tag>.so
Select it at run time with MARO_B200_LIB=<path>.  Used for A/B measurements on the GPU box; never shipped."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

tag, defs = sys.argv[1], sys.argv[2:]
csrc = os.path.join(ROOT, "maro_b200", "csrc")
out = os.path.join(ROOT, "maro_b200", f"libmaro_b200_{tag}.so")
units = [os.path.join(csrc, f) for f in ("cim_env.cu", "bike_env.cu", "vm_env.cu")]
subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + defs + ["--threads", "3"] + units + ["-o", out])
print(out)
