"""Aggregate an `ncu --page source --csv` (SASS view) dump by CUDA source line using nvdisasm -g line markers.

usage: python tools/ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top_n]
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    dis = ""  # one cubin per translation unit: take the one that holds the kernel
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        if any(ln.startswith("//---") and ".text." in ln and kern in ln for ln in d.splitlines()):
            dis = d
            break
    # address -> (file, line) for the requested kernel
    amap, cur, infn = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            infn = kern in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
        if m and cur:
            amap[int(m.group(1), 16)] = cur
    skip = os.environ.get("NCU_SKIP")  # NCU_SKIP=n: use the (n+1)-th launch of the report instead of the first
    extra = ["--launch-skip", skip, "--launch-count", "1"] if skip else []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # first kernel block only
    hdr = rows[1]
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            break
        data.append(dict(zip(hdr, r)))
    base = None
    agg = collections.defaultdict(lambda: [0.0, 0.0])
    tot_i = tot_s = 0.0
    for d in data:
        a = int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"])
        if base is None:
            base = a
        key = amap.get(a - base, ("?", 0))
        i = float(d["Instructions Executed"] or 0)
        s = float(d["# Samples"] or 0)
        agg[key][0] += i
        agg[key][1] += s
        tot_i += i
        tot_s += s
    src_cache = {}
    def src(f, l):
        for root in ("maro_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.isfile(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                L = src_cache[p]
                return L[l - 1].strip()[:100] if 0 < l <= len(L) else ""
        return ""
    print(f"kernel '{kern}': {tot_i:.0f} warp-instructions, {tot_s:.0f} samples")
    for (f, l), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*i/tot_i:5.1f}% inst {100*s/max(1,tot_s):5.1f}% smp  {f}:{l}  {src(f, l)}")


if __name__ == "__main__":
    main()
