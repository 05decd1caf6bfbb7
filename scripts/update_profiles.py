#!/usr/bin/env python
"""Refresh profiles/ from the gpurun_out captures of one tag (scripts/gpu_profile.sh TAG):
per-kernel ncu summaries, the launch list + its per-kernel shares, traffic.json and the
bench line.   usage: update_profiles.py TAG [bench.json]"""
import collections
import csv
import io
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1]
RND = "r2" if tag.startswith("r2") else "r1"
ALG = {"c2": 121, "c3": 97, "c4": 457, "c5": 457}

traffic = {"_how": f"ncu --set full --clock-control none -k regex:step_kernel -c 1 after 500 propagation + 5 "
                   f"warm-up steps, python bench.py --workload W --steps 5 --warmup 5 (B200, capture {tag}, "
                   "scripts/gpu_profile_r2.sh); dram__bytes_read.sum + dram__bytes_write.sum of one launch"}
old = json.loads((PROF / "traffic.json").read_text())
for w in ("c2", "c3", "c4", "c5"):
    rep = OUT / f"prof_{w}_{tag}.ncu-rep"
    if not rep.exists():
        traffic[w] = old.get(w)
        continue
    txt = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_summary.py"), str(rep)],
                         capture_output=True, text=True).stdout
    (PROF / f"{RND}_ncu_{w}_step_kernel.txt").write_text(txt)
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(k):
        i = hdr.index(k)
        v = float(r[i].replace(",", ""))
        u = units[i]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3,
                 "s": 1e6}.get(u, 1)
        return v * scale
    traffic[w] = {"kernel": r[hdr.index("Kernel Name")],
                  "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                  "algorithmic_bytes": (old.get(w) or {}).get("algorithmic_bytes"),
                  "gpu_time_us": val("gpu__time_duration.sum")}
(PROF / "traffic.json").write_text(json.dumps(traffic, indent=1))

ll = OUT / f"launches_{tag}.csv"
if ll.exists():
    shutil.copy(ll, PROF / f"{RND}_launches_bench.csv")
    lines = [l for l in ll.read_text().splitlines() if not l.startswith("==")]
    rows = list(csv.reader(lines))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        us = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r[mu], 1e-3)
        name = r[kn].split("(")[0][:70]
        tot[name] += us
        cnt[name] += 1
    total = sum(tot.values())
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400   python bench.py --steps 20 --warmup 5 --propagate 100 --no-e2e --no-cpu --no-extras" if RND == "r2" else
           "ncu --metrics gpu__time_duration.sum --clock-control none -c 600   python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu",
           "(launch list of the default bench command, B200; per-launch times are cold-cache and serialised: compare SHARES)",
           f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>10s}"]
    for name, us in tot.most_common(25):
        out.append(f"{name:70s} {cnt[name]:8d} {us:12.1f} {100 * us / total:6.1f}% {us / cnt[name]:10.1f}")
    (PROF / f"{RND}_launch_list_summary.txt").write_text("\n".join(out) + "\n")
if len(sys.argv) > 2:
    shutil.copy(sys.argv[2], PROF / f"{RND}_bench_n1.json")
print("profiles/ refreshed from tag", tag)
