"""Copy the judged artefacts of one tools/collect_profiles.sh run from gpurun_out/ (scratch) into profiles/ (tracked).
    python tools/publish_profiles.py <tag-in-gpurun_out> <round-prefix>      e.g.  s3 r2
Writes  profiles/<r>_bench.json            the plain bench line of the run (gpurun_out/<tag>_bench.json if present)
        profiles/<r>_launches.csv          ncu launch list of the same bench command
        profiles/<r>_vq_ncu_summary.txt    key metrics of the full capture of the VQ kernels
        profiles/<r>_gc_ncu_summary.txt    ... of gc_forward_kernel
        profiles/<r>_ncu_traffic.json      DRAM bytes per launch of the dominant kernels (what bench.py's `traffic` reads)
        profiles/<r>_sass_tcgen05.txt      tcgen05 / TMEM / TMA mnemonic counts per kernel of the built library
Runs here (no GPU): ncu -i and cuobjdump only read files."""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def num(s):
    return float(s.replace(",", "")) if s not in ("", None) else 0.0


bench = os.path.join(G, f"{tag}_bench.json")
if not os.path.exists(bench):
    bench = os.path.join(G, f"{tag}_plain.json")
line = [l for l in open(bench).read().splitlines() if l.startswith("{")][-1]
open(os.path.join(P, f"{rnd}_bench.json"), "w").write(line + "\n")
shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{rnd}_launches.csv"))

traffic = {}
for what, kern in (("vq", "vq_fused_kernel"), ("gc", "gc_forward_kernel")):
    rep = os.path.join(G, f"{tag}_{what}_full.ncu-rep")
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True,
                         text=True).stdout
    open(os.path.join(P, f"{rnd}_{what}_ncu_summary.txt"), "w").write(txt)
    # units: ask ncu for base units so that bytes are bytes
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    recs = [dict(zip(rows[0], r)) for r in rows[2:]]
    hit = [r for r in recs if kern in r["Kernel Name"]]
    if hit:
        r = hit[-1]
        traffic[kern] = {"dram_bytes_read": int(num(r["dram__bytes_read.sum"])),
                         "dram_bytes_write": int(num(r["dram__bytes_write.sum"])),
                         "gpu_time_us": num(r["gpu__time_duration.sum"]) / 1e3,
                         "source": f"{tag}_{what}_full.ncu-rep (ncu --set full --clock-control none, one launch)"}
json.dump(traffic, open(os.path.join(P, f"{rnd}_ncu_traffic.json"), "w"), indent=1)

lib = os.path.join(ROOT, "dc_vic_b200", "lib", "libdcvic_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
per = {}
cur = None
pat = re.compile(r"\b(UTC[A-Za-z0-9_.]+|UTMA[A-Za-z0-9_.]+|LDTM[A-Za-z0-9_.]*|STTM[A-Za-z0-9_.]*|SYNCS[A-Za-z0-9_.]*|UBLKCP[A-Za-z0-9_.]*)")
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    if cur:
        m = pat.search(l)
        if m:
            per.setdefault(cur, Counter())[m.group(1)] += 1
with open(os.path.join(P, f"{rnd}_sass_tcgen05.txt"), "w") as f:
    f.write("# SASS evidence (cuobjdump -sass dc_vic_b200/lib/libdcvic_b200.so): tcgen05 / TMEM / TMA mnemonics per kernel\n")
    for k in sorted(per):
        if not any(x.startswith(("UTC", "UTMA", "LDTM", "STTM")) for x in per[k]):
            continue
        f.write("== " + subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:140] + "\n")
        for mn, c in sorted(per[k].items()):
            f.write(f"{c:7d} {mn}\n")
print(json.dumps(traffic, indent=1))
