"""Top stall lines of a kernel from an .ncu-rep source page (SASS view).
    python tools/ncu_hot.py rep kernel-regex [topN]
"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may be concatenated; take the first table
hdr = rows[1]
body = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Address":
        break
    body.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {s: sum(int(r[ci[s]] or 0) for r in body) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]]))[:top]
for i in sorted(order):
    r = body[i]
    st = {s[6:]: int(r[ci[s]] or 0) for s in stalls if int(r[ci[s]] or 0)}
    print(f"{i:5d} {int(r[ci['# Samples']]):7d} {r[ci['Instructions Executed']]:>9s}  {r[ci['Source']].strip()[:90]:90s} {st}")
