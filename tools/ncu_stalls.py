"""Stall-reason totals of an ncu report (source page), split into the writer's streaming instructions
(STG.E.EF.128 / the LDS.128 feeding them) and everything else, plus the top SASS instructions by samples.
    python tools/ncu_stalls.py rep.ncu-rep [top]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
si = hdr.index("Source"); wi = hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
groups = defaultdict(lambda: defaultdict(int))
tot = 0
for r in data:
    s = r[si]
    g = "writer STG.128" if "STG.E.EF.128" in s else ("LDS.128" if "LDS.128" in s else ("other STG" if "STG" in s else "rest"))
    for i, h in stall_cols:
        groups[g][h] += int(r[i] or 0)
    tot += int(r[wi] or 0)
print("total samples", tot)
for g, d in groups.items():
    t = sum(d.values())
    print(f"{g:16s} {t:8d} ({100*t/tot:5.1f} %)  " + "  ".join(f"{h[6:]}={v}" for h, v in sorted(d.items(), key=lambda kv: -kv[1])[:7] if v))
print("top instructions:")
for r in sorted(data, key=lambda r: -int(r[wi] or 0))[:top]:
    d = {h[6:]: int(r[i] or 0) for i, h in stall_cols if int(r[i] or 0)}
    print(f"  {int(r[wi]):7d}  {r[si].strip()[:70]:70s} " + " ".join(f"{k}={v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:4]))
