"""Join the per-instruction stall samples of an ncu report with source lines.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep 'estimate_kernelILi2ELb0ELb0E' [top]

ncu's CSV source page is SASS level; nvdisasm -g gives the source line of every SASS instruction
of the same cubin, in the same order, so the two are joined by instruction index."""
import csv, re, subprocess, sys, tempfile, os
from collections import defaultdict
from pathlib import Path

rep, mangled = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = Path(__file__).resolve().parents[1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", str(root / "srsran_ce_pytorch_b200" / "libsrsce.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("srsce_kernels")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for ln in sass:
    if ln.startswith(".text."):
        on = mangled in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append((cur, ln.strip()))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
wi, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
assert len(data) == len(lines), (len(data), len(lines))
by_line, inst_line = defaultdict(int), defaultdict(int)
for (ln, _), r in zip(lines, data):
    by_line[ln] += int(r[wi] or 0)
    inst_line[ln] += int(r[ii] or 0)
tot, toti = sum(by_line.values()), sum(inst_line.values())
files = {}
print(f"total samples {tot}, warp instructions {toti}")
for key, s in sorted(by_line.items(), key=lambda kv: -kv[1])[:top]:
    f, ln = key if key else ("?", 0)
    if f not in files:
        files[f] = Path(f).read_text().splitlines() if os.path.exists(f) else []
    text = files[f][ln - 1].strip()[:100] if ln and ln <= len(files[f]) else "?"
    print(f"{os.path.basename(f)}:{ln:<5d} {100 * s / tot:5.1f}% samp {100 * inst_line[key] / toti:5.1f}% inst  {text}")
