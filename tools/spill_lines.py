"""Where does a kernel variant spill?  Counts LDL / STL per source line of one instantiation in libsrsce.so.
    python tools/spill_lines.py estimate_kernelILi512ELi2ELb0ELb0ELb1E"""
import os, re, subprocess, sys, tempfile
from collections import defaultdict
from pathlib import Path
root = Path(__file__).resolve().parents[1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", str(root / "srsran_ce_pytorch_b200" / "libsrsce.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("srsce_kernels")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
files = {}
for variant in sys.argv[1:]:
    on, cur, curf, by, total = False, None, None, defaultdict(lambda: [0, 0]), 0
    for ln in sass:
        if ln.startswith(".text."):
            on = variant in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "(.*)", line (\d+)', ln)
        if m:
            curf, cur = m.group(1), int(m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            total += 1
            if " LDL" in ln:
                by[(curf, cur)][0] += 1
            if " STL" in ln:
                by[(curf, cur)][1] += 1
    print(variant, "instructions", total)
    for (f, l), v in sorted(by.items()):
        if f not in files:
            files[f] = open(f).read().splitlines() if os.path.exists(f) else []
        text = files[f][l - 1].strip()[:100] if l and l <= len(files[f]) else ""
        print(f"   {os.path.basename(f)}:{l}  LDL {v[0]} STL {v[1]}  {text}")
