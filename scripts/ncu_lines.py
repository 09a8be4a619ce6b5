"""Warp-state samples of one kernel in an .ncu-rep, attributed to source lines of the library it was built from
(`ncu --import-source on` keeps SASS only; line numbers come from `nvdisasm --print-line-info` on the same .so).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [kernel mangled name] [lib.so]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
fn = sys.argv[2] if len(sys.argv) > 2 else "_ZN2rb16rssm_rows_kernelILi1EEEvNS_10RowsParamsE"
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "repo_b200", "librepo_b200.so")

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, capture_output=True, text=True).stdout
lines = sass.split("\n")
start = next(i for i, l in enumerate(lines) if ".section" in l and ".text." + fn in l)
end = next((i for i in range(start + 1, len(lines)) if ".section" in lines[i] and ".text." in lines[i]), len(lines))
off2line, cur = {}, None
for l in lines[start:end]:
    m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur

rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per_line = collections.Counter()
per_line_stall = collections.defaultdict(collections.Counter)
total = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    n = int(r[ix["# Samples"]] or 0)
    ln = off2line.get(off)
    per_line[ln] += n
    total += n
    for s in stalls:
        per_line_stall[ln][s] += int(r[ix[s]] or 0)
print(f"{rep}: {total} samples over {len(data)} SASS instructions; top source lines:")
src_cache = {}
for ln, n in per_line.most_common(40):
    txt = ""
    if ln:
        path = os.path.join(ROOT, "repo_b200", "csrc", ln[0])
        if os.path.exists(path):
            src_cache.setdefault(path, open(path).read().split("\n"))
            txt = src_cache[path][ln[1] - 1].strip()[:70]
    top = ", ".join(f"{k[6:]} {v}" for k, v in per_line_stall[ln].most_common(3))
    print(f"  {100.0 * n / total:5.1f}%  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<5d} {txt:<70s} [{top}]")
