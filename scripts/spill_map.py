"""Per-source-line attribution of register spills (STL/LDL) in one kernel of librepo_b200.so, from
`nvdisasm --print-line-info` (the library is built with -lineinfo).

    python scripts/spill_map.py [mangled kernel name]

The rows kernel runs with 228 KB of shared memory, i.e. next to no L1: a spill is an L2 round trip."""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "repo_b200", "librepo_b200.so")
fn = sys.argv[1] if len(sys.argv) > 1 else "_ZN2rb16rssm_rows_kernelILi1EEEvNS_10RowsParamsE"

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, capture_output=True, text=True).stdout
lines = sass.split("\n")
start = next(i for i, l in enumerate(lines) if ".section" in l and ".text." + fn in l)
end = next((i for i in range(start + 1, len(lines)) if ".section" in lines[i] and ".text." in lines[i]), len(lines))
cur = None
spill, total = Counter(), Counter()
for l in lines[start:end]:
    m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", l):
        total[cur] += 1
        if " STL" in l or " LDL" in l:
            spill[cur] += 1
print(f"{fn}: {sum(total.values())} SASS instructions, {sum(spill.values())} of them local-memory spills")
for k, v in sorted(spill.items(), key=lambda kv: -kv[1])[:30]:
    print(f"  {k[0]}:{k[1]:<5d} {v:4d} spill / {total[k]:4d} instructions")
