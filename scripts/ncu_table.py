"""One line per kernel launch of an .ncu-rep: duration, tensor-pipe / issue / L1 / L2 / DRAM utilisation and DRAM bytes.

    python scripts/ncu_table.py gpurun_out/x.ncu-rep [labels,comma,separated] > profiles/x.txt
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
labels = sys.argv[2].split(",") if len(sys.argv) > 2 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
c = lambda n: hdr.index(n)
_T = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
_B = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def scaled(name, raw):
    """ncu picks a unit per report: bring durations to ms and byte counts to MB"""
    try:
        v = float(raw.replace(",", ""))
    except ValueError:
        return raw
    u = units[c(name)]
    if name == "gpu__time_duration.sum":
        v *= _T.get(u, 1.0)
    elif name.startswith("dram__bytes"):
        v *= _B.get(u, 1.0)
    return v
tens = "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"
cols = [("dur_ms", "gpu__time_duration.sum"), ("tensor%", tens), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("l1tex%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("dram_rd_MB", "dram__bytes_read.sum"),
        ("dram_wr_MB", "dram__bytes_write.sum"), ("regs", "launch__registers_per_thread"), ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic")]
print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised launches)")
print(f"{'launch':28s} {'kernel':22s} {'grid':>5s} " + " ".join(f"{n:>10s}" for n, _ in cols))
tot = 0.0
for i, r in enumerate(rows[2:]):
    name = r[c("Kernel Name")].split("(")[0].replace("rb::", "")[:22]
    lab = labels[i] if i < len(labels) else str(i)
    vals = []
    for n, k in cols:
        v = scaled(k, r[c(k)]) if k in hdr else "-"
        vals.append(f"{v:.3f}" if isinstance(v, float) else v)
    tot += scaled("gpu__time_duration.sum", r[c("gpu__time_duration.sum")])
    print(f"{lab:28s} {name:22s} {r[c('launch__grid_size')]:>5s} " + " ".join(f"{v:>10s}" for v in vals))
print(f"# total {tot:.3f} ms over {len(rows) - 2} launches")
