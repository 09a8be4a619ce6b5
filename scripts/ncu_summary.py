"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    # L2 -> SM traffic and local-memory (spill) instructions: what the round-1 review asked to see next to lts__throughput
    "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    lines = [f"# ncu summary of {rep}", ""]
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"## kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for h, u, v in zip(hdr, units, r):
            if any(h == k or h.endswith("." + k) or h.endswith(k) for k in KEYS):
                lines.append(f"{h} [{u}] = {v}")
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    if len(src) > 2:
        h = src[1]
        ix = {k: i for i, k in enumerate(h)}
        # (a report with several kernels repeats the header rows per kernel: keep the numeric rows, all kernels pooled)
        data = [r for r in src[2:] if len(r) == len(h) and (r[ix["# Samples"]] or "0").isdigit()]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        stall = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        agg = collections.Counter()
        byop = collections.Counter()
        for r in data:
            s = r[ix["Source"]].split()
            op = (s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?"))
            byop[op] += int(r[ix["# Samples"]] or 0)
            for c in stall:
                agg[c] += int(r[ix[c]] or 0)
        lines.append(f"## warp-state samples: {tot} over {len(data)} SASS instructions")
        lines.append("stall reasons: " + ", ".join(f"{c[6:]}={100 * v / tot:.1f}%" for c, v in agg.most_common(8)))
        lines.append("top opcodes by samples: " + ", ".join(f"{o}={100 * v / tot:.1f}%" for o, v in byop.most_common(12)))
        top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:15]
        lines.append("hottest instructions:")
        for r in top:
            lines.append(f"  {100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:90]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
