#!/usr/bin/env python
"""Folds an .ncu-rep (``ncu --set full --import-source on``) into the two text files kept under profiles/:
``<out>_summary.csv`` (time, DRAM bytes, pipes, occupancy, stall ratios of every captured launch) and
``<out>_top_stalls.txt`` (the 30 instructions with the most stall samples + the totals per stall reason).

    python profiles/tools/ncu_summary.py gpurun_out/r2_edge_ws_f32.ncu-rep profiles/r2_ncu_edge_ws_f32
"""
import csv
import io
import subprocess
import sys

KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_tma_ld.sum", "smsp__sass_inst_executed_op_tma_st.sum",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_elapsed")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"))]
    with open(out + "_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch {k}" for k in range(len(rows) - 2)])
        for i in cols:
            w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])
    src = page(rep, "source")
    h = src[1]
    ix = {n: i for i, n in enumerate(h)}
    data = [r for r in src[2:] if len(r) == len(h)]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    with open(out + "_top_stalls.txt", "w") as f:
        f.write(f"{src[0][1]}\n{tot} warp-state samples over {len(data)} SASS instructions\n\n")
        f.write("share of samples per stall reason: " + ", ".join(
            f"{n[6:]} {100 * sum(int(r[ix[n]] or 0) for r in data) / tot:.1f}%" for n in stalls) + "\n\n")
        f.write("top instructions by samples (share, instruction, dominant reason)\n")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:30]:
            s = int(r[ix["# Samples"]])
            dom = max(stalls, key=lambda n: int(r[ix[n]] or 0))
            f.write(f"{100 * s / tot:5.1f}%  {r[ix['Source']].strip()[:80]:80s} {dom[6:]}\n")


if __name__ == "__main__":
    main()
