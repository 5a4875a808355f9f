#!/usr/bin/env python
"""SASS instruction-count table of the tcgen05 / TMA kernels in libgtb200.so (the evidence the profiling
guide asks for: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async).

    python profiles/tools/sass_table.py > profiles/r2_sass_counts.md
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
LIB = ROOT / "gnn_tracking_b200" / "csrc" / "libgtb200.so"
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMALDG.2D.GATHER4", "UTMASTG", "UTMASTG.2D.SCATTER4", "UBLKCP",
             "SYNCS", "LDGSTS", "REDG", "FFMA2", "FADD2", "HMMA"]
KERNELS = re.compile(r"in_edge_ws_kernel|fused_mlp_tc_kernel|in_node_ws_kernel|edge_encoder_ws_kernel|ec_head_ws_kernel|rows_atb_tc_kernel")


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    table: "OrderedDict[str, Counter]" = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = name if KERNELS.search(name) else None
            if cur:
                table[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        table[cur]["total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                table[cur][mn] += 1
        if op.startswith("UTMALDG.2D.GATHER4"):
            table[cur]["UTMALDG.2D.GATHER4"] += 0  # counted by the prefix rule above
    print("# SASS instruction counts (libgtb200.so, sm_100a) -- `python profiles/tools/sass_table.py`\n")
    print("| kernel | total | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|" + "---|" * len(MNEMONICS))
    for k, c in table.items():
        short = re.sub(r"\(.*", "", k).replace("void gtb::", "")
        print(f"| `{short}` | {c['total']} | " + " | ".join(str(c[m]) for m in MNEMONICS) + " |")
    print("\nTemplate arguments: `in_edge_ws_kernel<BF16, RELU_E, PROF>`, `fused_mlp_tc_kernel<W64, PROF>`.")
    print("`UTMALDG` / `UTMASTG` include their `.2D.GATHER4` / `.2D.SCATTER4` forms (listed again in their own columns).")


if __name__ == "__main__":
    sys.exit(main())
