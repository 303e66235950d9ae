"""
Per-source-line summary of an `ncu --set full --import-source on` capture: joins the SASS page of the
report (instructions executed, stall samples per instruction) with the line table of the .so the kernel
came from (nvdisasm -g), and prints the lines / line ranges that hold the samples.

    python tools/profile_by_line.py <report.ncu-rep> <kernel .so> [kernel-name-regex] [top N]
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path


def line_table(so: Path, kernel_re: str):
    """[(sass text, file, line)] in instruction order for the first kernel whose name matches."""
    tmp = Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(so)], cwd=tmp, capture_output=True, text=True)
    for cubin in sorted(tmp.glob("*.cubin")):
        out = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
        table, cur_file, cur_line, active = [], "?", 0, False
        for ln in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                active = re.search(kernel_re, m.group(1)) is not None and not table
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_file, cur_line = m.group(1), int(m.group(2))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and active:
                table.append((m.group(2).strip(), Path(cur_file).name, cur_line))
        if table:
            return table
    return []


def main():
    rep, so = Path(sys.argv[1]).resolve(), Path(sys.argv[2]).resolve()
    kernel_re = sys.argv[3] if len(sys.argv) > 3 else "post_kernel"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[1], rows[2:]
    col = {k: hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed", "stall_barrier", "stall_long_sb",
                                    "stall_short_sb", "stall_wait", "stall_no_inst", "stall_branch_resolving", "stall_lg", "stall_mio")}
    table = line_table(so, kernel_re)
    if len(table) != len(data):
        print(f"warning: {len(data)} instructions in the report, {len(table)} in {so.name}: line mapping by position may be off")
    agg = defaultdict(lambda: defaultdict(int))
    tot_s = tot_i = 0
    for k, r in enumerate(data):
        f, line = (table[k][1], table[k][2]) if k < len(table) else ("?", 0)
        a = agg[(f, line)]
        a["samples"] += int(r[col["# Samples"]])
        a["instr"] += int(r[col["Instructions Executed"]])
        for key in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_no_inst", "stall_lg", "stall_mio"):
            a[key] += int(r[col[key]])
        tot_s += int(r[col["# Samples"]])
        tot_i += int(r[col["Instructions Executed"]])
    print(f"{rep.name}: {tot_s} samples, {tot_i} warp instructions, {len(data)} SASS instructions")
    print(f"{'file:line':34s} {'samples%':>8s} {'instr%':>7s}  barrier long_sb short_sb wait no_inst lg mio")
    for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        print(f"{f + ':' + str(line):34s} {100 * a['samples'] / tot_s:8.1f} {100 * a['instr'] / tot_i:7.1f}  "
              f"{a['stall_barrier']:7d} {a['stall_long_sb']:7d} {a['stall_short_sb']:8d} {a['stall_wait']:4d} {a['stall_no_inst']:7d} "
              f"{a['stall_lg']:3d} {a['stall_mio']:3d}")
    # coarse view: 50-line buckets of post_kernel.cuh
    buckets = defaultdict(lambda: defaultdict(int))
    for (f, line), a in agg.items():
        b = (f, line // 50 * 50)
        for k2, v in a.items():
            buckets[b][k2] += v
    print("\n50-line buckets:")
    for (f, b), a in sorted(buckets.items()):
        if a["samples"] * 200 < tot_s:
            continue
        print(f"{f}:{b:4d}-{b + 49:<4d} samples {100 * a['samples'] / tot_s:5.1f}%  instr {100 * a['instr'] / tot_i:5.1f}%  barrier {a['stall_barrier']:6d} "
              f"long_sb {a['stall_long_sb']:6d} short_sb {a['stall_short_sb']:6d}")


if __name__ == "__main__":
    main()
