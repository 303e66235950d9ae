"""
Turn the artefacts of tools/gpu_profile.sh (gpurun_out/) into the text summaries kept under profiles/.

    python tools/summarize_profile.py <tag> [<out-name> [<config>]]

Reads gpurun_out/post_<tag>.ncu-rep (one `ncu --set full` capture of post_kernel),
gpurun_out/launches_<tag>.csv (the `--metrics gpu__time_duration.sum` launch list of a short bench
run) and gpurun_out/bench_<tag>.json (the bench line of the same build, NOT taken under ncu).
"""
import csv
import json
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
out_name = sys.argv[2] if len(sys.argv) > 2 else tag
G = ROOT / "gpurun_out"
out = []

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]

rep = G / f"post_{tag}.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out.append(f"== ncu --set full --clock-control none, kernel post_kernel, capture post_{tag}.ncu-rep ==")
    for r in rows[2:3]:
        out.append(f"kernel: {r[hdr.index('Kernel Name')]}")
        for m in METRICS:
            if m in hdr:
                out.append(f"  {m:72s} {r[hdr.index(m)]} {units[hdr.index(m)]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        out.append("  warps stalled per issue-active cycle: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
        # measured DRAM traffic of this launch -> profiles/traffic.json (bench.py reports it as roofline.traffic)
        try:
            def _bytes(metric):
                v, u = float(r[hdr.index(metric)].replace(",", "")), units[hdr.index(metric)].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

            traffic = _bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum")
            # (profiles/traffic.json itself is written by tools/measure_traffic.py, stamped with the hash of
            #  the kernel sources; this summary only states the figure of this capture)
            out.append(f"  DRAM traffic of the launch (read + write): {traffic / 1e6:.1f} MB")
        except (ValueError, KeyError) as e:
            out.append(f"  (traffic not recorded: {e})")
    cs = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv", "--print-source", "cuda,sass"],
                        capture_output=True, text=True).stdout
    rows = list(csv.reader(cs.splitlines()))
    src = (ROOT / "genesis_forge_b200/csrc/post_kernel.cuh").read_text().splitlines()
    marks = [(i, m.group(1)) for i, l in enumerate(src, 1)
             if (m := re.match(r"\s*// (slab loads|entity:|contacts:|terminations|rewards|command resample|in-library part|per-env outputs|slab outputs|slab partials|observations:)", l))]

    def region(f, line):
        if f != "post_kernel.cuh":
            return f
        name = "prologue"
        for ln, nm in marks:
            if line >= ln:
                name = nm
        return name

    cur, hdr2, agg = None, None, defaultdict(lambda: [0, 0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr2 = r
        elif r[0].isdigit() and hdr2:
            try:
                a = agg[region(cur, int(r[0]))]
                a[0] += int(r[hdr2.index("Instructions Executed")])
                a[1] += int(r[hdr2.index("# Samples")])
            except ValueError:
                pass
    tot, ts = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
    out.append("  warp-instructions and stall samples by kernel region (source-correlated):")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]:
        out.append(f"    {k:24s} {v[0] / tot * 100:5.1f}% of instructions   {v[1] / ts * 100:5.1f}% of stall samples")

launch = G / f"launches_{tag}.csv"
if launch.exists():
    rows = [r for r in csv.reader(open(launch)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values()) or 1
    out.append("")
    out.append(f"== launch list (ncu --metrics gpu__time_duration.sum --clock-control none), launches_{tag}.csv ==")
    out.append("   per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))[:12]:
        out.append(f"  {sum(v) / tot * 100:5.1f}%  n={len(v):4d}  avg {sum(v) / len(v) / 1000:8.1f} us  {k[:90]}")

bench = G / f"bench_{tag}.json"
if bench.exists() and bench.read_text().strip():
    line = json.loads(bench.read_text().strip().splitlines()[-1])
    out.append("")
    out.append(f"== bench line of the same build (not under a profiler), bench_{tag}.json ==")
    out.append(json.dumps(line, indent=1))

(ROOT / "profiles").mkdir(exist_ok=True)
(ROOT / "profiles" / f"{out_name}.txt").write_text("\n".join(out) + "\n")
print("\n".join(out[:40]))
