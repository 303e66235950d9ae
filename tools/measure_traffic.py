#!/usr/bin/env python
"""
DRAM traffic of ONE post_kernel launch per workload, measured with ncu, written to profiles/traffic.json
together with the hash of the kernel sources (bench.py refuses entries whose hash is stale).

    python tools/measure_traffic.py [config:num_envs ...]        (on the GPU box, under gpurun)

For each workload: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:post_kernel` around
a short bench.py run (2 launches skipped, the next one measured).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = ["command_direction:1048576", "contacts:65536", "contacts:1048576", "gait_trainer:65536",
           "rough_terrain:262144", "berkeley_humanoid:65536", "berkeley_humanoid:262144", "berkeley_humanoid:1048576"]


def measure(config: str, n: int):
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", "regex:post_kernel", "-s", "6", "-c", "1", "--csv",
           sys.executable, os.path.join(ROOT, "bench.py"), "--config", config, "--num-envs", str(n), "--steps", "3",
           "--warmup", "3", "--no-sweep", "--no-cpu", "--no-e2e", "--no-configs"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600).stdout
    start = out.find('"ID"')
    if start < 0:
        return None
    rows = list(csv.DictReader(io.StringIO(out[start:])))
    vals = {}
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "msecond": 1e3,
                 "nsecond": 1e-3}.get(unit, 1)
        vals[r["Metric Name"]] = v * scale
    if "dram__bytes_read.sum" not in vals:
        return None
    return {"dram_bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
            "dram_read": vals["dram__bytes_read.sum"], "dram_write": vals["dram__bytes_write.sum"],
            "kernel_us_under_ncu": vals.get("gpu__time_duration.sum"), "kernel": rows[0]["Kernel Name"][:60]}


def main():
    from genesis_forge_b200 import spec

    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        table = json.load(open(path))
    except (OSError, ValueError):
        table = {}
    for item in sys.argv[1:] or DEFAULT:
        config, n = item.split(":")
        m = measure(config, int(n))
        if m is None:
            print(item, "no measurement")
            continue
        m["kernel_source_hash"] = spec.source_hash()
        table[item] = m
        print(item, m)
    out = os.path.join(ROOT, "gpurun_out", "traffic.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(table, open(out, "w"), indent=1, sort_keys=True)  # (copy to profiles/traffic.json and commit)
    print("written", out)


if __name__ == "__main__":
    main()
