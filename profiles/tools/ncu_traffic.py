#!/usr/bin/env python
"""DRAM traffic per launch of the dominant step kernel, from ncu captures of STEADY-STATE launches:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none \
        --clock-control none -k regex:'k_fused|k_shard' -s 40 -c 6 --csv --log-file gpurun_out/traffic_<wl>.csv \
        python bench.py --workload <wl> --steps 60 --no-cpu --no-c5 --no-rollout

(no cache flush between the replayed launches, 40 launches skipped: the L2 holds what it holds in a real rollout).
This script averages the captured launches per workload and writes profiles/r2_traffic.json, stamped with the hash
of the CUDA sources the capture was taken on -- bench.py prints `traffic_stale: true` when the kernels have changed
since.

    python profiles/tools/ncu_traffic.py c4=gpurun_out/traffic_c4.csv c3=... c5=...
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bench import source_hash  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def parse(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = next(r for r in rows if "Metric Name" in r)
    iid, ik, im, iu, iv = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = {}
    for r in rows:
        if r is hdr or not r[iid].isdigit():
            continue
        d = launches.setdefault(int(r[iid]), {"kernel": r[ik]})
        d[r[im]] = float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
    return list(launches.values())


def main():
    out = {"source_hash": source_hash(),
           "how": "ncu --cache-control none --clock-control none, 40 launches skipped, mean over the captured launches "
                  "(profiles/tools/ncu_traffic.py)"}
    for arg in sys.argv[1:]:
        wl, path = arg.split("=", 1)
        steps = 1
        if ":" in path:            # wl=file.csv:64 -- every captured launch advances 64 steps (in-kernel step loop)
            path, k = path.rsplit(":", 1)
            steps = int(k)
        ls = parse(path)
        # the dominant kernel = the one with the largest total duration among the captured launches
        by = {}
        for d in ls:
            by.setdefault(d["kernel"], []).append(d)
        name, group = max(by.items(), key=lambda kv: sum(x.get("gpu__time_duration.sum", 0.0) for x in kv[1]))
        n = len(group)
        out[wl] = {"kernel": name.split("(")[0], "launches": n, "steps_per_launch": steps,
                   "dram_bytes_read": sum(x["dram__bytes_read.sum"] for x in group) / n,
                   "dram_bytes_write": sum(x["dram__bytes_write.sum"] for x in group) / n,
                   "gpu_time_us_under_ncu": sum(x["gpu__time_duration.sum"] for x in group) / n,
                   "per_launch": [{"read": x["dram__bytes_read.sum"], "write": x["dram__bytes_write.sum"],
                                   "us": x["gpu__time_duration.sum"]} for x in group]}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "per_launch"}) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
