#!/usr/bin/env python
"""DRAM traffic per launch of the five passes of the CH-3D-512 substep (roofline.traffic of bench.py).

ncu cannot run inside a timed bench, so this script captures one steady-state AB2 substep of tools/pass_times.py under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` (one replay pass, clocks untouched) and writes
profiles/traffic.json: bytes per launch for every pass, the kernel names, and the hash of the library build they were
measured on.  bench.py reports a traffic figure only when that hash equals the build it is timing.
Run on the GPU box in the same session as the bench:  python tools/measure_traffic.py
"""
import csv
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_bytes, lib_sha256  # noqa: E402


def main():
    n = 512
    out_csv = os.path.join(ROOT, "gpurun_out", "traffic_ncu.csv")
    os.makedirs(os.path.dirname(out_csv), exist_ok=True)
    # pass_times.py: 1 AB1 + 3 AB2 substeps before anything is timed = 20 launches of this library's kernels (the filter
    # keeps torch's own kernels - checksums, fills - out of the count); take the next substep
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "--print-units", "base", "-k", "regex:^k_", "-s", "20", "-c", "5", "--csv", "--log-file", out_csv,
           sys.executable, os.path.join(ROOT, "tools", "pass_times.py"), str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        raise SystemExit("ncu failed")
    rows = []
    with open(out_csv) as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for row in csv.DictReader(lines):
        rows.append(row)
    per = {}
    order = []
    for row in rows:
        k = row["ID"]
        if k not in per:
            per[k] = {"kernel": row["Kernel Name"]}
            order.append(k)
        per[k][row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
    _, _, alg = algorithmic_bytes(n, 1)
    names = list(alg.keys())
    assert len(order) == len(names), (len(order), [per[k]["kernel"] for k in order])
    expect = ["k_zfwd", "k_strided", "k_fused", "k_strided", "k_zinv"]
    assert all(e in per[k]["kernel"] for e, k in zip(expect, order)), [per[k]["kernel"] for k in order]
    passes, kernels, detail = {}, {}, {}
    for nm, k in zip(names, order):
        d = per[k]
        b = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        passes[nm] = b
        kernels[nm] = d["kernel"]
        detail[nm] = {"read": d["dram__bytes_read.sum"], "write": d["dram__bytes_write.sum"], "algorithmic": alg[nm],
                      "traffic_over_algorithmic": round(b / alg[nm], 4), "ncu_duration_ns": d.get("gpu__time_duration.sum")}
    out = {"passes": passes, "kernels": kernels, "detail": detail, "lib_sha256": lib_sha256(), "grid": [n, n, n],
           "_source": "tools/measure_traffic.py: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one "
                      "steady-state AB2 substep, bytes per launch; " + time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
