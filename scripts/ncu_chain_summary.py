"""Summarise an `ncu --set full` report of the chain launches of one training step (read with `ncu -i REP --page raw --csv`):
per launch time, tensor-pipe activity, DRAM bytes / throughput; writes the CSV summary and profiles/r2_chain_traffic.json
(dram bytes per launch, what bench.py reports as roofline.traffic)."""
import csv
import json
import subprocess
import sys

rep, out_csv, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
cols = {"Kernel Name": "kernel", "gpu__time_duration.sum": "time_ms",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
        "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "launch__registers_per_thread": "regs",
        "launch__grid_size": "grid", "launch__block_size": "block"}
units = rows[1]
idx = {h: i for i, h in enumerate(H)}
out = []
for r in rows[2:]:
    if len(r) < len(H):
        continue
    d = {}
    for h, name in cols.items():
        if h in idx:
            v = r[idx[h]]
            u = units[idx[h]]
            if name in ("dram_read", "dram_write"):
                f = float(v.replace(",", ""))
                f *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
                v = f
            elif name == "time_ms":
                f = float(v.replace(",", ""))
                f *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(u, 1.0)
                v = f
            d[name] = v
    d["kernel"] = d["kernel"].replace("<unnamed>::", "").split("(")[0]
    out.append(d)
with open(out_csv, "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=list(cols.values()))
    w.writeheader()
    for d in out:
        w.writerow(d)
tot = sum(d["dram_read"] + d["dram_write"] for d in out)
js = {"dram_bytes_per_launch": tot / max(len(out), 1), "launches": len(out),
      "per_launch": [{"kernel": d["kernel"], "time_ms": d["time_ms"], "dram_bytes": d["dram_read"] + d["dram_write"],
                      "tensor_pipe_pct": d.get("tensor_pipe_pct")} for d in out],
      "source": f"{out_csv} (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full --clock-control none, the chain launches "
                "of one 8192-ray step of this build)"}
with open(out_json, "w") as f:
    json.dump(js, f, indent=1)
for d in out:
    print(d)
print("mean dram bytes per launch: %.3f GB" % (js["dram_bytes_per_launch"] / 1e9))
