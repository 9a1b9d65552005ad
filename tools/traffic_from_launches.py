"""Derive profiles/traffic_*.json (bench.py's roofline.traffic) and a per-family share table from an ncu launch list.
    python tools/traffic_from_launches.py profiles/launches_r1h_step.csv profiles/traffic_r1h.json"""
import csv, json, sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
L = {}
for r in rows[1:]:
    d = L.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]], "stream": r[ix["Stream"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    name = r[ix["Metric Name"]]
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}[unit]
    elif name.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    d[name] = v
TENSOR = ("conv3x3_halo_kernel", "wgrad3x3_halo_kernel", "igemm_kernel")
fam = {}
tot = 0.0
for d in L.values():
    k = d["kernel"]
    name = next((t for t in TENSOR if t in k), None) or k.split("(")[0].split("::")[-1].split("<")[0]
    f = fam.setdefault(name, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    f["launches"] += 1
    f["us"] += d["gpu__time_duration.sum"]
    f["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot += d["gpu__time_duration.sum"]
tl = sum(fam[t]["launches"] for t in TENSOR if t in fam)
tb = sum(fam[t]["dram_bytes"] for t in TENSOR if t in fam)
tu = sum(fam[t]["us"] for t in TENSOR if t in fam)
out = {"source": f"{src} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,...; one CubeNET-64 batch-2 training step, serialised)",
       "launches": len(L), "step_us_serialised": tot,
       "tensor_family_launches": tl, "tensor_family_dram_bytes_per_step": tb,
       "tensor_family_dram_bytes_per_launch": tb / max(tl, 1), "tensor_family_share_of_step_ncu": tu / tot,
       "families": {k: {"launches": v["launches"], "us": round(v["us"], 1), "share": round(v["us"] / tot, 4),
                        "dram_mb": round(v["dram_bytes"] / 1e6, 1)} for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["us"])}}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
