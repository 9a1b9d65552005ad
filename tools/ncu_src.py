"""Summarise an ncu report: key raw metrics + hottest SASS lines by stall samples (needs ncu on PATH)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
vals = rows[2 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "lts__t_sectors_srcunit_tex_op_read.sum", "launch__registers_per_thread", "sm__throughput.avg.pct"]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k) for k in keys) and not h.endswith((".max", ".min")) and ".max." not in h and ".min." not in h and ".sum.p" not in h:
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # k-th profiled launch of the report
hdr = rows[1]
data, sec = [], -1
for r in rows:                          # sections start with a "Kernel Name" row, followed by the header row
    if r and r[0] == "Kernel Name":
        sec += 1
        continue
    if sec == which and len(r) == len(hdr) and r[0] != "Address":
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot)
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(enumerate(data), key=lambda kv: -int(kv[1][ix["# Samples"]]))[:n]
for i, r in sorted(top):
    st = {s: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0}
    print(i, r[ix["Source"]][:72], r[ix["# Samples"]], r[ix["Instructions Executed"]], sorted(st.items(), key=lambda kv: -kv[1])[:3])
