"""Flatten ncu reports into one small CSV (one row per profiled launch) for profiles/.
    python tools/ncu_summary.py out.csv rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv, io, subprocess, sys
COLS = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
        ("gpu__time_duration.sum", "duration"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct_elapsed"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_active"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct_active"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("sm__cycles_elapsed.avg.per_second", "sm_ghz")]
out = csv.writer(open(sys.argv[1], "w", newline=""))
out.writerow(["report"] + [c[1] for c in COLS])
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        line = [rep.split("/")[-1]]
        for key, _ in COLS:
            i = ix.get(key)
            line.append("" if i is None else (r[i] + (" " + units[i] if units[i] and key not in ("Kernel Name",) else "")))
        out.writerow(line)
