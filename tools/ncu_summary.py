"""Summarise an .ncu-rep (read on the CPU box): per kernel time, instructions, pipes, DRAM bytes, stall mix.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [max_rows]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
maxrows = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
short = {"gpu__time_duration.sum": "us", "smsp__inst_executed.sum": "inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
         "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma%",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu%",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%", "launch__registers_per_thread": "regs",
         "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram%",
         "lts__t_sectors.sum": "l2_sectors", "launch__grid_size": "grid", "launch__block_size": "block"}
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
seen = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > 2 or sum(seen.values()) > maxrows * 2:
        continue
    out = [name[:44]]
    for k, s in short.items():
        if k in idx:
            v = r[idx[k]]
            try:
                f = float(v.replace(",", "")); v = ("%.4g" % f)
            except ValueError:
                pass
            out.append(f"{s}={v}")
    st = sorted(((float(r[idx[n]]), n) for n in stalls if r[idx[n]] not in ("", "n/a")), reverse=True)[:5]
    out.append("stalls: " + " ".join("%s=%.2f" % (n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for v, n in st))
    print("  ".join(out))
