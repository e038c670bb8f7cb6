"""Turn an .ncu-rep into the small CSV/markdown summaries committed under profiles/ (run where ncu is installed)."""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum"]
with open(out, "w") as f:
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write("## %s\n" % d.get("Kernel Name", "?")[:110])
        for k in keys[1:]:
            if k in d:
                f.write("%-66s %s %s\n" % (k, d[k], units[hdr.index(k)]))
        st = sorted(((float(d[h].replace(",", "")) if d[h] not in ("", "n/a") else 0, h) for h in hdr
                     if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h), reverse=True)[:6]
        f.write("top stalls (warps per issue): " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in st) + "\n\n")
print("wrote", out)
