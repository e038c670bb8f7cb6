"""Per-CUDA-source-line roll-up of an .ncu-rep captured with `--set full --import-source on` (kernels built -lineinfo).

usage: python tools/ncu_source.py report.ncu-rep out.txt [top_n]

For every kernel in the report: warp instructions executed and warp-stall samples summed over the SASS of each source
line, as a share of the kernel, with the two dominant stall reasons of the line, top lines first.  Small enough to
commit under profiles/.
"""
import csv, io, subprocess, sys
from collections import defaultdict

rep, out = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout

blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "File Path":
        cur = {"file": row[1], "fn": "?", "hdr": None, "rows": []}
        blocks.append(cur)
    elif row[0] == "Function Name" and cur is not None:
        cur["fn"] = row[1]
    elif row[0] == "Line No" and cur is not None:
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)

with open(out, "w") as f:
    per_fn = defaultdict(list)
    for b in blocks:
        per_fn[b["fn"]].append(b)
    for fn, bs in per_fn.items():
        inst, stall, text, reasons = defaultdict(float), defaultdict(float), {}, defaultdict(lambda: defaultdict(float))
        for b in bs:
            h = b["hdr"]
            i_line, i_src = 0, 1
            i_inst = h.index("Instructions Executed")
            i_stall = h.index("Warp Stall Sampling (All Samples)")
            st_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
            short = b["file"].split("/")[-1]
            for r in b["rows"]:
                try:
                    key = (short, int(r[i_line]))
                except ValueError:
                    continue
                text[key] = r[i_src].strip()
                try:
                    inst[key] += float(r[i_inst] or 0)
                    stall[key] += float(r[i_stall] or 0)
                    for i, c in st_cols:
                        reasons[key][c[6:]] += float(r[i] or 0)
                except ValueError:
                    pass
        ti, ts = sum(inst.values()) or 1.0, sum(stall.values()) or 1.0
        f.write("## %s\n   warp instructions %.0f, stall samples %.0f\n" % (fn[:140], ti, ts))
        f.write("   %-24s %7s %7s  %-28s %s\n" % ("file:line", "inst%", "stall%", "top stall reasons", "source"))
        for key in sorted(inst, key=lambda k: -(inst[k] / ti + stall[k] / ts))[:top_n]:
            rs = sorted(reasons[key].items(), key=lambda kv: -kv[1])[:2]
            f.write("   %-24s %6.2f%% %6.2f%%  %-28s %s\n" % ("%s:%d" % key, 100 * inst[key] / ti, 100 * stall[key] / ts,
                                                          ",".join("%s %.0f" % kv for kv in rs if kv[1] > 0), text[key][:100]))
        f.write("\n")
print("wrote", out)
