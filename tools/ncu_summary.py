"""Summarise an .ncu-rep: headline metrics + per-source-line instruction counts and stall samples.
usage: python tools/ncu_summary.py REPORT.ncu-rep PIXELS [kernel-substring] [top-N]"""
import csv, io, subprocess, sys
rep, pixels = sys.argv[1], float(sys.argv[2])
kfilter = sys.argv[3] if len(sys.argv) > 3 else ""
norm = lambda t: (t or "").replace("(int)", "").replace("(bool)", "").replace("true", "1").replace("false", "0").replace(" ", "")
kfilter = norm(kfilter)
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rows[2:]:
    if kfilter not in norm(r[hdr.index("Kernel Name")]):
        continue
    print("=" * 100)
    for w in want:
        if w in hdr:
            print(f"{w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
    try:
        print("instructions / pixel =", float(r[hdr.index("smsp__inst_executed.sum")]) / pixels)
    except ValueError:
        pass
    for i, hname in enumerate(hdr):
        if "issue_stalled" in hname and hname.endswith("per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.05:
                print(f"   stall {hname.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
out, cur_file, cur_fn, h2 = [], None, None, None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": h2 = r; continue
    if h2 and r[0].isdigit() and len(r) > 8 and kfilter in norm(cur_fn):
        d = dict(zip(h2, r))
        try:
            out.append((cur_file.split("/")[-1], int(r[0]), int(d["Instructions Executed"]), int(d["# Samples"]), r[1].strip()[:100]))
        except (ValueError, KeyError):
            pass
tot_i, tot_s = sum(o[2] for o in out) or 1, sum(o[3] for o in out) or 1
print(f"--- per source line (top {topn} by stall samples); total {tot_i / pixels:.1f} inst/pixel attributed")
for o in sorted(out, key=lambda o: -o[3])[:topn]:
    print(f"{o[0]}:{o[1]:4d} inst/px={o[2] / pixels:7.2f} samples%={100 * o[3] / tot_s:5.1f}  {o[4]}")
