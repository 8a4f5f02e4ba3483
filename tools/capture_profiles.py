"""Round-2 profile capture, run ON the GPU box (one gpurun call):

    python tools/capture_profiles.py [case ...]          # default: every case below

For every case: the command runs once without ncu (must exit 0), then under
`ncu --set full --clock-control none --import-source on` with a kernel-name filter.  Outputs (gpurun_out/):
    r02_<case>.ncu-rep, r02_<case>.txt (tools/ncu_summary.py), and r02_ncu.json = one record per kernel family with
    dram bytes / pixel, instructions / pixel, pipe utilisation, active lanes per instruction, stamped with
    the fingerprint of the kernel sources (nblic_image_compression_b200.build.source_fingerprint), the git HEAD the
    snapshot was taken from (env NBLIC_GIT_HEAD: the box has no .git) and the launch shape.
bench.py quotes these figures only when the fingerprint equals that of the build it runs.
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nblic_image_compression_b200.build import source_fingerprint

OUT = os.path.join(ROOT, "gpurun_out")
PY = sys.executable
# case -> (profile_case args, ncu -k regex, launches to skip, launches to capture, {record key: kernel-name substring})
CASES = {
    # the bench shape: 1024 x 1024 images, a full residency of the encoder; "warp4" = the packed decoder the bench's 10 000 streams run on
    "e1": (["3552", "1024", "1024", "1", "0", "warp4"], "regex:nblic_kernel|subwarp_decode", 0, 2, {"e1_encode_lossless": "ENC", "e1_decode": "DEC"}),
    "e1n2": (["3552", "512", "512", "1", "2"], "regex:nblic_kernel", 0, 2, {"e1n2_encode": "ENC", "e1n2_decode": "DEC"}),
    "e0": (["3552", "1024", "1024", "0", "0"], "regex:coop_q_kernel", 0, 2, {"e0_encode": "ENC", "e0_decode": "DEC"}),
    "e2": (["2368", "256", "256", "2", "0"], "regex:nblic_kernel", 0, 2, {"e2_encode": "ENC", "e2_decode": "DEC"}),
    "e3": (["2368", "128", "256", "3", "0"], "regex:nblic_kernel", 0, 2, {"e3_encode": "ENC", "e3_decode": "DEC"}),
}
METRICS = {
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "ipc_per_sm": "sm__inst_executed.avg.per_cycle_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "active_lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
}


def num(x):
    """A metric as a float; None for missing or not-a-number values (ncu reports nan for counters it could not collect)."""
    try:
        v = float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return None
    return v if v == v and abs(v) != float("inf") else None


def main():
    os.makedirs(OUT, exist_ok=True)
    wanted = sys.argv[1:] or list(CASES)
    path = os.path.join(OUT, "r02_ncu.json")
    rec = {"csrc_fingerprint": source_fingerprint(), "git_head": os.environ.get("NBLIC_GIT_HEAD", "unknown"),
           "how": "ncu --set full --clock-control none --import-source on over tools/profile_case.py; per-launch figures", "kernels": {}}
    if os.path.exists(path):
        old = json.load(open(path))
        if old.get("csrc_fingerprint") == rec["csrc_fingerprint"]:
            rec["kernels"] = old.get("kernels", {})
    for case in wanted:
        args, kfilter, skip, count, keys = CASES[case]
        cmd = [PY, os.path.join(ROOT, "tools", "profile_case.py"), *args]
        plain = subprocess.run(cmd, capture_output=True, text=True)
        open(os.path.join(OUT, f"r02_{case}_plain.log"), "w").write(plain.stdout + plain.stderr)
        if plain.returncode != 0:
            print(f"{case}: plain run failed, not profiled"); continue
        rep = os.path.join(OUT, f"r02_{case}")
        r = subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", kfilter, "-s", str(skip), "-c", str(count),
                            "-f", "-o", rep, *cmd], capture_output=True, text=True)
        open(os.path.join(OUT, f"r02_{case}_ncu.log"), "w").write(r.stdout + r.stderr)
        if r.returncode != 0 or not os.path.exists(rep + ".ncu-rep"):
            print(f"{case}: ncu failed rc={r.returncode}"); continue
        n, h, w = int(args[0]), int(args[1]), int(args[2])
        pixels = n * h * w
        raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = rows[0]
        launches = rows[2:]
        for idx, (key, _tag) in enumerate(keys.items()):  # profile_case runs encode then decode: launch order = key order
            if idx >= len(launches):
                continue
            row = dict(zip(hdr, launches[idx]))
            k = {"kernel": row.get("Kernel Name"), "shape": f"{n} x {h}x{w} -e{args[3]} -n{args[4]}", "grid": row.get("launch__grid_size"),
                 "registers": row.get("launch__registers_per_thread"), "time_ms_under_ncu": (num(row.get("gpu__time_duration.sum")) or 0) / 1e6,  # base unit: ns
                 "pixels": pixels}
            inst = num(row.get("smsp__inst_executed.sum"))
            rd, wr = num(row.get("dram__bytes_read.sum")), num(row.get("dram__bytes_write.sum"))
            k["inst_per_pixel"] = round(inst / pixels, 2) if inst else None
            k["dram_bytes_per_pixel"] = round((rd + wr) / pixels, 3) if rd is not None and wr is not None else None
            k["pipes"] = {name: num(row.get(metric)) for name, metric in METRICS.items()}
            rec["kernels"][key] = k
        summ = subprocess.run([PY, os.path.join(ROOT, "tools", "ncu_summary.py"), rep + ".ncu-rep", str(pixels), "", "40"], capture_output=True, text=True)
        head = f"# git {rec['git_head']}  csrc {rec['csrc_fingerprint']}  case {case}: profile_case.py {' '.join(args)}\n"
        open(os.path.join(OUT, f"r02_{case}.txt"), "w").write(head + summ.stdout + summ.stderr[-2000:])
        json.dump(rec, open(path, "w"), indent=1)
        print(f"{case}: captured {[ (k, rec['kernels'][k]['inst_per_pixel']) for k in keys if k in rec['kernels']]}")


if __name__ == "__main__":
    main()
