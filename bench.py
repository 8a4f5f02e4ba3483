#!/usr/bin/env python
"""bench.py -- NBLIC batch encode+decode throughput on B200 (BASELINE.json metric), one JSON line.

Workload = BASELINE.json configs[4]: B = 10000 synthetic 1024x1024 gray images per GPU (deterministic
generator of SURVEY.md Appendix B, seeds disjoint across ranks), lossless -n0 -e1 (NBLIC), one step =
encode the batch, then decode the streams it produced (10.5 GPixel each way).  The whole 10k-image
config fits one GPU (about 95 GB of HBM with slots and scratch), so N=1 runs it as named; images are
independent, so with N ranks every rank runs its own 10k images and shares nothing ("scaling": "weak",
no collective on the data path; torch.distributed is only the barrier and the max-over-ranks).

    value     (pixels encoded + pixels decoded) / s, inputs resident in HBM, CUDA events on the codec's stream
    e2e       the same through the host-buffer C ABI (nblic_b200_encode_batch / nblic_b200_decode_batch)
              with pinned host buffers: H2D of pixels, D2H of streams, H2D of streams, D2H of pixels
    roofline  dominant kernel = the longer of the two coop_nblic_kernel launches (decode), timed by the
              library's own CUDA events on its stream
    cpu_baseline  the unmodified reference (oracle/_ref/libnblic_ref.so; else the oracle port) on all
              host cores, one process per core, bounded sample of the same workload (rank 0, N=1)

`--impl reference` times only that CPU arm and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 1024
NEAR, EFFORT = 0, 1
METRIC = "nblic_batch_encode_decode_throughput"
UNIT = "MPixel/s"
# contract issue-slot weights per pixel (e0, e1, e2, e3) and the lane-issue peak
SLOTS_PER_PIXEL = {0: 275, 1: 626, 2: 18256, 3: 65376}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference codec, one process per host core
# ------------------------------------------------------------------------------------------------
def _cpu_lib():
    ref = os.path.join(ROOT, "oracle", "_ref", "libnblic_ref.so")
    if os.path.exists(ref):
        return ref, "reference"
    port = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(port):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return port, "port"


_CPU_IMAGES = None  # inherited by forked workers


def _cpu_worker(args):
    idxs, path, kind, near, effort = args
    lib = C.CDLL(path)
    u8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    px = 0
    for i in idxs:
        img = _CPU_IMAGES[i].copy()
        h, w = img.shape
        out = np.zeros(2 * h * w + 65536, dtype=np.uint8)
        dec = np.zeros(h * w, dtype=np.uint8)
        n_, e_, hh, ww = C.c_int(near), C.c_int(effort), C.c_int(), C.c_int()
        if kind == "reference":
            n = lib.NBLICcompress(0, out.ctypes.data_as(u8p), img.ctypes.data_as(u8p), h, w, C.byref(n_), C.byref(e_))
            rc = lib.NBLICdecompress(0, out.ctypes.data_as(u8p), dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww), C.byref(n_), C.byref(e_))
        else:
            lib.oracle_n_decode.argtypes = [u8p, C.c_long, u8p, ip, ip, ip, ip]
            n = lib.oracle_n_encode(img.ctypes.data_as(u8p), h, w, C.byref(n_), C.byref(e_), out.ctypes.data_as(u8p))
            rc = lib.oracle_n_decode(out.ctypes.data_as(u8p), n, dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww), C.byref(n_), C.byref(e_))
        assert n > 0 and rc == 0 and np.array_equal(dec.reshape(h, w), _CPU_IMAGES[i])
        px += 2 * h * w
    return px


def _cpu_hash_worker(args):
    """SHA-256 of the CPU codec's stream for each image index (parity spot check, outside every timed region)."""
    import hashlib
    idxs, path, kind, near, effort = args
    lib = C.CDLL(path)
    u8p = C.POINTER(C.c_uint8)
    out = []
    for i in idxs:
        img = _CPU_IMAGES[i].copy()
        h, w = img.shape
        buf = np.zeros(2 * h * w + 65536, dtype=np.uint8)
        n_, e_ = C.c_int(near), C.c_int(effort)
        if kind == "reference":
            n = lib.NBLICcompress(0, buf.ctypes.data_as(u8p), img.ctypes.data_as(u8p), h, w, C.byref(n_), C.byref(e_))
        else:
            n = lib.oracle_n_encode(img.ctypes.data_as(u8p), h, w, C.byref(n_), C.byref(e_), buf.ctypes.data_as(u8p))
        out.append((i, int(n), hashlib.sha256(buf[:max(n, 0)].tobytes()).hexdigest()))
    return out


def cpu_stream_hashes(images, indices, cores, near, effort):
    import multiprocessing as mp
    global _CPU_IMAGES
    _CPU_IMAGES = images
    path, kind = _cpu_lib()
    jobs = [(list(indices[k::cores]), path, kind, near, effort) for k in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        res = [r for part in pool.map(_cpu_hash_worker, jobs, chunksize=1) for r in part]
    return {i: (n, hsh) for i, n, hsh in res}, kind


def cpu_throughput(images, cores, near, effort):
    """(MPixel/s over encode+decode, seconds) of the CPU codec on `cores` processes over `images`."""
    import multiprocessing as mp
    global _CPU_IMAGES
    _CPU_IMAGES = images
    path, kind = _cpu_lib()
    jobs = [(list(range(k, len(images), cores)), path, kind, near, effort) for k in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [([], path, kind, near, effort)] * cores)  # spin the workers up
        t0 = time.perf_counter()
        done = pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return sum(done) / dt / 1e6, dt, kind


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=int(os.environ.get("NBLIC_BENCH_IMAGES", "10000")), help="images per GPU per step")
    ap.add_argument("--mapping", default=os.environ.get("NBLIC_BENCH_MAPPING", "auto"), choices=["auto", "warp", "lane"])
    ap.add_argument("--cpu-images-per-core", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    workload = f"configs[4]: {args.images} synthetic {H}x{W} gray images per GPU, -n{NEAR} -e{EFFORT} (NBLIC), encode then decode"

    if args.impl == "reference":
        if rank != 0:
            return
        from nblic_image_compression_b200.synth import gen
        per_step = cores * 4
        uniq = [gen(H, W, s) for s in range(min(per_step, 16))]
        images = [uniq[i % len(uniq)] for i in range(per_step)]
        vals, secs, kind = [], [], "reference"
        for it in range(args.warmup + args.steps):
            v, dt, kind = cpu_throughput(images, cores, NEAR, EFFORT)
            if it >= args.warmup:
                vals.append(v); secs.append(dt)
        value = sum(2 * im.size for im in images) * len(secs) / sum(secs) / 1e6
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * sum(secs) / len(secs), 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload, "sample": f"{per_step} images per step ({len(uniq)} distinct seeds)"},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{per_step} images of the workload per step, one process per core, encode+decode each"},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    from nblic_image_compression_b200 import api
    from nblic_image_compression_b200.synth import occluders  # noqa: F401  (numpy PCG64 table for the device generator)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    codec = api.Codec(local_rank, {"auto": api.MAP_AUTO, "warp": api.MAP_WARP, "lane": api.MAP_LANE}[args.mapping])
    B = args.images
    npx = H * W
    dev = torch.device("cuda", local_rank)
    d_pixels = torch.empty(B * npx, dtype=torch.uint8, device=dev)
    for i in range(B):
        codec.synth_device(d_pixels.data_ptr() + i * npx, H, W, rank * B + i)
    pix_off = np.arange(B, dtype=np.uint64) * npx
    hs = np.full(B, H, dtype=np.int32)
    ws = np.full(B, W, dtype=np.int32)
    stream_cap = B * api.stream_bound(H, W)
    d_streams = torch.empty(stream_cap, dtype=torch.uint8, device=dev)
    d_decoded = torch.empty(B * npx, dtype=torch.uint8, device=dev)
    ext = torch.cuda.ExternalStream(codec.stream_handle, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm ----------------------------------------------------------------
    state = {}

    def step_device():
        off, st, rc = codec.encode_device(d_pixels.data_ptr(), pix_off, hs, ws, NEAR, EFFORT, d_streams.data_ptr(), stream_cap)
        state["enc_ms"] = codec.last_coder_ms
        assert rc == 0, st
        st2, rc2 = codec.decode_device(d_streams.data_ptr(), off, d_decoded.data_ptr(), pix_off)
        state["dec_ms"] = codec.last_coder_ms
        assert rc2 == 0, st2
        state["stream_off"] = off

    for _ in range(args.warmup):
        step_device()
    assert torch.equal(d_pixels, d_decoded), "decode(encode(x)) != x"
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = codec.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc_ms, dec_ms = [], []
    e0.record(ext)
    for _ in range(args.steps):
        step_device()
        enc_ms.append(state["enc_ms"]); dec_ms.append(state["dec_ms"])
    e1.record(ext)
    barrier()
    launches = codec.launches - launches0
    dev_s = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    stream_bytes = int(state["stream_off"][-1])
    value = world * B * npx * 2 * args.steps / dev_s / 1e6

    # ---- end-to-end arm: pinned host buffers through the host C ABI --------------------------
    h_pixels = torch.empty(B * npx, dtype=torch.uint8, pin_memory=True)
    h_pixels.copy_(d_pixels)
    h_np = h_pixels.numpy()
    bound = H * W + 8192  # caller-side capacity per stream: 8 bpp + header room (lossless -e1 needs ~3.5 bpp here; a larger stream is reported as overflow)
    h_streams = torch.empty(B * bound, dtype=torch.uint8, pin_memory=True)
    h_decoded = torch.empty(B * npx, dtype=torch.uint8, pin_memory=True)
    images = [h_np[i * npx:(i + 1) * npx].reshape(H, W) for i in range(B)]
    outs = [h_streams.numpy()[i * bound:(i + 1) * bound] for i in range(B)]
    dec_views = [h_decoded.numpy()[i * npx:(i + 1) * npx] for i in range(B)]
    lib = codec.lib
    img_ptrs, out_ptrs, dec_ptrs = api._ptr_array(images), api._ptr_array(outs), api._ptr_array(dec_views)
    hs_c = (C.c_int * B)(*([H] * B)); ws_c = (C.c_int * B)(*([W] * B))
    caps = (C.c_size_t * B)(*([bound] * B)); lens = (C.c_size_t * B)()
    dcaps = (C.c_size_t * B)(*([npx] * B))
    status = (C.c_int * B)()

    def step_e2e():
        rc = lib.nblic_b200_encode_batch(codec.ctx, B, img_ptrs, hs_c, ws_c, NEAR, EFFORT, out_ptrs, caps, lens, None, status)
        assert rc == 0, codec._err()
        rc = lib.nblic_b200_decode_batch(codec.ctx, B, out_ptrs, lens, dec_ptrs, dcaps, None, None, None, None, status)
        assert rc == 0, codec._err()

    step_e2e()  # one untimed pass: first touch of the pinned buffers and of the library's staging allocations
    assert np.array_equal(h_decoded.numpy(), h_np), "e2e decode(encode(x)) != x"
    barrier()
    e0.record(ext)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e1.record(ext)
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_s = max_over_ranks(max(e0.elapsed_time(e1) / 1e3, 0.0))
    clocks = sampler.stop()
    e2e_value = world * B * npx * 2 * args.steps / e2e_s / 1e6
    e2e_stream_bytes = int(sum(lens[i] for i in range(B)))

    # ---- parity check of a random 1 % sample against the CPU codec (checker only, outside the timed regions) ----
    parity = None
    if rank == 0:
        import hashlib
        try:
            sample = sorted(np.random.default_rng(0).choice(B, size=max(1, B // 100), replace=False).tolist())
            ref_hashes, kind = cpu_stream_hashes(images, sample, cores, NEAR, EFFORT)
            bad = [i for i in sample if ref_hashes[i] != (int(lens[i]), hashlib.sha256(bytes(outs[i][: lens[i]])).hexdigest())]
            who = "the unmodified reference" if kind == "reference" else "the oracle port"
            parity = (f"bit-exact vs {who} on {len(sample)} random images of the batch (1 %)" if not bad
                      else f"MISMATCH vs {who} on images {bad[:8]}")
        except Exception as ex:  # pragma: no cover
            parity = f"CPU checker unavailable: {ex}"

    # ---- roofline of the dominant kernel (the longer of the encode / decode coder launches) --------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    enc_s = float(np.mean(enc_ms)) / 1e3
    dec_s = float(np.mean(dec_ms)) / 1e3
    dom_is_dec = dec_s >= enc_s
    dom_s = dec_s if dom_is_dec else enc_s
    # algorithmic bytes per launch: every pixel and every stream byte cross HBM once (read one, write the other)
    alg_bytes = B * npx + stream_bytes
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_mhz_max = float(peaks.get("sm_max_mhz", 1965.0))
    issue_peak = sm_count * 4 * 32 * sm_mhz_max * 1e6  # lane-issue slots / s
    issue_dom = B * npx / dom_s * SLOTS_PER_PIXEL[EFFORT]
    traffic, traffic_note, pipes = None, "no ncu capture on record", None
    try:  # dram bytes per pixel of the same kernel from the committed ncu --set full capture, scaled to this launch
        figures = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        pipes = figures.get("pipes", {}).get("coop_nblic_kernel<0, 2, 1>" if dom_is_dec else "coop_nblic_kernel<0, 0, 1>")
        cap = figures["e1_decode" if dom_is_dec else "e1_encode_lossless"]
        traffic = int(cap["dram_bytes_per_pixel"] * B * npx)
        traffic_note = "dram__bytes_read+write per pixel of the ncu capture (%s) x pixels of this launch; above the algorithmic bytes because the per-stream rank/frequency tables (50 KB x resident streams) exceed L2" % cap["capture"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "coop_nblic_kernel<effort 1, %s>" % ("decode" if dom_is_dec else "lossless encode"),
        "achieved": round(alg_bytes / dom_s / 1e9, 3), "peak": hbm_peak, "unit": "GB/s",
        "frac": round(alg_bytes / dom_s / 1e9 / hbm_peak, 6), "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
        "kernel_ms": round(1e3 * dom_s, 3), "encode_kernel_ms": round(1e3 * enc_s, 3), "decode_kernel_ms": round(1e3 * dec_s, 3),
        "algorithmic_bytes_per_launch": int(alg_bytes),
        "note": "the path is bound by dependent integer issue, not HBM or tensor throughput (SURVEY.md 8(d)): see `issue`",
        "issue": {"unit": "T lane-issue-slots/s", "slots_per_pixel": SLOTS_PER_PIXEL[EFFORT], "achieved": round(issue_dom / 1e12, 4),
                  "peak": round(issue_peak / 1e12, 3), "frac": round(issue_dom / issue_peak, 6),
                  "ncu_pipes_pct_of_peak": pipes,
                  "ncu_note": "pipe utilisation of the same kernel from the committed ncu capture (profiles/r01_final_*.txt): the integer ALU pipe is the binding unit"},
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = cores * args.cpu_images_per_core
        cpu_imgs = [images[i % B].copy() for i in range(n_cpu)]
        v, dt, kind = cpu_throughput(cpu_imgs, cores, NEAR, EFFORT)
        cpu = {"value": round(v, 3), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_cpu} images of the workload ({args.cpu_images_per_core} per core), one process per core, encode+decode, {dt:.1f} s"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * dev_s / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload, "mapping": codec.last_mapping, "l2": "inputs larger than L2 (batch pixels >> 126 MB)",
                       "encode_mpix_s": round(world * B * npx / enc_s / 1e6, 3), "decode_mpix_s": round(world * B * npx / dec_s / 1e6, 3),
                       "bits_per_pixel": round(8.0 * stream_bytes / (B * npx), 4), "parity": parity},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": B * npx + e2e_stream_bytes,
                    "d2h_bytes_per_step": e2e_stream_bytes + B * npx, "ms_per_step": round(1e3 * e2e_s / args.steps, 3),
                    "host_wall_ms_per_step": round(1e3 * e2e_wall / args.steps, 3)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
