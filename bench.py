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
    roofline  dominant kernel = the longer of the two coder launches (decode), timed by the library's own CUDA
              events on its stream; the integer-issue peak is the MEASURED one (profiles/r02_int_issue_peak.json,
              tools/microbench/int_issue_peak.cu); ncu pipe / dram figures are quoted only from captures stamped
              with the fingerprint of the kernel sources this run was built from, else null
    cpu_baseline  the unmodified reference (oracle/_ref/libnblic_ref.so; else the oracle port) on all
              host cores, one process per core, bounded sample of the same workload (rank 0, N=1)

    config.per_effort   the other settings (e0, e1 -n2, e2 -n2, e3) on 1536 images of the same shape, device-resident
              and end to end, each with its own issue-roofline fraction (fixed 1 warm-up + 1 timed step each)
    strong    (N > 1) configs[4] as named: 10 000 images in total, split image by image over the ranks
              (shard.plan_shards), lengths gathered on rank 0; value = total pixels / max-over-ranks time

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

# More hardware work queues than the default 8 (must be set before CUDA initialises): the streams of the host lanes and
# of the single-image pipelines then get a queue each.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
def _load_json(*parts):
    try:
        return json.load(open(os.path.join(ROOT, *parts)))
    except Exception:
        return None


def issue_peak_lane_ops(sm_count, sm_mhz_max):
    """(lane-issue slots / s, source).  Measured by tools/microbench/int_issue_peak.cu on this pool's B200s: the best
    sustained rate of any integer instruction mix (IMAD on the FMA pipe and IADD3/LOP3 on the ALU pipe together)."""
    rec = _load_json("profiles", "r02_int_issue_peak.json")
    if rec and rec.get("peak_lane_ops_per_s"):
        return float(rec["peak_lane_ops_per_s"]), "measured: profiles/r02_int_issue_peak.json (%s)" % max(rec["ops"], key=lambda k: rec["ops"][k]["lane_ops_per_s"])
    return sm_count * 4 * 32 * sm_mhz_max * 1e6, "theoretical 1 warp-instruction / cycle / scheduler (no measurement on record)"


def profile_figures(kernel_key):
    """ncu figures of `kernel_key` from profiles/r02_ncu.json, or (None, reason) when the capture was not taken from
    the kernel sources this process is running."""
    from nblic_image_compression_b200.build import source_fingerprint
    rec = _load_json("profiles", "r02_ncu.json")
    if not rec:
        return None, "no ncu capture on record (profiles/r02_ncu.json)"
    if rec.get("csrc_fingerprint") != source_fingerprint():
        return None, "stale: profiles/r02_ncu.json was captured from kernel sources %s, this build is %s" % (rec.get("csrc_fingerprint"), source_fingerprint())
    k = rec.get("kernels", {}).get(kernel_key)
    if not k:
        return None, "kernel %s not in profiles/r02_ncu.json" % kernel_key
    if k.get("inst_per_pixel") is None:
        return None, "ncu could not collect the counters of %s in the capture on record (profiles/r02_ncu.json: %s)" % (kernel_key, k.get("note", "no figures"))
    return k, "ncu --set full capture of this build (git %s, shape %s)" % (rec.get("git_head", "?")[:12], k.get("shape"))


def _cpu_one_core(img, near, effort):
    """(encode ms, decode ms, stream bytes) of the CPU codec on one core for one image (reference build when present)."""
    path, kind = _cpu_lib()
    lib = C.CDLL(path)
    u8p, u16p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_int)
    h, w = img.shape
    src = img.copy()
    out = np.zeros(2 * h * w + 65536, dtype=np.uint8)
    dec = np.zeros(h * w, dtype=np.uint8)
    n_, e_, hh, ww = C.c_int(near), C.c_int(effort), C.c_int(), C.c_int()
    t0 = time.perf_counter()
    if kind == "reference":
        if effort == 0:
            n = 2 * lib.QNBLICcompress(out.ctypes.data_as(u16p), src.ctypes.data_as(u8p), h, w)
        else:
            n = lib.NBLICcompress(0, out.ctypes.data_as(u8p), src.ctypes.data_as(u8p), h, w, C.byref(n_), C.byref(e_))
        t1 = time.perf_counter()
        if effort == 0:
            rc = lib.QNBLICdecompress(out.ctypes.data_as(u16p), dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww))
        else:
            rc = lib.NBLICdecompress(0, out.ctypes.data_as(u8p), dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww), C.byref(n_), C.byref(e_))
    else:
        lib.oracle_n_decode.argtypes = [u8p, C.c_long, u8p, ip, ip, ip, ip]
        lib.oracle_q_decode.argtypes = [u16p, C.c_long, u8p, ip, ip]
        if effort == 0:
            n = 2 * lib.oracle_q_encode(src.ctypes.data_as(u8p), h, w, out.ctypes.data_as(u16p))
        else:
            n = lib.oracle_n_encode(src.ctypes.data_as(u8p), h, w, C.byref(n_), C.byref(e_), out.ctypes.data_as(u8p))
        t1 = time.perf_counter()
        if effort == 0:
            rc = lib.oracle_q_decode(out.ctypes.data_as(u16p), n // 2, dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww))
        else:
            rc = lib.oracle_n_decode(out.ctypes.data_as(u8p), n, dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww), C.byref(n_), C.byref(e_))
    t2 = time.perf_counter()
    assert n > 0 and rc == 0 and np.array_equal(dec.reshape(h, w), img)
    return 1e3 * (t1 - t0), 1e3 * (t2 - t1), int(n), bytes(out[:n]), kind


def latency_records(codec, api, cores):
    """Few-stream configs through the host-buffer API (best of 3 wall-clock calls after one warm-up), with the CPU
    codec on ONE core beside them; every GPU stream is compared byte for byte with the CPU's."""
    from nblic_image_compression_b200.synth import gen
    out = {}
    try:
        gold = os.path.join(ROOT, "tests", "golden", "kodak_e1n0")
        names = sorted(f for f in os.listdir(gold) if f.endswith(".nblic"))
        kodak = [d[0] for d in codec.decode_batch([open(os.path.join(gold, f), "rb").read() for f in names])]
        cases = [("configs[0]: Kodak 01, -n0 -e0", [kodak[0]], 0), ("configs[1]: 24 Kodak images, -n0 -e1", kodak, 1),
                 ("Kodak 01, -n0 -e1", [kodak[0]], 1), ("one synthetic 4096x4096 image, -n0 -e1", [gen(4096, 4096, 0)], 1)]
        for label, imgs, effort in cases:
            px = sum(im.size for im in imgs)
            enc_ms, dec_ms = [], []
            for it in range(2 if px > 4e6 and len(imgs) == 1 else 4):  # one warm-up, then best of 3 (of 1 for the 19-second decode of the large image)
                t0 = time.perf_counter()
                streams, _, st = codec.encode_batch(imgs, 0, effort)
                t1 = time.perf_counter()
                mapping = codec.last_mapping
                dec = codec.decode_batch(streams)
                t2 = time.perf_counter()
                if it:
                    enc_ms.append(1e3 * (t1 - t0)); dec_ms.append(1e3 * (t2 - t1))
            assert all(np.array_equal(d[0], im) for d, im in zip(dec, imgs))
            ce, cd, cn, cbytes, kind = _cpu_one_core(imgs[0], 0, effort)
            rec = {"images": len(imgs), "mpixel": round(px / 1e6, 3), "encode_ms": round(min(enc_ms), 3), "decode_ms": round(min(dec_ms), 3),
                   "encode_mpix_s": round(px / min(enc_ms) / 1e3, 2), "decode_mpix_s": round(px / min(dec_ms) / 1e3, 2), "encode_mapping": mapping,
                   "cpu_one_core_first_image": {"encode_ms": round(ce, 2), "decode_ms": round(cd, 2), "kind": kind},
                   "first_stream_equals_cpu": streams[0] == cbytes, "timing": "host wall clock around the host-buffer calls (pageable numpy buffers), best of %d after a warm-up" % len(enc_ms)}
            out[label] = rec
    except Exception as ex:  # pragma: no cover
        out["error"] = repr(ex)
    return out


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
    ap.add_argument("--per-effort-images", type=int, default=int(os.environ.get("NBLIC_BENCH_PER_EFFORT", "1536")),
                    help="images of the sub-records for the other efforts (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (kernel experiments only; the line is then not a valid bench line)")
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

    import hashlib
    import torch
    from nblic_image_compression_b200 import api
    from nblic_image_compression_b200.shard import plan_shards

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    codec = api.Codec(local_rank, {"auto": api.MAP_AUTO, "warp": api.MAP_WARP, "lane": api.MAP_LANE}[args.mapping])
    lib = codec.lib
    B = args.images
    npx = H * W
    dev = torch.device("cuda", local_rank)
    d_pixels = torch.empty(B * npx, dtype=torch.uint8, device=dev)
    codec.synth_device_batch(d_pixels.data_ptr(), B, H, W, rank * B)  # one launch per 32768 images
    stream_cap = B * api.stream_bound(H, W)
    d_streams = torch.empty(stream_cap, dtype=torch.uint8, device=dev)
    d_decoded = torch.empty(B * npx, dtype=torch.uint8, device=dev)
    d_recon = torch.empty(max(args.per_effort_images, 1) * npx, dtype=torch.uint8, device=dev)
    ext = torch.cuda.ExternalStream(codec.stream_handle, device=dev)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    peaks = _load_json("MEASURED_PEAKS.json") or {}
    issue_peak, issue_src = issue_peak_lane_ops(sm_count, float(peaks.get("sm_max_mhz", 1965.0)))

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

    # pinned host staging of the end-to-end arm (allocated once, first-touched by the untimed pass)
    bound = H * W + 8192  # caller-side capacity per stream: 8 bpp + header room (a larger stream is reported as overflow)
    h_pixels = torch.empty(B * npx, dtype=torch.uint8, pin_memory=True)
    h_streams = torch.empty(B * bound, dtype=torch.uint8, pin_memory=True)
    h_decoded = torch.empty(B * npx, dtype=torch.uint8, pin_memory=True)
    h_np = h_pixels.numpy()

    def measure(n, near, effort, warmup, steps, e2e=True, check=True):
        """One setting on the first n images of the batch: device-resident and (optionally) host-buffer arms."""
        pix_off = np.arange(n, dtype=np.uint64) * npx
        pix_cap = np.full(n, npx, dtype=np.uint64)
        hs = np.full(n, H, dtype=np.int32)
        ws = np.full(n, W, dtype=np.int32)
        state = {}
        target = d_recon if near else d_pixels

        def step_device():
            off, st, rc = codec.encode_device(d_pixels.data_ptr(), pix_off, hs, ws, near, effort, d_streams.data_ptr(), stream_cap,
                                              d_recon.data_ptr() if near else 0)
            state["enc_ms"] = codec.last_coder_ms
            state["enc_slots"] = codec.last_slots
            assert rc == 0, st
            st2, rc2 = codec.decode_device(d_streams.data_ptr(), off, d_decoded.data_ptr(), pix_off, pix_cap)
            state["dec_ms"] = codec.last_coder_ms
            state["dec_slots"] = codec.last_slots
            assert rc2 == 0, st2
            state["stream_off"] = off

        for _ in range(warmup):
            step_device()
        if check:
            assert torch.equal(target[: n * npx], d_decoded[: n * npx]), "decode(encode(x)) != x (or != the reconstruction)"
            if near:
                assert int((d_recon[: n * npx].to(torch.int16) - d_pixels[: n * npx].to(torch.int16)).abs().max()) <= near
        barrier()
        launches0 = codec.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        enc_ms, dec_ms = [], []
        e0.record(ext)
        for _ in range(steps):
            step_device()
            enc_ms.append(state["enc_ms"]); dec_ms.append(state["dec_ms"])
        e1.record(ext)
        barrier()
        res = {"n": n, "near": near, "effort": effort, "launches": codec.launches - launches0,
               "dev_s": max_over_ranks(e0.elapsed_time(e1) / 1e3), "enc_s": float(np.mean(enc_ms)) / 1e3, "dec_s": float(np.mean(dec_ms)) / 1e3,
               "stream_off": state["stream_off"].copy(), "stream_bytes": int(state["stream_off"][-1]), "mapping": codec.last_mapping,
               "enc_slots": state["enc_slots"], "dec_slots": state["dec_slots"]}
        if not e2e:
            return res
        h_pixels[: n * npx].copy_(d_pixels[: n * npx])
        images = [h_np[i * npx:(i + 1) * npx].reshape(H, W) for i in range(n)]
        outs = [h_streams.numpy()[i * bound:(i + 1) * bound] for i in range(n)]
        dec_views = [h_decoded.numpy()[i * npx:(i + 1) * npx] for i in range(n)]
        img_ptrs, out_ptrs, dec_ptrs = api._ptr_array(images), api._ptr_array(outs), api._ptr_array(dec_views)
        hs_c = (C.c_int * n)(*([H] * n)); ws_c = (C.c_int * n)(*([W] * n))
        caps = (C.c_size_t * n)(*([bound] * n)); lens = (C.c_size_t * n)()
        dcaps = (C.c_size_t * n)(*([npx] * n))
        status = (C.c_int * n)()

        def step_e2e():
            rc = lib.nblic_b200_encode_batch(codec.ctx, n, img_ptrs, hs_c, ws_c, near, effort, out_ptrs, caps, lens, None, status)
            assert rc == 0, codec._err()
            rc = lib.nblic_b200_decode_batch(codec.ctx, n, out_ptrs, lens, dec_ptrs, dcaps, None, None, None, None, status)
            assert rc == 0, codec._err()

        step_e2e()  # untimed: first touch of the pinned buffers and of the library's staging allocations
        if check:
            assert torch.equal(h_decoded[: n * npx], target[: n * npx].cpu()), "e2e decode(encode(x)) mismatch"
        barrier()
        e0.record(ext)
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        e1.record(ext)
        barrier()
        res["e2e_wall"] = time.perf_counter() - t0
        res["e2e_s"] = max_over_ranks(max(e0.elapsed_time(e1) / 1e3, 0.0))
        res["e2e_lens"] = [int(lens[i]) for i in range(n)]
        res["images"], res["outs"] = images, outs
        return res

    def issue_record(pixels, seconds, effort):
        rate = pixels / seconds * SLOTS_PER_PIXEL[effort]
        return {"slots_per_pixel": SLOTS_PER_PIXEL[effort], "achieved": round(rate / 1e12, 4), "peak": round(issue_peak / 1e12, 3),
                "frac": round(rate / issue_peak, 6)}

    # ---- the workload of the metric: configs[4] ------------------------------------------------------
    sampler = ClockSampler(local_rank)
    main_r = measure(B, NEAR, EFFORT, args.warmup, args.steps, e2e=not args.no_e2e)
    clocks = sampler.stop()
    value = world * B * npx * 2 * args.steps / main_r["dev_s"] / 1e6
    e2e_rec = None
    if not args.no_e2e:
        e2e_value = world * B * npx * 2 * args.steps / main_r["e2e_s"] / 1e6
        e2e_bytes = int(sum(main_r["e2e_lens"]))
        e2e_rec = {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": B * npx + e2e_bytes, "d2h_bytes_per_step": e2e_bytes + B * npx,
                   "ms_per_step": round(1e3 * main_r["e2e_s"] / args.steps, 3), "host_wall_ms_per_step": round(1e3 * main_r["e2e_wall"] / args.steps, 3)}

    # ---- parity of a random 1 % sample against the CPU codec: BOTH arms (checker only, outside the timed regions) ----
    parity = None
    if rank == 0:
        try:
            sample = sorted(np.random.default_rng(0).choice(B, size=max(1, B // 100), replace=False).tolist())
            off = main_r["stream_off"]
            dev_streams = {i: bytes(d_streams[int(off[i]):int(off[i + 1])].cpu().numpy()) for i in sample}  # device-resident arm's packed output
            imgs = main_r.get("images") or [d_pixels[i * npx:(i + 1) * npx].cpu().numpy().reshape(H, W) if i in dev_streams else None for i in range(B)]
            ref_hashes, kind = cpu_stream_hashes(imgs, sample, cores, NEAR, EFFORT)
            who = "the unmodified reference" if kind == "reference" else "the oracle port"
            bad = [i for i in sample if ref_hashes[i] != (len(dev_streams[i]), hashlib.sha256(dev_streams[i]).hexdigest())]
            arms = "device-resident arm"
            if not args.no_e2e:
                lens_, outs_ = main_r["e2e_lens"], main_r["outs"]
                bad += [i for i in sample if ref_hashes[i] != (lens_[i], hashlib.sha256(bytes(outs_[i][: lens_[i]])).hexdigest())]
                arms = "device-resident and host-buffer arms"
            cpu_total = sum(ref_hashes[i][0] for i in sample)
            gpu_total = sum(len(dev_streams[i]) for i in sample)
            parity = (f"bit-exact vs {who} on {len(sample)} random images of the batch (1 %), {arms}; sample bytes GPU {gpu_total} == CPU {cpu_total}"
                      if not bad and cpu_total == gpu_total else f"MISMATCH vs {who} on images {sorted(set(bad))[:8]} (bytes GPU {gpu_total} / CPU {cpu_total})")
        except Exception as ex:  # pragma: no cover
            parity = f"CPU checker unavailable: {ex}"

    # ---- roofline of the dominant kernel (the longer of the encode / decode coder launches) --------
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    enc_s, dec_s = main_r["enc_s"], main_r["dec_s"]
    dom_is_dec = dec_s >= enc_s
    dom_s = dec_s if dom_is_dec else enc_s
    alg_bytes = B * npx + main_r["stream_bytes"]  # every pixel and every stream byte cross HBM once (read one, write the other)
    fig, fig_note = profile_figures("e1_decode" if dom_is_dec else "e1_encode_lossless")
    roofline = {
        "bound": "hbm", "kernel": "effort-1 %s coder kernel (%s)" % ("decode" if dom_is_dec else "lossless encode", (fig or {}).get("kernel", "see profiles/")),
        "achieved": round(alg_bytes / dom_s / 1e9, 3), "peak": hbm_peak, "unit": "GB/s",
        "frac": round(alg_bytes / dom_s / 1e9 / hbm_peak, 6),
        "traffic": int(fig["dram_bytes_per_pixel"] * B * npx) if fig and fig.get("dram_bytes_per_pixel") is not None else None,
        "traffic_note": fig_note, "peak_source": peak_src,
        "kernel_ms": round(1e3 * dom_s, 3), "encode_kernel_ms": round(1e3 * enc_s, 3), "decode_kernel_ms": round(1e3 * dec_s, 3),
        "algorithmic_bytes_per_launch": int(alg_bytes),
        "note": "the path is bound by dependent integer issue, not HBM or tensor throughput (SURVEY.md 8(d)): see `issue`",
        "issue": dict(issue_record(B * npx, dom_s, EFFORT), unit="T lane-issue-slots/s", peak_source=issue_src,
                      encode=issue_record(B * npx, enc_s, EFFORT)["frac"], decode=issue_record(B * npx, dec_s, EFFORT)["frac"],
                      ncu=(fig or {}).get("pipes"), ncu_note=fig_note),
    }

    # ---- the other settings (sub-records; fixed 1 warm-up + 1 timed step, not scaled by --steps) ----
    per_effort = None
    if args.per_effort_images > 0:
        per_effort = {}
        ne = min(args.per_effort_images, B)
        for effort, near in ((0, 0), (1, 2), (2, 2), (3, 0)):
            r = measure(ne, near, effort, 1, 1, e2e=not args.no_e2e)
            px = world * ne * npx
            rec = {"images_per_gpu": ne, "value": round(2 * px / r["dev_s"] / 1e6, 3),
                   "encode_mpix_s": round(px / r["enc_s"] / 1e6, 3), "decode_mpix_s": round(px / r["dec_s"] / 1e6, 3),
                   "bits_per_pixel": round(8.0 * r["stream_bytes"] / (ne * npx), 4), "mapping": r["mapping"],
                   "issue_frac_encode": issue_record(ne * npx, r["enc_s"], effort)["frac"],
                   "issue_frac_decode": issue_record(ne * npx, r["dec_s"], effort)["frac"],
                   "fill_encode": round(ne / max(r["enc_slots"], 1), 3), "fill_decode": round(ne / max(r["dec_slots"], 1), 3)}
            if not args.no_e2e:
                rec["e2e"] = round(2 * px / r["e2e_s"] / 1e6, 3)
            per_effort[f"e{effort}n{near}"] = rec

    # ---- latency configs (configs[0], configs[1], and a single 4096 x 4096 image): few streams, so the lossless encoders
    # run as whole-GPU pipelines (pipe_qnblic.cuh / pipe_nblic.cuh); a decoder is one serial chain per image ----
    latency = None
    if rank == 0 and world == 1 and args.per_effort_images > 0:
        latency = latency_records(codec, api, cores)

    # ---- strong scaling of configs[4] as named: B images in total over the ranks ------------------------
    strong = None
    if world > 1:
        mine = plan_shards([npx] * B, world)[rank]  # equal sizes: image i goes to rank i % world
        assert mine == list(range(rank, B, world))
        codec.synth_device_batch(d_pixels.data_ptr(), len(mine), H, W, rank, world)  # seeds rank, rank + world, ...
        r = measure(len(mine), NEAR, EFFORT, 1, max(args.steps, 2), e2e=False)
        lens_local = np.diff(r["stream_off"].astype(np.int64)).tolist()
        parts = [None] * world
        dist.all_gather_object(parts, (mine, lens_local))  # rank 0 learns every stream length, in batch order
        if rank == 0:
            lengths = np.zeros(B, dtype=np.int64)
            for idxs, ls in parts:
                lengths[np.asarray(idxs, dtype=np.int64)] = ls
            strong = {"images_total": B, "images_per_gpu": len(mine), "value": round(2 * B * npx * max(args.steps, 2) / r["dev_s"] / 1e6, 3), "unit": UNIT,
                      "stream_bytes_total": int(lengths.sum()), "fill_encode": round(len(mine) / max(r["enc_slots"], 1), 3),
                      "fill_decode": round(len(mine) / max(r["dec_slots"], 1), 3), "mapping": r["mapping"],
                      "note": "scaling 'strong': total work fixed; efficiency = value / (N x the N=1 line's value)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = cores * args.cpu_images_per_core
        cpu_imgs = [d_pixels[(i % B) * npx:((i % B) + 1) * npx].cpu().numpy().reshape(H, W) for i in range(n_cpu)]
        v, dt, kind = cpu_throughput(cpu_imgs, cores, NEAR, EFFORT)
        cpu = {"value": round(v, 3), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_cpu} images of the workload ({args.cpu_images_per_core} per core), one process per core, encode+decode, {dt:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * main_r["dev_s"] / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload, "mapping": main_r["mapping"], "l2": "inputs larger than L2 (batch pixels >> 126 MB)",
                       "encode_mpix_s": round(world * B * npx / enc_s / 1e6, 3), "decode_mpix_s": round(world * B * npx / dec_s / 1e6, 3),
                       "bits_per_pixel": round(8.0 * main_r["stream_bytes"] / (B * npx), 4), "parity": parity,
                       "resident_slots": {"encode": main_r["enc_slots"], "decode": main_r["dec_slots"]}, "per_effort": per_effort,
                       "latency": latency},
            "e2e": e2e_rec, "gpu_launches": int(main_r["launches"]), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        if strong:
            line["strong"] = strong
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
