"""One-off runs of the BASELINE.json configs that are too slow for the test suite (single large
streams): prints size / SHA-256 / timings next to the reference's golden values (tests/golden/manifest.json).
usage: python tools/config_runs.py [config3] [config4]"""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nblic_image_compression_b200 import api

man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
codec = api.Codec(0)

def run(h, w, seed, effort, near, key):
    npx = h * w
    d = torch.empty(npx, dtype=torch.uint8, device="cuda:0")
    codec.synth_device(d.data_ptr(), h, w, seed)
    cap = api.stream_bound(h, w)
    ds = torch.empty(cap, dtype=torch.uint8, device="cuda:0")
    dd = torch.empty(npx, dtype=torch.uint8, device="cuda:0")
    dr = torch.empty(npx, dtype=torch.uint8, device="cuda:0") if near else None
    off = np.zeros(1, np.uint64)
    t0 = time.time()
    so, st, rc = codec.encode_device(d.data_ptr(), off, np.array([h], np.int32), np.array([w], np.int32), near, effort, ds.data_ptr(), cap,
                                     dr.data_ptr() if near else 0)
    t1 = time.time()
    stream = ds[: int(so[1])].cpu().numpy().tobytes()
    ent = man["synthetic"][f"{h}x{w}_s{seed}"]["streams"][key]
    ok_enc = len(stream) == ent["bytes"] and hashlib.sha256(stream).hexdigest() == ent["sha256"]
    st2, rc2 = codec.decode_device(ds.data_ptr(), so, dd.data_ptr(), off)
    t2 = time.time()
    ok_dec = bool(torch.equal(dd, dr if near else d))
    print(json.dumps({"image": f"synthetic {h}x{w} seed {seed}", "setting": key, "bytes": len(stream), "reference_bytes": ent["bytes"],
                      "sha256_16": hashlib.sha256(stream).hexdigest()[:16], "encode_bit_exact": ok_enc, "decode_matches": ok_dec,
                      "encode_s": round(t1 - t0, 2), "decode_s": round(t2 - t1, 2), "mapping": codec.last_mapping}), flush=True)

which = sys.argv[1:] or ["config4", "config3"]
if "config4" in which:
    run(2048, 2048, 0, 2, 2, "e2n2")   # configs[3] synthetic part: reference 707198 B, 63fdbf3b53cdfe8e
if "config3" in which:
    run(4096, 4096, 0, 3, 0, "e3n0")   # configs[2]: reference 7260363 B, c79098a1831a46c6
