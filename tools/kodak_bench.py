"""BASELINE.json configs[1]: all 24 Kodak images, -n0 -e1, one stream per warp -- plus the set replicated x64
(1536 images) to fill the GPU, and the other efforts.  The Kodak rasters are recovered by decoding the
committed reference streams (tests/golden/kodak_e1n0) on the GPU itself.  Prints one JSON line per case."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nblic_image_compression_b200 import api

codec = api.Codec(0)
gold = os.path.join(ROOT, "tests", "golden", "kodak_e1n0")
files = [open(os.path.join(gold, f), "rb").read() for f in sorted(os.listdir(gold))]
imgs = [d[0] for d in codec.decode_batch(files)]
assert len(imgs) == 24

def case(rep, effort, near, reps=2):
    batch = imgs * rep
    n = len(batch)
    sizes = np.array([im.size for im in batch], dtype=np.uint64)
    off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    total = int(sizes.sum())
    d_pix = torch.from_numpy(np.concatenate([im.reshape(-1) for im in batch])).to("cuda:0")
    cap = sum(api.stream_bound(*im.shape) for im in batch)
    d_str = torch.empty(cap, dtype=torch.uint8, device="cuda:0")
    d_dec = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    d_rec = torch.empty(total, dtype=torch.uint8, device="cuda:0") if near else None
    hs = np.array([im.shape[0] for im in batch], np.int32); ws = np.array([im.shape[1] for im in batch], np.int32)
    for _ in range(reps):
        so, st, rc = codec.encode_device(d_pix.data_ptr(), off, hs, ws, near, effort, d_str.data_ptr(), cap, d_rec.data_ptr() if near else 0)
        enc_ms = codec.last_coder_ms
        st2, rc2 = codec.decode_device(d_str.data_ptr(), so, d_dec.data_ptr(), off)
        dec_ms = codec.last_coder_ms
        assert rc == 0 and rc2 == 0 and torch.equal(d_dec, d_rec if near else d_pix)
    print(json.dumps({"case": f"Kodak x{rep} ({n} images, {total / 1e6:.1f} MPix)", "effort": effort, "near": near, "stream_bytes": int(so[-1]),
                      "bpp": round(8 * int(so[-1]) / total, 4), "encode_ms": round(enc_ms, 2), "decode_ms": round(dec_ms, 2),
                      "encode_mpix_s": round(total / enc_ms / 1e3, 1), "decode_mpix_s": round(total / dec_ms / 1e3, 1), "mapping": codec.last_mapping}), flush=True)

case(1, 1, 0)       # configs[1] as named: 24 streams -> a latency figure (expect sum 4891174 bytes)
case(64, 1, 0)      # the set replicated to fill the GPU
case(1, 0, 0); case(64, 0, 0)
case(1, 2, 2); case(64, 2, 2)
case(1, 3, 0); case(64, 3, 0)
