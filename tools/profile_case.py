"""One small batch through encode + decode, for ncu: python tools/profile_case.py N H W EFFORT NEAR [auto|warp|warp4|lane] [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nblic_image_compression_b200 import api

n, h, w, effort, near = (int(x) for x in sys.argv[1:6])
mapping = {"warp": api.MAP_WARP, "warp4": api.MAP_WARP4, "lane": api.MAP_LANE, "auto": api.MAP_AUTO}[sys.argv[6] if len(sys.argv) > 6 else "auto"]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 1
codec = api.Codec(0, mapping)
npx = h * w
d_pix = torch.empty(n * npx, dtype=torch.uint8, device="cuda:0")
codec.synth_device_batch(d_pix.data_ptr(), n, h, w, 0)
cap = n * api.stream_bound(h, w)
d_str = torch.empty(cap, dtype=torch.uint8, device="cuda:0")
d_dec = torch.empty(n * npx, dtype=torch.uint8, device="cuda:0")
d_rec = torch.empty(n * npx, dtype=torch.uint8, device="cuda:0") if near else None
off = np.arange(n, dtype=np.uint64) * npx
hs, ws = np.full(n, h, np.int32), np.full(n, w, np.int32)
for _ in range(reps):
    so, st, rc = codec.encode_device(d_pix.data_ptr(), off, hs, ws, near, effort, d_str.data_ptr(), cap, d_rec.data_ptr() if near else 0)
    enc_ms = codec.last_coder_ms
    assert rc == 0
    st, rc = codec.decode_device(d_str.data_ptr(), so, d_dec.data_ptr(), off)
    dec_ms = codec.last_coder_ms
    assert rc == 0
    assert torch.equal(d_dec, d_rec if near else d_pix)
    print(f"n={n} {h}x{w} e{effort}n{near} map={codec.last_mapping} enc {enc_ms:.2f} ms ({n*npx/enc_ms/1e3:.1f} MPix/s) dec {dec_ms:.2f} ms ({n*npx/dec_ms/1e3:.1f} MPix/s) bpp {8*int(so[-1])/(n*npx):.3f}", flush=True)
