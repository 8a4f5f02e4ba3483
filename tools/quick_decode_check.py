"""Development check of the effort-1 decoder variants (lanes per stream, bias table placement):
correctness on small mixed batches against the library's own one-stream-per-warp decoder and the source pixels,
then timing at a bench-like shape.  usage: python tools/quick_decode_check.py [N H W]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nblic_image_compression_b200 import api
from nblic_image_compression_b200.synth import gen

VARIANTS = [("32", None), ("8", None), ("8", "1")]


def make_codec(lps, ctx_smem):
    os.environ["NBLIC_B200_LPS"] = lps
    if ctx_smem:
        os.environ["NBLIC_B200_FOREST_SMEM"] = ctx_smem
    else:
        os.environ.pop("NBLIC_B200_FOREST_SMEM", None)
    return api.Codec(0)


def correctness():
    rng = np.random.default_rng(1)
    shapes = [(40, 56)] * 9 + [(33, 17)] * 3 + [(1, 5)] + [(9, 13)] * 5 + [(64, 200)] * 2 + [(2, 2)] + [(5, 1)] * 4 + [(3, 70)] * 6 + [(130, 9)] * 2
    imgs = []
    for k, (h, w) in enumerate(shapes):
        kind = k % 4
        if kind == 0: im = gen(h, w, 100 + k)
        elif kind == 1: im = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        elif kind == 2: im = (rng.integers(0, 2, size=(h, w)) * 255).astype(np.uint8)
        else: im = np.clip(np.cumsum(rng.normal(0, 6, size=(h, w)), axis=1) + 128, 0, 255).astype(np.uint8)
        imgs.append(im)
    base = make_codec("32", None)
    ok = True
    for near in (0, 1, 2, 4, 9):
        streams, recs, st = base.encode_batch(imgs, near, 1, want_recon=near > 0)
        want = [r if near else im for r, im in zip(recs, imgs)]
        for lps, cs in VARIANTS:
            c = make_codec(lps, cs)
            out = c.decode_batch(streams)
            bad = [k for k, (d, t) in enumerate(zip(out, want)) if d is None or not np.array_equal(d[0], t)]
            if bad:
                ok = False
                k = bad[0]
                d = out[k]
                where = None if d is None else np.argwhere(d[0] != want[k])[:3].tolist()
                print(f"MISMATCH near={near} lps={lps} forest_smem={cs}: images {bad[:10]} shape {imgs[k].shape} first diffs {where} status {c.last_status[k]}", flush=True)
            else:
                print(f"ok near={near} lps={lps} forest_smem={cs} ({len(imgs)} images, mapping {c.last_mapping})", flush=True)
            c.close()
        # corrupt payloads must not fault
        bad_streams = [s[:16] + bytes(rng.integers(0, 256, size=len(s) - 16, dtype=np.uint8)) for s in streams[:12]] + [s[: 16 + (len(s) - 16) // 2] for s in streams[:12]]
        for lps, cs in VARIANTS[1:]:
            c = make_codec(lps, cs)
            c.decode_batch(bad_streams)
            c.close()
    base.close()
    return ok


def timing(n, h, w):
    npx = h * w
    base = make_codec("32", None)
    d_pix = torch.empty(n * npx, dtype=torch.uint8, device="cuda:0")
    base.synth_device_batch(d_pix.data_ptr(), n, h, w, 0)
    cap = n * api.stream_bound(h, w)
    d_str = torch.empty(cap, dtype=torch.uint8, device="cuda:0")
    d_dec = torch.empty(n * npx, dtype=torch.uint8, device="cuda:0")
    off = np.arange(n, dtype=np.uint64) * npx
    hs, ws = np.full(n, h, np.int32), np.full(n, w, np.int32)
    so, st, rc = base.encode_device(d_pix.data_ptr(), off, hs, ws, 0, 1, d_str.data_ptr(), cap)
    print(f"encode: {base.last_coder_ms:.1f} ms ({n * npx / base.last_coder_ms / 1e3:.0f} MPix/s), slots {base.last_slots}", flush=True)
    base.close()
    for lps, cs in VARIANTS:
        c = make_codec(lps, cs)
        best = 1e30
        for rep in range(2):
            d_dec.zero_()
            st, rc = c.decode_device(d_str.data_ptr(), so, d_dec.data_ptr(), off, np.full(n, npx, np.uint64))
            assert rc == 0
            best = min(best, c.last_coder_ms)
        same = torch.equal(d_dec, d_pix)
        print(f"decode lps={lps} forest_smem={cs}: {best:.1f} ms ({n * npx / best / 1e3:.0f} MPix/s) slots {c.last_slots} exact={same}", flush=True)
        c.close()


if __name__ == "__main__":
    good = correctness()
    print("correctness:", "PASS" if good else "FAIL", flush=True)
    a = [int(x) for x in sys.argv[1:4]] if len(sys.argv) >= 4 else [5000, 512, 512]
    timing(*a)
