"""First-light script for the GPU box: parity of every mode on small inputs + rough timings.
Not a test and not the bench; prints what it finds."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nblic_image_compression_b200 import api
from nblic_image_compression_b200.synth import gen
from cpu_codecs import Oracle, build_oracle
from cases import edge_cases, SETTINGS_EDGE

build_oracle()
orc = Oracle()
codec = api.Codec(0, sequential=True)  # test build: also carries the sequential kernels behind MAP_LANE

def oracle_enc(img, effort, near):
    if effort == 0:
        return orc.q_encode(img), img
    s, rec, _, _ = orc.n_encode(img, near, effort)
    return s, rec

def check(images, effort, near, mapping, label):
    codec.set_mapping(mapping)
    t0 = time.time()
    streams, recs, status = codec.encode_batch(images, near, effort, want_recon=near > 0)
    t1 = time.time()
    enc_ms = codec.last_coder_ms
    bad = 0
    exp = [oracle_enc(im, effort, near) for im in images]
    for k, (s, e) in enumerate(zip(streams, exp)):
        if s != e[0]:
            bad += 1
            if bad <= 3:
                n = next((i for i in range(min(len(s or b''), len(e[0]))) if s[i] != e[0][i]), -1)
                print(f"   MISMATCH enc {label} img{k} shape={images[k].shape} len {len(s) if s else None} vs {len(e[0])} first diff @{n} status={status[k]}")
        if near > 0 and recs[k] is not None and not np.array_equal(recs[k], e[1]):
            bad += 1
            print(f"   MISMATCH recon {label} img{k}")
    t2 = time.time()
    dec = codec.decode_batch([e[0] for e in exp])
    t3 = time.time()
    dec_ms = codec.last_coder_ms
    for k, (d, e) in enumerate(zip(dec, exp)):
        if d is None or not np.array_equal(d[0], e[1]):
            bad += 1
            if bad <= 6:
                print(f"   MISMATCH dec {label} img{k} shape={images[k].shape} {'None' if d is None else int((d[0]!=e[1]).sum())}")
    px = sum(im.size for im in images)
    print(f"{label:40s} map={codec.last_mapping} n={len(images)} px={px} bad={bad} enc_kernel={enc_ms:.1f}ms dec_kernel={dec_ms:.1f}ms "
          f"enc_wall={1e3*(t1-t0):.0f}ms dec_wall={1e3*(t3-t2):.0f}ms", flush=True)
    return bad

total_bad = 0
edges = [im for _, im in edge_cases()]
for effort, near in SETTINGS_EDGE:
    for mapping in (api.MAP_WARP, api.MAP_LANE):
        total_bad += check(edges, effort, near, mapping, f"edge e{effort}n{near}")

# Kodak-size synthetic images (same shape as Kodak), e0/e1 timing
imgs = [gen(512, 768, s) for s in range(8)]
for effort, near in [(0, 0), (1, 0), (1, 2)]:
    for mapping in (api.MAP_WARP, api.MAP_LANE):
        total_bad += check(imgs, effort, near, mapping, f"synth 512x768 e{effort}n{near}")
small = [gen(96, 128, s) for s in range(8)]
for effort, near in [(2, 0), (2, 2), (3, 0), (3, 3)]:
    for mapping in (api.MAP_WARP, api.MAP_LANE):
        total_bad += check(small, effort, near, mapping, f"synth 96x128 e{effort}n{near}")
print("TOTAL BAD", total_bad)
sys.exit(1 if total_bad else 0)
