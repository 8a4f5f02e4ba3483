"""Single-image / small-batch encode latency through the host-buffer API (development aid).
usage: python tools/latency_check.py [h w [n [effort]]]   (NBLIC_B200_E1PIPE_TIMING=1 prints the pipeline's stage times)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nblic_image_compression_b200 import api
from nblic_image_compression_b200.synth import gen

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 768)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
effort = int(sys.argv[4]) if len(sys.argv) > 4 else 1
imgs = [gen(h, w, s) for s in range(n)]
c = api.Codec(0)
for mode, name in ((api.PIPE_NEVER, "one warp per image"), (api.PIPE_ALWAYS, "whole-GPU pipeline")):
    c.set_pipeline(mode)
    best, ref = 1e9, None
    for it in range(3):
        t0 = time.perf_counter()
        streams, _, st = c.encode_batch(imgs, 0, effort)
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    print(f"{n} x {h}x{w} e{effort} {name}: {1e3 * best:.2f} ms ({n * h * w / best / 1e6:.1f} MPix/s), coder kernels {c.last_coder_ms:.2f} ms, mapping {c.last_mapping}", flush=True)
    if ref is None:
        ref = streams
keep = streams
c.set_pipeline(api.PIPE_NEVER)
assert c.encode_batch(imgs, 0, effort)[0] == keep
print("bytes equal")
c.close()
