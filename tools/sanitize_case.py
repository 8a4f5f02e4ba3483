"""Small mixed workload for compute-sanitizer: every kernel family once, checked against round trips."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nblic_image_compression_b200 import api
from nblic_image_compression_b200.synth import gen
codec = api.Codec(0)
imgs = [gen(24, 40, 1), gen(17, 33, 2), gen(9, 70, 3), gen(3, 5, 4), gen(40, 7, 5)]
for effort, near in [(0, 0), (1, 0), (1, 3), (2, 0), (2, 2), (3, 0), (3, 1)]:
    streams, recs, st = codec.encode_batch(imgs, near, effort, want_recon=near > 0)
    assert all(s == 0 for s in st), st
    dec = codec.decode_batch(streams)
    for im, r, d in zip(imgs, recs, dec):
        assert np.array_equal(d[0], r if near else im)
codec.close()
codec = api.Codec(0, sequential=True)  # the sequential kernels live in the test build
codec.set_mapping(api.MAP_LANE)
streams, _, _ = codec.encode_batch(imgs, 0, 1)
assert all(np.array_equal(d[0], im) for im, d in zip(imgs, codec.decode_batch(streams)))
print("sanitize_case ok", codec.launches, "launches")
