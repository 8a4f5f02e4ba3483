"""One-off: the size limit of the reference (NBLIC_MAX_IMG_SIZE = 100 000 000 pixels, NBLIC.h:31) through the
GPU path: a 10000 x 10000 synthetic image, effort 0 (encode + decode) and effort 1 (encode), streams compared
with the CPU checker's (tests/cpu_codecs.py; checker only)."""
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from nblic_image_compression_b200 import api
from cpu_codecs import Oracle, Ref, build_oracle

build_oracle()
chk = Ref() if Ref.available() else Oracle()
codec = api.Codec(0)
h = w = 10000
d = torch.empty(h * w, dtype=torch.uint8, device="cuda:0")
codec.synth_device(d.data_ptr(), h, w, 42)
img = d.cpu().numpy().reshape(h, w)
for effort in (0, 1):
    t0 = time.time()
    streams, _, status = codec.encode_batch([img], 0, effort)
    t1 = time.time()
    exp = chk.q_encode(img) if effort == 0 else chk.n_encode(img, 0, 1)[0]
    rec = {"image": f"synthetic {h}x{w} seed 42", "effort": effort, "status": status, "bytes": len(streams[0] or b""), "checker_bytes": len(exp),
           "bit_exact": streams[0] == exp, "encode_s": round(t1 - t0, 1), "checker": chk.name}
    rec["encode_mapping"] = codec.last_mapping
    if effort == 0 and not os.environ.get("NBLIC_SKIP_DECODE"):
        t2 = time.time()
        dec = codec.decode_batch(streams)
        rec["decode_s"] = round(time.time() - t2, 1)
        rec["decode_matches"] = bool(dec[0] is not None and np.array_equal(dec[0][0], img))
    print(json.dumps(rec), flush=True)
