#!/usr/bin/env python
"""verify.py-equivalent regression sweep on the GPU codec.

The reference's verify.py (verify.py:26-36,58-72) hard-codes effort 0 and drives the CLI one file at a
time; this sweep takes a directory of 8-bit gray images (.pgm / .bmp) or --synthetic N, runs every
(effort, near) pair as ONE batch through libnblic_b200.so, and checks for every image
  * the stream is byte-identical to the CPU checker's (the unmodified reference when oracle/_ref is built),
  * decode(stream) == the reference reconstruction, and max |decoded - original| <= near,
printing bpp and kernel time per setting.  The CPU checker is test infrastructure (oracle/), used only to
verify; nothing it computes is shipped.
"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nblic_image_compression_b200 import api
from nblic_image_compression_b200.synth import gen
from cpu_codecs import Oracle, Ref, build_oracle, load_bmp_gray


def load_pgm(path):
    d = open(path, "rb").read()
    assert d[:2] == b"P5"
    toks, p = [], 2
    while len(toks) < 3:
        while d[p:p + 1].isspace(): p += 1
        if d[p:p + 1] == b"#":
            p = d.index(b"\n", p); continue
        q = p
        while d[q:q + 1].isdigit(): q += 1
        toks.append(int(d[p:q])); p = q
    w, h, _ = toks
    return np.frombuffer(d, np.uint8, count=w * h, offset=p + 1).reshape(h, w).copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("directory", nargs="?")
    ap.add_argument("--synthetic", type=int, default=0, help="use N synthetic 256x384 images instead of a directory")
    ap.add_argument("--efforts", default="0,1,2,3")
    ap.add_argument("--nears", default="0,1,2,3")
    args = ap.parse_args()
    if args.synthetic:
        images = [gen(256, 384, s) for s in range(args.synthetic)]
    else:
        files = sorted(f for f in os.listdir(args.directory) if f.lower().endswith((".pgm", ".bmp")))
        images = [load_pgm(os.path.join(args.directory, f)) if f.lower().endswith(".pgm") else load_bmp_gray(os.path.join(args.directory, f)) for f in files]
    build_oracle()
    chk = Ref() if Ref.available() else Oracle()
    codec = api.Codec(0)
    px = sum(im.size for im in images)
    bad = 0
    for effort in map(int, args.efforts.split(",")):
        for near in map(int, args.nears.split(",")):
            if effort == 0 and near:
                continue
            streams, recs, status = codec.encode_batch(images, near, effort, want_recon=near > 0)
            enc_ms = codec.last_coder_ms
            dec = codec.decode_batch(streams)
            dec_ms = codec.last_coder_ms
            n_bad = 0
            for im, s, r, d in zip(images, streams, recs, dec):
                exp, rec = (chk.q_encode(im), im) if effort == 0 else chk.n_encode(im, near, effort)[:2]
                ok = s == exp and d is not None and np.array_equal(d[0], rec) and int(np.abs(d[0].astype(int) - im.astype(int)).max()) <= near
                if near:
                    ok = ok and np.array_equal(r, rec)
                n_bad += not ok
            bad += n_bad
            print(f"effort {effort} near {near}: {len(images)} images, {8 * sum(map(len, streams)) / px:.4f} bpp, "
                  f"encode {px / enc_ms / 1e3:.1f} MPix/s, decode {px / dec_ms / 1e3:.1f} MPix/s, checker={chk.name}, mismatches={n_bad}", flush=True)
    print("TOTAL MISMATCHES", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
