#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference (run in the authoring
container only: needs /root/reference and oracle/_ref/libnblic_ref.so, see oracle/Makefile).

Outputs (all under tests/golden/):
  kodak_e1n0/NN.nblic   the reference's `-n0 -e1` streams of img_kodak/NN.bmp (config 2).  They double
                        as the pixel source on the GPU box: decoding them gives the Kodak rasters.
  kodak_01_e0n0.nblic   config 1's stream.
  manifest.json         sizes + SHA-256 of every reference stream for 24 Kodak x {e0n0, e1..3 x n0..3},
                        of the decoded/reconstructed pixels, of the synthetic images of
                        SURVEY.md Appendix B, and of a deterministic edge-case suite (tests/cases.py).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

from cases import edge_cases, SETTINGS_EDGE  # noqa: E402
from cpu_codecs import Ref, load_bmp_gray  # noqa: E402
from nblic_image_compression_b200.synth import gen  # noqa: E402

KODAK = "/root/reference/img_kodak"


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def encode(ref, img, effort, near):
    if effort == 0:
        s = ref.q_encode(img)
        return s, img
    s, rec, _, _ = ref.n_encode(img, near, effort)
    return s, rec


def main():
    ref = Ref()
    man = {"generator": "tests/golden/make_golden.py", "reference": "WangXuan95/NBLIC-Image-Compression src/*.c, gcc -O3",
           "kodak": {}, "synthetic": {}, "edge": {}}
    os.makedirs(os.path.join(HERE, "kodak_e1n0"), exist_ok=True)
    full = "--quick" not in sys.argv
    for k in range(1, 25):
        name = f"{k:02d}"
        img = load_bmp_gray(os.path.join(KODAK, name + ".bmp"))
        ent = {"shape": list(img.shape), "pixels_sha256": sha(img.tobytes()), "streams": {}}
        for effort in (0, 1, 2, 3):
            for near in ((0,) if effort == 0 else (0, 1, 2, 3)):
                if not full and effort >= 2 and k > 2:
                    continue
                s, rec = encode(ref, img, effort, near)
                e = {"bytes": len(s), "sha256": sha(s)}
                if near:
                    e["recon_sha256"] = sha(rec.tobytes())
                ent["streams"][f"e{effort}n{near}"] = e
                if effort == 1 and near == 0:
                    open(os.path.join(HERE, "kodak_e1n0", name + ".nblic"), "wb").write(s)
                if effort == 0 and k == 1:
                    open(os.path.join(HERE, "kodak_01_e0n0.nblic"), "wb").write(s)
        man["kodak"][name] = ent
        print("kodak", name, {k_: v["bytes"] for k_, v in ent["streams"].items()}, flush=True)

    synth_jobs = [(64, 64, 0, [(0, 0), (1, 0), (2, 2), (3, 0)]),
                  (200, 333, 7, [(0, 0), (1, 0), (1, 3), (2, 0), (2, 2), (3, 0), (3, 1)]),
                  (1024, 1024, 0, [(0, 0), (1, 0), (2, 2), (3, 0)]),
                  (1024, 1024, 1, [(1, 0)]),
                  (2048, 2048, 0, [(0, 0), (1, 0), (2, 2)])]
    if full:
        synth_jobs.append((4096, 4096, 0, [(0, 0), (1, 0), (3, 0)]))
    for h, w, seed, sets in synth_jobs:
        img = gen(h, w, seed)
        ent = {"pixels_sha256": sha(img.tobytes()), "streams": {}}
        for effort, near in sets:
            s, rec = encode(ref, img, effort, near)
            e = {"bytes": len(s), "sha256": sha(s)}
            if near:
                e["recon_sha256"] = sha(rec.tobytes())
            ent["streams"][f"e{effort}n{near}"] = e
        man["synthetic"][f"{h}x{w}_s{seed}"] = ent
        print("synthetic", h, w, seed, {k_: v["bytes"] for k_, v in ent["streams"].items()}, flush=True)

    for name, img in edge_cases():
        ent = {"shape": list(img.shape), "pixels_sha256": sha(img.tobytes()), "streams": {}}
        for effort, near in SETTINGS_EDGE:
            s, rec = encode(ref, img, effort, near)
            e = {"bytes": len(s), "sha256": sha(s)}
            if len(s) <= 96:
                e["hex"] = s.hex()
            if near:
                e["recon_sha256"] = sha(rec.tobytes())
            ent["streams"][f"e{effort}n{near}"] = e
        man["edge"][name] = ent
    print("edge cases:", len(man["edge"]))
    json.dump(man, open(os.path.join(HERE, "manifest.json"), "w"), indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
