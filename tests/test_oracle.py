"""CPU tests: the oracle (oracle/nblic_oracle.c) against the reference's golden vectors and,
when oracle/_ref is present, against the reference itself.  No GPU needed."""
import numpy as np
import pytest

from cases import SETTINGS_EDGE, edge_cases
from conftest import sha
from nblic_image_compression_b200.synth import gen


def _enc(codec, img, effort, near):
    if effort == 0:
        return codec.q_encode(img), img
    s, rec, _, _ = codec.n_encode(img, near, effort)
    return s, rec


def _dec(codec, data):
    r = codec.q_decode(data)
    if r is not None:
        return r
    return codec.n_decode(data)[0]


def test_synth_generator_hashes(manifest):
    assert sha(gen(1024, 1024, 0).tobytes())[:16] == "a04f5e66bcd9927f"
    for key in ("64x64_s0", "200x333_s7", "1024x1024_s0"):
        h, w, s = int(key.split("x")[0]), int(key.split("x")[1].split("_")[0]), int(key.split("_s")[1])
        assert sha(gen(h, w, s).tobytes()) == manifest["synthetic"][key]["pixels_sha256"]


def test_kodak_e0_e1_all(oracle, kodak, manifest):
    total = {"e0n0": 0, "e1n0": 0}
    for name, img in kodak.items():
        for key, (effort, near) in {"e0n0": (0, 0), "e1n0": (1, 0)}.items():
            s, _ = _enc(oracle, img, effort, near)
            ent = manifest["kodak"][name]["streams"][key]
            assert len(s) == ent["bytes"] and sha(s) == ent["sha256"], (name, key)
            total[key] += len(s)
            assert np.array_equal(_dec(oracle, s), img)
    assert total == {"e0n0": 4985986, "e1n0": 4891174}  # BASELINE.md section 4 sums


@pytest.mark.parametrize("name,key", [("01", "e1n2"), ("05", "e1n1"), ("23", "e1n3"), ("01", "e2n0"), ("23", "e2n2"),
                                      ("04", "e2n3"), ("01", "e3n0"), ("09", "e3n2")])
def test_kodak_other_settings(oracle, kodak, manifest, name, key):
    effort, near = int(key[1]), int(key[3])
    s, rec = _enc(oracle, kodak[name], effort, near)
    ent = manifest["kodak"][name]["streams"][key]
    assert len(s) == ent["bytes"] and sha(s) == ent["sha256"]
    if near:
        assert sha(rec.tobytes()) == ent["recon_sha256"]
        assert int(np.abs(rec.astype(int) - kodak[name].astype(int)).max()) <= near
    assert np.array_equal(_dec(oracle, s), rec)


def test_config1_stream_file(oracle, kodak):
    import os
    from conftest import GOLDEN
    data = open(os.path.join(GOLDEN, "kodak_01_e0n0.nblic"), "rb").read()
    assert len(data) == 255738 and sha(data)[:16] == "e61745cecc6a9ef4"
    assert np.array_equal(oracle.q_decode(data), kodak["01"])
    assert oracle.q_encode(kodak["01"]) == data


def test_synthetic(oracle, manifest):
    for key in ("64x64_s0", "200x333_s7"):
        h, w, s = int(key.split("x")[0]), int(key.split("x")[1].split("_")[0]), int(key.split("_s")[1])
        img = gen(h, w, s)
        for skey, ent in manifest["synthetic"][key]["streams"].items():
            st, rec = _enc(oracle, img, int(skey[1]), int(skey[3]))
            assert len(st) == ent["bytes"] and sha(st) == ent["sha256"], (key, skey)
            assert np.array_equal(_dec(oracle, st), rec)
    img = gen(1024, 1024, 0)
    for skey in ("e0n0", "e1n0"):
        st, _ = _enc(oracle, img, int(skey[1]), 0)
        assert sha(st) == manifest["synthetic"]["1024x1024_s0"]["streams"][skey]["sha256"]


def test_edge_cases(oracle, manifest):
    for name, img in edge_cases():
        ent = manifest["edge"][name]
        assert sha(img.tobytes()) == ent["pixels_sha256"], name
        for effort, near in SETTINGS_EDGE:
            st, rec = _enc(oracle, img, effort, near)
            g = ent["streams"][f"e{effort}n{near}"]
            assert len(st) == g["bytes"] and sha(st) == g["sha256"], (name, effort, near)
            if "hex" in g:
                assert st.hex() == g["hex"]
            if near:
                assert sha(rec.tobytes()) == g["recon_sha256"]
                assert int(np.abs(rec.astype(int) - img.astype(int)).max()) <= near
            assert np.array_equal(_dec(oracle, st), rec), (name, effort, near)


def test_bad_input(oracle):
    assert oracle.q_decode(b"\0" * 64) is None
    assert oracle.n_decode(b"NBLIC0.2" + b"\0" * 64) is None
    hdr = bytearray(b"NBLIC0.3\x01\x00\x04\x00\x04\x00\x03\x01" + b"\0" * 32)
    hdr[14] = 2  # k_step below the minimum of 3 (NBLIC.c:740)
    assert oracle.n_decode(bytes(hdr)) is None
    hdr[14], hdr[15] = 3, 4  # effort out of range
    assert oracle.n_decode(bytes(hdr)) is None


def test_parameter_clipping(oracle):
    img = gen(16, 16, 3)
    s, _, near, effort = oracle.n_encode(img, 50, 9)  # clipped to 9 / 3 in place (NBLIC.c:768-770)
    assert (near, effort) == (9, 3) and s[13] == 9 and s[14] == 16 and s[15] == 3
    s, _, near, effort = oracle.n_encode(img, -4, 0)
    assert (near, effort) == (0, 1)


def test_live_against_reference(oracle, ref):
    rng = np.random.default_rng(1234)
    for trial in range(40):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 60))
        kind = trial % 4
        if kind == 0:
            img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        elif kind == 1:
            img = gen(h, w, trial)
        elif kind == 2:
            img = (rng.integers(0, 2, size=(h, w)) * 255).astype(np.uint8)
        else:
            img = np.clip(rng.normal(128, 20, size=(h, w)).cumsum(axis=1) / 8 + 64, 0, 255).astype(np.uint8)
        assert oracle.q_encode(img) == ref.q_encode(img)
        for effort in (1, 2, 3):
            near = int(rng.integers(0, 10))
            a, b = oracle.n_encode(img, near, effort), ref.n_encode(img, near, effort)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]), (trial, effort, near)
            d = ref.n_decode(a[0])
            assert np.array_equal(d[0], a[1])
