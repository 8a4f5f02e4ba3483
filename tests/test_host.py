"""CPU-only tests of the host side: the C-ABI library loads and exports what include/nblic_b200.h
declares, header parsing, the no-GPU failure mode (no fallback), and the multi-GPU shard planner
under a world_size-2 gloo group."""
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from nblic_image_compression_b200.build import build_library
    build_library()
    from nblic_image_compression_b200 import api
    return api.load_library()


def test_header_symbols_exported(lib):
    from nblic_image_compression_b200 import api
    text = open(os.path.join(ROOT, "include", "nblic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(nblic_b200_\w+|Q?NBLIC(?:de)?compress\w*)\s*\(", text))
    assert declared, "no declarations parsed"
    assert declared == set(api.SYMBOLS), declared ^ set(api.SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_product_and_test_builds(lib):
    """Two libraries from the same sources: the product carries no sequential kernels (it says so in its version string
    and its cubin holds no coder_kernel), the test build (-DNBLIC_B200_SEQUENTIAL) does; both export the same ABI."""
    import subprocess
    from nblic_image_compression_b200 import api
    from nblic_image_compression_b200.build import LIB, LIB_SEQ, build_library
    build_library(sequential=True)
    seq = api.load_library(sequential=True)
    assert b"test build" not in lib.nblic_b200_version() and b"test build" in seq.nblic_b200_version()
    for sym in api.SYMBOLS:
        assert getattr(seq, sym) is not None
    names = {path: subprocess.run(["cuobjdump", "-elf", path], capture_output=True, text=True).stdout for path in (LIB, LIB_SEQ)}
    if names[LIB]:  # cuobjdump present (it is part of the CUDA toolkit of this image)
        assert "12coder_kernelI" not in names[LIB] and "12coder_kernelI" in names[LIB_SEQ]  # mangled template name (not e1p_coder_kernel)
        for kernel in ("coop_nblic_kernel", "subwarp_decode_kernel", "coop_q_kernel", "e1p_coder_kernel", "qpipe_finish_kernel", "psort_scatter_kernel"):
            assert kernel in names[LIB], kernel


def test_batch_cli_usage():
    """nblic_batch (csrc/nblic_batch_cli.c) rejects a bad command line before touching CUDA."""
    import subprocess
    from nblic_image_compression_b200.build import CLI, build_library
    build_library()
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode != 0 and "usage:" in r.stderr and "-b<MiB per group>" in r.stderr
    r = subprocess.run([CLI, "-cq", "/tmp", "x.pgm"], capture_output=True, text=True)
    assert r.returncode != 0 and "unknown switch -q" in r.stderr


def _write_bmp(path, img, top_down=False):
    """8-bit palettised BMP, rows padded to 4 bytes; bottom-up unless top_down (negative height)."""
    import struct
    h, w = img.shape
    stride = (w + 3) & ~3
    rows = np.zeros((h, stride), np.uint8)
    rows[:, :w] = img if top_down else img[::-1]
    pal = b"".join(bytes((k, k, k, 0)) for k in range(256))
    off = 14 + 40 + 1024
    hdr = b"BM" + struct.pack("<IHHI", off + stride * h, 0, 0, off) + struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, 8, 0, stride * h, 2835, 2835, 256, 0)
    open(path, "wb").write(hdr + pal + rows.tobytes())


def test_batch_cli_pipeline_on_a_cpu_stub(tmp_path):
    """csrc/nblic_batch_cli.c end to end WITHOUT a GPU: the CLI is linked against tests/stubs/cli_stub_backend.c (the batch
    entry points answered by the parity oracle -- test infrastructure, the product links libnblic_b200.so) so that its own
    logic runs here: PGM (with comments) and BMP (bottom-up, top-down, padded rows) readers, grouping (-b0: one file per
    group, the three-stage thread pipeline over more groups than ring slots), writers, and the error exits."""
    import subprocess
    from cpu_codecs import Oracle
    from nblic_image_compression_b200.synth import gen
    exe = tmp_path / "nblic_batch_stub"
    subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-std=gnu99", "-fwrapv", "-ffp-contract=off", "-o", str(exe),
                    os.path.join(ROOT, "nblic_image_compression_b200", "csrc", "nblic_batch_cli.c"),
                    os.path.join(ROOT, "tests", "stubs", "cli_stub_backend.c"), os.path.join(ROOT, "oracle", "nblic_oracle.c"), "-lpthread"], check=True)
    orc = Oracle()
    imgs = [gen(37 + 5 * k, 50 + 3 * k, k) for k in range(7)]  # widths 50..68: BMP rows with 0..3 bytes of padding
    inputs = []
    for k, im in enumerate(imgs):
        path = tmp_path / f"in{k}.{'bmp' if k % 3 else 'pgm'}"
        if k % 3 == 0:
            path.write_bytes(b"P5\n# a comment\n%d %d\n255\n" % (im.shape[1], im.shape[0]) + im.tobytes())
        else:
            _write_bmp(path, im, top_down=k % 3 == 2)
        inputs.append(path)
    for switches, extra, near, effort in (("-cn0e0", ["-b0", "-v"], 0, 0), ("-cn2e1", [], 2, 1), ("-ce2", ["-b0"], 0, 2)):
        out = tmp_path / ("enc" + switches[1:]); out.mkdir()
        r = subprocess.run([str(exe), switches, *extra, str(out), *map(str, inputs)], check=True, capture_output=True, text=True)
        if "-v" in extra:
            assert "7 images in 7 groups" in r.stdout and "pipeline of read" in r.stdout, r.stdout
        streams = []
        for im, path in zip(imgs, inputs):
            got = (out / (path.stem + ".nblic")).read_bytes()
            exp, rec = (orc.q_encode(im), im) if effort == 0 else orc.n_encode(im, near, effort)[:2]
            assert got == exp, (switches, path.name)
            streams.append((out / (path.stem + ".nblic"), rec))
        dec = tmp_path / ("dec" + switches[1:]); dec.mkdir()
        subprocess.run([str(exe), "-d", *extra[:1], str(dec), *[str(p) for p, _ in streams]], check=True, stdout=subprocess.DEVNULL)
        for path, rec in streams:
            data = (dec / (path.stem + ".pgm")).read_bytes()
            head = b"P5\n%d %d\n255\n" % (rec.shape[1], rec.shape[0])
            assert data == head + rec.tobytes(), (switches, path.name)
    bad = tmp_path / "noise.pgm"
    bad.write_bytes(b"not an image at all")
    r = subprocess.run([str(exe), "-c", str(tmp_path), str(inputs[0]), str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "neither a binary PGM nor an 8-bit BMP" in r.stderr
    r = subprocess.run([str(exe), "-c", str(tmp_path), str(tmp_path / "missing.pgm")], capture_output=True, text=True)
    assert r.returncode != 0 and "open" in r.stderr
    r = subprocess.run([str(exe), "-d", str(tmp_path), str(inputs[0])], capture_output=True, text=True)
    assert r.returncode != 0 and "is not a .nblic stream" in r.stderr


def test_dropin_wrappers_on_a_cpu_stub(tmp_path):
    """csrc/nblic_dropin.c (the reference's five entry points) WITHOUT a GPU: built into a shared object with the stub
    backend, called through ctypes, compared with the oracle: header emitted before validation, near / effort clipped in
    place, near > 0 leaves the reconstruction in the image buffer, decode from an exact-size buffer without a length hint,
    with a hint, foreign magics rejected without output."""
    import ctypes as C
    import subprocess
    from cpu_codecs import Oracle
    from nblic_image_compression_b200.synth import gen
    so = tmp_path / "libdropin_stub.so"
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-std=gnu99", "-fwrapv", "-ffp-contract=off", "-o", str(so),
                    os.path.join(ROOT, "nblic_image_compression_b200", "csrc", "nblic_dropin.c"),
                    os.path.join(ROOT, "tests", "stubs", "cli_stub_backend.c"), os.path.join(ROOT, "oracle", "nblic_oracle.c"), "-lpthread"], check=True)
    lib = C.CDLL(str(so))
    u8p, u16p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_int)
    lib.nblic_b200_hint_input_len.argtypes = [C.c_size_t]
    orc = Oracle()
    img = gen(45, 67, 5)
    h, w = img.shape
    for near_in, effort_in in ((0, 1), (2, 1), (1, 3), (17, 0), (3, 9)):  # the last two are clipped to (9, 1) and (3, 3)
        work = img.copy()
        out = np.zeros(2 * h * w + 8192, np.uint8)
        n_, e_ = C.c_int(near_in), C.c_int(effort_in)
        n = lib.NBLICcompress(0, out.ctypes.data_as(u8p), work.ctypes.data_as(u8p), h, w, C.byref(n_), C.byref(e_))
        near, effort = min(max(near_in, 0), 9), min(max(effort_in, 1), 3)
        assert (n_.value, e_.value) == (near, effort)
        exp, rec, _, _ = orc.n_encode(img, near, effort)
        assert n == len(exp) and bytes(out[:n]) == exp
        assert np.array_equal(work, rec)  # near > 0: the reconstruction; lossless: untouched
        for hint in (False, True):
            exact = np.frombuffer(exp, np.uint8).copy()  # exactly the stream, nothing behind it
            dec = np.zeros(h * w, np.uint8)
            hh, ww, nn, ee = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            if hint:
                lib.nblic_b200_hint_input_len(len(exp))
            rc = lib.NBLICdecompress(0, exact.ctypes.data_as(u8p), dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww), C.byref(nn), C.byref(ee))
            assert rc == 0 and (hh.value, ww.value, nn.value, ee.value) == (h, w, near, effort)
            assert np.array_equal(dec.reshape(h, w), rec)
    # the header is written before the dimensions are validated (NBLIC.c:768-775)
    out = np.zeros(64, np.uint8)
    n_, e_ = C.c_int(1), C.c_int(2)
    assert lib.NBLICcompress(0, out.ctypes.data_as(u8p), img.ctypes.data_as(u8p), 0, 5, C.byref(n_), C.byref(e_)) == -1
    assert bytes(out[:8]) == b"NBLIC0.3" and out[13] == 1 and out[15] == 2
    # QNBLIC: words, the -t alias, sniffing
    qexp = orc.q_encode(img)
    for fn in (lib.QNBLICcompress, lib.QNBLICcompressMultiThread):
        qout = np.zeros(h * w + 4096, np.uint16)
        work = img.copy()
        words = fn(qout.ctypes.data_as(u16p), work.ctypes.data_as(u8p), h, w)
        assert words * 2 == len(qexp) and qout[:words].tobytes() == qexp and np.array_equal(work, img)
    dec = np.zeros(h * w, np.uint8)
    hh, ww = C.c_int(), C.c_int()
    qbuf = np.frombuffer(qexp, np.uint16).copy()
    assert lib.QNBLICdecompress(qbuf.ctypes.data_as(u16p), dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww)) == 0
    assert (hh.value, ww.value) == (h, w) and np.array_equal(dec.reshape(h, w), img)
    hh, ww = C.c_int(-7), C.c_int(-7)
    nbuf = np.frombuffer(orc.n_encode(img, 0, 1)[0], np.uint8).copy()
    assert lib.QNBLICdecompress(nbuf.ctypes.data_as(u16p), dec.ctypes.data_as(u8p), C.byref(hh), C.byref(ww)) == -1  # the CLI's format sniff (NBLIC_main.c:223)
    assert (hh.value, ww.value) == (-7, -7)
    assert lib.QNBLICcompress(qbuf.ctypes.data_as(u16p), img.ctypes.data_as(u8p), 0, 9) == -1


def test_header_is_plain_c(tmp_path):
    """include/nblic_b200.h must be consumable by a C99 compiler (the reference and its CLI are C)."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "nblic_b200.h"\n'
                   "int probe(void) { int (*f)(int, unsigned char *, unsigned char *, int, int, int *, int *) = NBLICcompress;\n"
                   "  size_t (*g)(int, int) = nblic_b200_stream_bound; return f != 0 && g != 0 && NBLIC_MAX_IMG_SIZE == 100000000; }\n")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "use_header.o")], check=True)


def test_peek_and_bound(lib):
    from nblic_image_compression_b200 import api
    from conftest import GOLDEN
    q = open(os.path.join(GOLDEN, "kodak_01_e0n0.nblic"), "rb").read()
    n = open(os.path.join(GOLDEN, "kodak_e1n0", "04.nblic"), "rb").read()
    assert api.peek(q[:16]) == (512, 768, 0, 0)
    assert api.peek(n[:16]) == (768, 512, 0, 1)
    assert api.peek(b"NBLIC0.2" + bytes(8)) is None
    assert api.peek(b"Q0.2" + bytes(4)) is None  # zero dims (QNBLIC.c:33-45)
    bad = bytearray(n[:16]); bad[14] = 2
    assert api.peek(bytes(bad)) is None  # k_step below 3 (NBLIC.c:740)
    assert api.stream_bound(512, 768) >= 2 * 512 * 768


def test_legacy_decode_never_reads_past_the_callers_mapping(tmp_path):
    """The reference decode ABI carries no stream length (NBLIC.h:72); the drop-in wrappers bound their read by the
    readable extent of p_buf.  readable_prefix() is exercised on a buffer that ends at an inaccessible page."""
    import subprocess
    src = tmp_path / "probe.c"
    src.write_text(r"""
#define _GNU_SOURCE
#include <sys/mman.h>
#include "%s"
/* the batch ABI is not linked into this probe */
nblic_b200_ctx *nblic_b200_create(int d) { (void)d; return 0; }
const char *nblic_b200_last_error(const nblic_b200_ctx *c) { (void)c; return ""; }
size_t nblic_b200_stream_bound(int h, int w) { return 2 * (size_t)h * w + 8192; }
int nblic_b200_peek(const uint8_t *s, size_t n, int *h, int *w, int *a, int *e) { (void)s; (void)n; (void)h; (void)w; (void)a; (void)e; return -1; }
int nblic_b200_encode_batch(nblic_b200_ctx *c, int n, const uint8_t *const *i, const int *h, const int *w, int a, int e, uint8_t *const *o,
                            const size_t *oc, size_t *ol, uint8_t *const *r, int *s) { (void)c; (void)n; (void)i; (void)h; (void)w; (void)a; (void)e; (void)o; (void)oc; (void)ol; (void)r; (void)s; return -1; }
int nblic_b200_decode_batch(nblic_b200_ctx *c, int n, const uint8_t *const *st, const size_t *sl, uint8_t *const *im, const size_t *ic, int *h,
                            int *w, int *a, int *e, int *s) { (void)c; (void)n; (void)st; (void)sl; (void)im; (void)ic; (void)h; (void)w; (void)a; (void)e; (void)s; return -1; }
int main(void) {
    const size_t page = 4096, pages = 40;
    uint8_t *m = mmap(0, (pages + 1) * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    size_t got = 0, k;
    uint8_t *copy;
    if (m == MAP_FAILED) return 2;
    for (k = 0; k < pages * page; k++) m[k] = (uint8_t)(k * 7 + 3);
    if (mprotect(m + pages * page, page, PROT_NONE)) return 3;
    copy = readable_prefix(m + 100, 10 * pages * page, &got);          /* asks for far more than is readable */
    if (!copy || got != pages * page - 100 || memcmp(copy, m + 100, got)) return 4;
    free(copy);
    copy = readable_prefix(m + 5, 70000, &got);                          /* fully readable request */
    if (!copy || got != 70000 || memcmp(copy, m + 5, got)) return 5;
    free(copy);
    copy = readable_prefix(m + pages * page, 64, &got);                  /* nothing readable */
    if (!copy || got != 0) return 6;
    free(copy);
    return 0;
}
""" % os.path.join(ROOT, "nblic_image_compression_b200", "csrc", "nblic_dropin.c"))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-O1", "-std=gnu99", "-Wall", "-o", str(exe), str(src), "-lpthread"], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_no_gpu_means_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nblic_image_compression_b200 import api
    with pytest.raises(RuntimeError):
        api.Codec(0)
    img = np.zeros((4, 4), np.uint8)
    assert api.legacy.qnblic_compress(img) is None
    s, _, near, effort = api.legacy.nblic_compress(img, 50, 0)
    assert s is None and (near, effort) == (9, 1)  # clipped in place even when the call fails (NBLIC.c:768-770)


def test_sniff_rejects_foreign_streams_without_cuda(lib):
    from nblic_image_compression_b200 import api
    from conftest import GOLDEN
    n = open(os.path.join(GOLDEN, "kodak_e1n0", "01.nblic"), "rb").read()
    assert api.legacy.qnblic_decompress(n) is None  # NBLIC_main.c:223 tries QNBLIC first on every file
    assert api.legacy.nblic_decompress(b"Q0.2" + bytes(64)) is None


def test_bench_reference_arm_emits_the_contract_line():
    """bench.py --impl reference needs no GPU: one JSON line with the base contract's keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()[-1]
    line = json.loads(out)
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1


def test_plan_shards_properties():
    from nblic_image_compression_b200.shard import plan_shards
    rng = np.random.default_rng(0)
    px = rng.integers(1, 5_000_000, size=101).tolist()
    for world in (1, 2, 4, 8):
        shards = plan_shards(px, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(px)))
        loads = [sum(px[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(px)
    assert plan_shards([], 4) == [[], [], [], []]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from nblic_image_compression_b200.shard import gather_streams, plan_shards
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    px = [(i * 7919) % 1000 + 1 for i in range(37)]
    mine = plan_shards(px, world)[rank]
    streams = [f"stream-{i}-{px[i]}".encode() for i in mine]  # stands in for the per-image .nblic bytes
    full = gather_streams(streams, mine, len(px), dist)
    q.put((rank, mine, full == [f"stream-{i}-{px[i]}".encode() for i in range(len(px))]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, _, ok in res)
    assert sorted(res[0][1] + res[1][1]) == list(range(37)) and not set(res[0][1]) & set(res[1][1])
