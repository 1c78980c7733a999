"""Inputs of the compression-ratio gate (tests/test_gpu_parity.py::test_deflate_ratio_gate_wide, tools/ratio_fixtures.py):
besides the makedata / alice29 fixtures, data where 3- and 4-byte matches carry the ratio — the LZ77 stage here only emits
matches of 5 bytes and more, so these are the cases that would show it."""
import gzip
import os
import random
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ratio_inputs(pg):
    alice = gzip.decompress(open(os.path.join(ROOT, "tests", "golden", "alice29.txt.gz"), "rb").read())
    rnd = random.Random(1)
    out = [("alice29", alice)]
    for seed in (1, 4, 5):
        out.append((f"makedata-s{seed}", pg.makedata(seed, 20, alice)))
    # source code: this repository's own largest kernel source, repeated to 1 MiB with line numbers changed
    src = open(os.path.join(ROOT, "power-gzip_b200", "csrc", "deflate.cu"), "rb").read()
    out.append(("source-code", (src + open(os.path.join(ROOT, "bench.py"), "rb").read() + open(os.path.join(ROOT, "power-gzip_b200", "csrc", "nxgpu_api.cu"), "rb").read())[:400000]))
    # the 33-symbol alphabet of the reference's random test data (test/test_utils.c:22-28), srand(1)-like fixed seed
    dict33 = b"abcdefghijklmnopqrstuvwxyz,.!?.{}"
    out.append(("alphabet33", bytes(rnd.choice(dict33) for _ in range(300000))))
    # binary records: little-endian structs with slowly varying fields (3- and 4-byte matches at fixed strides)
    recs = bytearray()
    t, x = 1_700_000_000, 1000.0
    for i in range(20000):
        t += rnd.choice((1, 1, 1, 2, 5)); x += rnd.uniform(-1, 1)
        recs += struct.pack("<IHHfB3x", t, i & 0xffff, rnd.randrange(8), x, rnd.randrange(4))
    out.append(("binary-records", bytes(recs)))
    # short-period data: periods 1..7 in runs of a few hundred bytes, separated by noise
    sp = bytearray()
    while len(sp) < 300000:
        p = rnd.randrange(1, 8)
        unit = bytes(rnd.randrange(256) for _ in range(p))
        sp += unit * rnd.randrange(20, 120) + bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 6)))
    out.append(("short-periods", bytes(sp[:300000])))
    # words from a small vocabulary: matches are mostly 3 to 6 bytes
    vocab = [bytes(rnd.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rnd.randrange(2, 7))) for _ in range(300)]
    out.append(("small-vocabulary", b" ".join(rnd.choice(vocab) for _ in range(60000))))
    # DNA-like: four symbols
    out.append(("four-symbols", bytes(rnd.choice(b"ACGT") for _ in range(300000))))
    return out
