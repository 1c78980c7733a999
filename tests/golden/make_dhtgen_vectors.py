#!/usr/bin/env python3
"""Regenerates tests/golden/dhtgen_vectors.json (build container only: needs oracle/_ref/dhtgen_test,
the reference's own lib/nx_dhtgen.c compiled with -D_DHTGEN_TEST by oracle/Makefile).

For a handful of LZ77 histograms (286 lit/len + 30 distance counts, zero counts raised to 1 exactly as
lib/nx_dht.c:627 fill_zero_lzcounts(.., 1) does before it calls dhtgen) the file holds the dynamic
Huffman table the REFERENCE generates: the bytes of cpb.in_dht and their length in bits.  The GPU's
nxgpu_dhtgen must produce a valid table for the same counts that costs no more bits."""
import gzip, json, os, random, subprocess, tempfile, zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TOOL = os.path.join(ROOT, "oracle", "_ref", "dhtgen_test")

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577]


def code_of(v, base):
    k = 0
    while k + 1 < len(base) and base[k + 1] <= v:
        k += 1
    return k


def histogram(data):
    """greedy LZ77 with a 3-byte hash of last positions (any reasonable tokenisation will do)"""
    ll, dd = [0] * 286, [0] * 30
    last, i, n = {}, 0, len(data)
    while i < n:
        best = 0
        if i + 3 <= n:
            key = data[i:i + 3]
            j = last.get(key)
            if j is not None and i - j <= 32768:
                l = 3
                while l < 258 and i + l < n and data[j + l] == data[i + l]:
                    l += 1
                best, dist = l, i - j
            last[key] = i
        if best >= 3:
            ll[257 + code_of(best, LEN_BASE)] += 1
            dd[code_of(dist, DIST_BASE)] += 1
            for k in range(i + 1, min(i + best, n - 2)):
                last[data[k:k + 3]] = k
            i += best
        else:
            ll[data[i]] += 1
            i += 1
    ll[256] = 1
    return ll + dd


def reference_dht(counts):
    with tempfile.TemporaryDirectory() as td:
        lz = os.path.join(td, "lz.txt")
        with open(lz, "w") as f:
            for s in range(286):
                f.write(f"{s} : {counts[s]}\n")
            for s in range(30):
                f.write(f"{s} : {counts[286 + s]}\n")
        subprocess.run([TOOL, lz, os.path.join(td, "dht.bin")], cwd=td, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        blob = open(os.path.join(td, "dht.bin"), "rb").read()
    bits = ((blob[14] & 0x0f) << 8) | blob[15]
    return blob[16:16 + (bits + 7) // 8], bits


def main():
    alice = gzip.decompress(open(os.path.join(HERE, "alice29.txt.gz"), "rb").read())
    rnd = random.Random(5)
    inputs = {
        "alice_64k": alice[:65536],
        "alice_tail": alice[100000:],
        "binaryish": bytes(rnd.choice(b"\x00\x01\x02\xff\x10 ") for _ in range(40000)),
        "random_8k": rnd.randbytes(8192),
        "zeros": bytes(30000),
        "abc": b"abc" * 5000 + alice[:3000],
    }
    out = []
    for name, data in inputs.items():
        c = histogram(data)
        c = [x if x else 1 for x in c]                      # lib/nx_dht.c:627
        dht, bits = reference_dht(c)
        out.append({"name": name, "counts": c, "ref_dht_hex": dht.hex(), "ref_bits": bits})
        print(name, "reference table:", bits, "bits")
    json.dump({"generator": "tests/golden/make_dhtgen_vectors.py over oracle/_ref/dhtgen_test (reference lib/nx_dhtgen.c, -D_DHTGEN_TEST)",
               "vectors": out}, open(os.path.join(HERE, "dhtgen_vectors.json"), "w"))


if __name__ == "__main__":
    main()
