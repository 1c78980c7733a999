#!/usr/bin/env python3
"""Regenerates tests/golden/* from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

  alice29.txt.gz        samples/alice29.txt (the reference's fixture, BASELINE.json config 1),
                        stored gzip-compressed; pinned by sha256 = oct/snappy-alice29.source:3
  kat_crc32.json        the 141 known-answer cases of test/test_crc32.c:38-180
  kat_adler32.json      the 140 known-answer cases of test/test_adler32.c:38-179
                        every expected value is re-computed by the reference's own
                        nx_crc32()/nx_adler32() (oracle/_ref/libnxz_ref.so) and must agree
  scp_stream.json       the 611-byte zlib stream of test/test_buf_error.c:107 (bug #74)
  ref_vectors.json      sizes/checksums the reference's software path (sw_zlib.c -> zlib 1.3)
                        produces: alice29 compress2 L1/L6/L9, makedata crc32 pins
"""
import ctypes, gzip, hashlib, json, os, re, subprocess, sys, zlib

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def c_string(tok: str) -> bytes:
    # decode a C string literal body (the tests only use printable ASCII and \\ escapes)
    return bytes(tok, "latin1").decode("unicode_escape").encode("latin1")


def parse_kats(path):
    src = open(path, encoding="latin1").read()
    body = src[src.index("tests[] = {"):]
    body = body[:body.index("};")]
    out = []
    for m in re.finditer(r"\{__LINE__,\s*(0x[0-9a-fA-F]+|\d+)L?,\s*(0x0|\(Byte \*\)\s*\"((?:[^\"\\]|\\.)*)\"),\s*(\d+),\s*(0x[0-9a-fA-F]+|\d+)L?\}", body):
        seed = int(m.group(1), 0)
        null = m.group(2) == "0x0"
        data = b"" if null else c_string(m.group(3))
        out.append({"seed": seed, "null": null, "data_hex": data.hex(), "len": int(m.group(4)), "expect": int(m.group(5), 0)})
    return out


def main():
    os.environ.setdefault("NX_GZIP_LOGFILE", "/tmp/nx.log")
    os.environ["NX_GZIP_TYPE_SELECTOR"] = "1"          # software path of the reference (sw_zlib.c)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    ref = ctypes.CDLL(os.path.join(ROOT, "oracle/_ref/libnxz_ref.so"))
    ref.nx_crc32.restype = ctypes.c_ulong
    ref.nx_crc32.argtypes = [ctypes.c_ulong, ctypes.c_char_p, ctypes.c_size_t]
    ref.nx_adler32.restype = ctypes.c_ulong
    ref.nx_adler32.argtypes = [ctypes.c_ulong, ctypes.c_char_p, ctypes.c_size_t]

    alice = open(f"{REF}/samples/alice29.txt", "rb").read()
    assert hashlib.sha256(alice).hexdigest().startswith("7467306e")
    with open(os.path.join(HERE, "alice29.txt.gz"), "wb") as f:
        f.write(gzip.compress(alice, 9, mtime=0))

    for name, fn in (("crc32", ref.nx_crc32), ("adler32", ref.nx_adler32)):
        kats = parse_kats(f"{REF}/test/test_{name}.c")
        for k in kats:
            buf = None if k["null"] else bytes.fromhex(k["data_hex"])
            got = fn(k["seed"], buf, k["len"]) & 0xffffffff
            assert got == k["expect"], (name, k, hex(got))
        print(name, len(kats), "KATs agree with the reference build")
        json.dump(kats, open(os.path.join(HERE, f"kat_{name}.json"), "w"), indent=0)

    src = open(f"{REF}/test/test_buf_error.c", encoding="latin1").read()
    body = src[src.index("compr[611]"):]
    body = body[body.index("{") + 1:body.index("};")]
    scp = bytes(int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+", body))
    assert len(scp) == 611
    json.dump({"zlib_stream_hex": scp.hex()}, open(os.path.join(HERE, "scp_stream.json"), "w"))

    # the reference's own compress2()/uncompress() through sw_zlib.c
    ref.compress2.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_ulong), ctypes.c_char_p, ctypes.c_ulong, ctypes.c_int]
    ref.uncompress.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_ulong), ctypes.c_char_p, ctypes.c_ulong]
    vec = {"zlib_version": zlib.ZLIB_RUNTIME_VERSION, "alice29": {"len": len(alice),
           "crc32": ref.nx_crc32(0, alice, len(alice)) & 0xffffffff, "adler32": ref.nx_adler32(1, alice, len(alice)) & 0xffffffff}}
    for lv in (1, 6, 9):
        cap = ctypes.c_ulong(len(alice) * 2)
        out = ctypes.create_string_buffer(cap.value)
        assert ref.compress2(out, ctypes.byref(cap), alice, len(alice), lv) == 0
        back = ctypes.create_string_buffer(len(alice))
        bl = ctypes.c_ulong(len(alice))
        assert ref.uncompress(back, ctypes.byref(bl), out.raw[:cap.value], cap.value) == 0 and back.raw == alice
        vec["alice29"][f"compress2_L{lv}"] = cap.value
    md = {}
    for seed, b in ((1, 20), (1, 24), (4, 20), (5, 20)):
        p = subprocess.run([os.path.join(ROOT, "oracle/_ref/makedata"), "-s", str(seed), "-b", str(b)],
                           input=alice, capture_output=True, check=True)
        md[f"s{seed}_b{b}"] = {"len": len(p.stdout), "crc32": zlib.crc32(p.stdout), "zlib_L1": len(zlib.compress(p.stdout, 1)),
                               "zlib_L6": len(zlib.compress(p.stdout, 6))}
    md["s1_b26"] = {"len": 1 << 26, "crc32": 0xece3d95e}
    vec["makedata"] = md
    json.dump(vec, open(os.path.join(HERE, "ref_vectors.json"), "w"), indent=1)
    print("fixtures written")


if __name__ == "__main__":
    main()
