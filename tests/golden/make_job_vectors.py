#!/usr/bin/env python3
"""Regenerates tests/golden/job_vectors.json.gz: per-function-code descriptor in/out dumps for the NX decompress
function codes (SURVEY.md §8c "golden vectors: per-FC descriptor in/out dumps").

INDEPENDENT of oracle/ and of the CUDA engine on purpose.  The expected CSB/CPB fields are derived here from

  * RFC 1951 (the walker below records every symbol boundary of a stream made by system zlib — the library the
    reference's software path dlopens, lib/sw_zlib.c:283-327 — and checks its own output against zlib's), and
  * the NX-GZIP user manual (doc/power_nx_gzip_um.pdf) §2.4 "SPBC indicates the number of compressed source bytes
    read by the accelerator; SUBC indicates the number of source bits that the accelerator discarded, because they
    were past the stream end", §5.2.5.5 Table 5-3 (SFBT / SUBC combinations), §5.2.5.6 (forward progress), and the
    field comments of inc_nx/nxu.h:296-540 ("For ZLIB and GZIP these values are 32 and 64 bits ... may range from
    32 to 39, and 64 to 71 bits", :454-465),
  * the way the reference's host code feeds a suspended job back (lib/nx_inflate.c:1447-1609): next source byte =
    spbc - histlen - (subc+7)/8, in_subc = subc % 8, in_sfbt / in_rembytecnt / in_dht copied from the outputs.

Both engines (oracle/nxemu.c on the CPU, libnxgpu.so on the B200) are checked against THESE vectors
(tests/test_job_vectors.py), not against each other.

The one free parameter of the model is how far the engine has read behind the point where it stopped by itself
(final end-of-block, single-block suspend): READ_AHEAD = 8 bytes, the smallest value that reproduces the manual's
SUBC ranges for a bare zlib (32..39) and gzip (64..71) trailer.
"""
import base64
import gzip
import json
import os
import random
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
READ_AHEAD = 8

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
             6145, 8193, 12289, 16385, 24577]
DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]
CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


class Bits:
    def __init__(self, data):
        self.d = data
        self.pos = 0

    def get(self, n):
        v = 0
        for i in range(n):
            p = self.pos + i
            v |= ((self.d[p >> 3] >> (p & 7)) & 1) << i
        self.pos += n
        return v


def canonical(lengths):
    """{(length, code): symbol} per RFC 1951 §3.2.2"""
    count = [0] * 16
    for l in lengths:
        count[l] += 1
    count[0] = 0
    code, nxt = 0, [0] * 16
    for b in range(1, 16):
        code = (code + count[b - 1]) << 1
        nxt[b] = code
    table = {}
    for sym, l in enumerate(lengths):
        if l:
            table[(l, nxt[l])] = sym
            nxt[l] += 1
    return table


def decode(bits, table):
    code = 0
    for l in range(1, 16):
        code = (code << 1) | bits.get(1)
        if (l, code) in table:
            return table[(l, code)]
    raise ValueError("no such code")


def walk(stream):
    """Every block of a raw deflate stream with every symbol boundary: the ground truth the vectors are cut from."""
    bits = Bits(stream + bytes(8))
    out = bytearray()
    blocks = []
    while True:
        b = {"hdr": bits.pos, "out0": len(out)}
        b["bfinal"] = bits.get(1)
        b["btype"] = bits.get(2)
        if b["btype"] == 0:
            bits.pos = (bits.pos + 7) & ~7
            ln, nl = bits.get(16), bits.get(16)
            assert ln ^ nl == 0xffff
            b["data"] = bits.pos
            b["len"] = ln
            out += stream[bits.pos >> 3: (bits.pos >> 3) + ln]
            bits.pos += 8 * ln
            b["end"] = bits.pos
        else:
            if b["btype"] == 1:
                ll = canonical([8] * 144 + [9] * 112 + [7] * 24 + [8] * 8)
                dd = canonical([5] * 30)
            else:
                b["dht_from"] = bits.pos                       # HLIT is the first bit the engine hands back (in_dht)
                hlit, hdist, hclen = bits.get(5) + 257, bits.get(5) + 1, bits.get(4) + 4
                cl = [0] * 19
                for i in range(hclen):
                    cl[CL_ORDER[i]] = bits.get(3)
                clt = canonical(cl)
                lens = []
                while len(lens) < hlit + hdist:
                    s = decode(bits, clt)
                    if s < 16:
                        lens.append(s)
                    elif s == 16:
                        lens += [lens[-1]] * (3 + bits.get(2))
                    elif s == 17:
                        lens += [0] * (3 + bits.get(3))
                    else:
                        lens += [0] * (11 + bits.get(7))
                assert len(lens) == hlit + hdist
                b["dht_bits"] = bits.pos - b["dht_from"]
                ll, dd = canonical(lens[:hlit]), canonical(lens[hlit:])
            b["data"] = bits.pos
            syms = []                                          # (bit position behind the symbol, output length behind it)
            while True:
                s = decode(bits, ll)
                if s == 256:
                    break
                if s < 256:
                    out.append(s)
                else:
                    ln = LEN_BASE[s - 257] + bits.get(LEN_EXTRA[s - 257])
                    ds = decode(bits, dd)
                    dist = DIST_BASE[ds] + bits.get(DIST_EXTRA[ds])
                    for _ in range(ln):
                        out.append(out[-dist])
                syms.append((bits.pos, len(out)))
            b["syms"] = syms
            b["end"] = bits.pos
        b["out1"] = len(out)
        blocks.append(b)
        if b["bfinal"]:
            return blocks, bytes(out)


def bitslice(stream, frm, n):
    v = 0
    for i in range(n):
        p = frm + i
        v |= ((stream[p >> 3] >> (p & 7)) & 1) << i
    return v.to_bytes(288, "little")


def run_job(stream, blocks, start_bit, avail_bytes_end, out_done, single_block, state):
    """One decompress job over stream bits [start_bit, 8*avail_bytes_end): where does the engine stop, and what does it
    report (manual Table 5-3)?  `state` = ("hdr", block) | ("stored", block, remaining) | ("huff", block).
    Returns dict(tpbc, sfbt, subc, rem, dht_from, dht_bits, stop_bit, state, src_read_end)."""
    E = 8 * avail_bytes_end
    pos = start_bit
    out = out_done
    bi = state[1]
    r = {"rem": 0, "dht_bits": 0, "dht_from": 0}
    first = True
    while True:
        b = blocks[bi]
        if state[0] == "hdr":
            if not first and (single_block or pos == E):
                # a BFINAL=0 block has just ended: single-block suspend, or the source ends exactly here (SFBT 1110)
                rd = min(avail_bytes_end, ((pos + 7) >> 3) + READ_AHEAD) if pos != E else avail_bytes_end
                r.update(sfbt=0xe, subc=8 * rd - pos, stop=pos, state=("hdr", bi), read_end=rd)
                break
            if E < b["data"]:
                # the header is incomplete: all of it is handed back; BFINAL is reported when its bit was there
                f = b["bfinal"] if E > b["hdr"] else 0
                r.update(sfbt=0xe | f, subc=E - b["hdr"], stop=b["hdr"], state=("hdr", bi), read_end=avail_bytes_end)
                break
            pos = b["data"]
            state = ("stored", bi, b["len"]) if b["btype"] == 0 else ("huff", bi)
        first = False
        if state[0] == "stored":
            rem = state[2]
            n = min(rem, (E - pos) // 8)
            pos += 8 * n
            out += n
            rem -= n
            if rem:
                r.update(sfbt=0x8 | b["bfinal"], subc=0, rem=rem, stop=pos, state=("stored", bi, rem), read_end=avail_bytes_end)
                break
        else:
            syms = [s for s in b["syms"] if s[0] > pos]
            done = [s for s in syms if s[0] <= E]
            if done:
                pos, out = done[-1]
            if b["end"] > E:
                # the next symbol (or the end-of-block code) is not all there
                r.update(sfbt=(0xa if b["btype"] == 1 else 0xc) | b["bfinal"], subc=E - pos, stop=pos, state=("huff", bi),
                         read_end=avail_bytes_end)
                if b["btype"] == 2:
                    r.update(dht_from=b["dht_from"], dht_bits=b["dht_bits"])
                break
            pos = b["end"]
        # the block is complete
        if b["bfinal"]:
            rd = min(avail_bytes_end, ((pos + 7) >> 3) + READ_AHEAD)
            r.update(sfbt=0, subc=8 * rd - pos, stop=pos, state=("final",), read_end=rd)
            break
        bi += 1
        state = ("hdr", bi)
    r["tpbc"] = out - out_done
    return r


def b64(b):
    return base64.b64encode(b).decode()


def main():
    alice = gzip.decompress(open(os.path.join(HERE, "alice29.txt.gz"), "rb").read())
    rnd = random.Random(2024)
    data = alice[:24000] + bytes(3000) + rnd.randbytes(2000) + alice[1000:9000]
    streams = {
        "dyn6": zlib.compress(data, 6)[2:-4],
        "dyn1": zlib.compress(data, 1)[2:-4],
        "stored": zlib.compress(data, 0)[2:-4],
    }
    fx = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
    streams["fixed"] = fx.compress(data) + fx.flush()
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    streams["multiblock"] = b"".join(co.compress(data[i:i + 5000]) + co.flush(zlib.Z_FULL_FLUSH if (i // 5000) % 2 else zlib.Z_SYNC_FLUSH)
                                     for i in range(0, len(data), 5000)) + co.flush()
    jobs = []
    chain_id = 0
    for name, stream in streams.items():
        blocks, out = walk(stream)
        assert out == data == zlib.decompress(stream, -15), name
        plans = []
        for trailing in (0, 4, 8, 13, 9000, 70000):
            plans.append(("oneshot", [len(stream) + trailing], trailing, False))
        plans.append(("p997", [997], 5, False))
        plans.append(("p61", [61] * 30 + [4099], 8, False))
        plans.append(("single", [len(stream) + 100], 100, True))
        for plan, pieces, trailing, single in plans:
            tail = bytes((i * 131 + 7) & 0xff for i in range(trailing))     # the test rebuilds the same bytes
            full = stream + tail
            byte_pos, start_bit, out_done = 0, 0, 0
            state = ("hdr", 0)
            in_sfbt = in_rem = in_dhtlen = 0
            in_dht = b""
            crc, adler = 0, 1
            step, extra, k = 0, 0, 0
            while True:
                piece = (pieces[k] if k < len(pieces) else pieces[-1]) + extra
                end = min(len(full), byte_pos + piece)
                first = step == 0
                hist = min(out_done, 32768)
                r = run_job(full, blocks, start_bit, end, out_done, single, state)
                exp_out = data[out_done: out_done + r["tpbc"]]
                crc, adler = zlib.crc32(exp_out, crc), zlib.adler32(exp_out, adler)
                hist_padded = (hist + 15) // 16 * 16
                job = {
                    "stream": name, "plan": plan, "chain": chain_id, "step": step, "trailing": trailing,
                    "fc": (0x12 if first else 0x16) if single else (0x10 if first else 0x14),
                    "src_from": byte_pos, "src_to": end, "hist": hist, "out_done": out_done,
                    "in_subc": (8 - (start_bit - 8 * byte_pos)) % 8, "in_sfbt": in_sfbt, "in_rem": in_rem, "in_dhtlen": in_dhtlen,
                    "in_dht": b64(in_dht),
                    "exp": {"cc": 3, "tpbc": r["tpbc"], "sfbt": r["sfbt"], "subc": r["subc"], "rem": r["rem"],
                            "dhtlen": r["dht_bits"], "spbc": hist_padded + (r["read_end"] - byte_pos),
                            "crc": crc, "adler": adler},
                }
                if r["dht_bits"]:
                    job["exp"]["dht"] = b64(bitslice(full, r["dht_from"], r["dht_bits"])[: (r["dht_bits"] + 7) // 8])
                jobs.append(job)
                assert r["subc"] <= 0xffff
                out_done += r["tpbc"]
                step += 1
                if r["state"][0] == "final":
                    assert out_done == len(data)
                    # the host finds the end of the stream at spbc - histlen - subc/8 (lib/nx_inflate.c:1452-1472)
                    assert (r["read_end"] - byte_pos) - r["subc"] // 8 + byte_pos == (r["stop"] + 7) // 8
                    break
                # feed back like lib/nx_inflate.c:1480-1609
                consumed = (r["read_end"] - byte_pos) - (r["subc"] + 7) // 8
                progress = consumed > 0 or r["tpbc"] > 0
                new_pos = byte_pos + consumed
                assert 8 * new_pos <= r["stop"] < 8 * new_pos + 8 or r["stop"] == 8 * new_pos
                start_bit = r["stop"]
                byte_pos = new_pos
                state = r["state"]
                in_sfbt = r["sfbt"]
                in_rem = r["rem"]
                in_dhtlen = r["dht_bits"]
                in_dht = bitslice(full, r["dht_from"], r["dht_bits"]) if r["dht_bits"] else b""
                extra = extra + piece if not progress else 0   # no forward progress (manual §5.2.5.6): give more source
                if progress:
                    k += 1
                assert step < 5000
            chain_id += 1
    doc = {
        "generator": "tests/golden/make_job_vectors.py (RFC 1951 walker + NX-GZIP manual Table 5-3; independent of oracle/ and csrc/)",
        "read_ahead_bytes": READ_AHEAD,
        "data_zlib_b64": b64(zlib.compress(data, 9)),
        "streams": {k: b64(v) for k, v in streams.items()},
        "jobs": jobs,
    }
    path = os.path.join(HERE, "job_vectors.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(doc, separators=(",", ":")).encode())
    print(f"{len(jobs)} jobs in {chain_id} chains -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
