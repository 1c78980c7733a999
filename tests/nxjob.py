"""A 2048-byte NX job descriptor (nx_gzip_crb_cpb_t, inc_nx/nxu.h:286-616) built and read from Python: shared by the
descriptor-level tests (test_dropin.py, test_job_vectors.py)."""
import ctypes as C


def _be32(v):
    return int(v).to_bytes(4, "big")


class Job:
    """A 2048-byte nx_gzip_crb_cpb_t (inc_nx/nxu.h:286-616) with direct or indirect DDEs."""

    def __init__(self, fc, src_parts, dst_cap, histlen_qw=0, subc=0, sfbt=0, rem_or_dhtlen=0, dht=b"", crc=0, adler=1, split_dst=False):
        raw = C.create_string_buffer(2048 + 2048)
        base = (C.addressof(raw) + 2047) & ~2047
        self.keep = [raw]
        self.buf = (C.c_uint8 * 2048).from_address(base)
        self.addr = base
        self.put(0, _be32(fc))
        self.put(8, (base + 240).to_bytes(8, "big"))
        self.srcs = [C.create_string_buffer(p, len(p)) for p in src_parts]
        self.dst_bufs = [C.create_string_buffer(dst_cap // 2 + 1), C.create_string_buffer(dst_cap - dst_cap // 2 - 1)] if split_dst and dst_cap > 2 \
            else [C.create_string_buffer(max(dst_cap, 1))]
        self.dst_caps = [dst_cap // 2 + 1, dst_cap - dst_cap // 2 - 1] if split_dst and dst_cap > 2 else [dst_cap]
        self.dde(16, [(C.addressof(b), len(p)) for b, p in zip(self.srcs, src_parts)])
        self.dde(32, [(C.addressof(b), n) for b, n in zip(self.dst_bufs, self.dst_caps)])
        self.put(256 + 0, _be32(adler))
        self.put(256 + 4, int(crc).to_bytes(4, "little"))
        self.put(256 + 8, _be32((histlen_qw & 0xfff) << 20 | (subc & 7)))
        self.put(256 + 12, _be32((sfbt & 0xf) << 16 | (rem_or_dhtlen & 0xffff)))
        self.put(256 + 16, dht[:288])

    def put(self, off, b):
        for i, x in enumerate(b):
            self.buf[off + i] = x

    def get(self, off, n):
        return bytes(self.buf[off:off + n])

    def dde(self, off, segs):
        if len(segs) == 1:
            self.put(off, _be32(0) + _be32(segs[0][1]) + segs[0][0].to_bytes(8, "big"))
            return
        lst = C.create_string_buffer(16 * len(segs))
        self.keep.append(lst)
        for i, (a, n) in enumerate(segs):
            C.memmove(C.addressof(lst) + 16 * i, _be32(0) + _be32(n) + a.to_bytes(8, "big"), 16)
        self.put(off, _be32(len(segs) << 8) + _be32(sum(n for _, n in segs)) + C.addressof(lst).to_bytes(8, "big"))

    # outputs
    def cc(self): return self.buf[240 + 2]
    def ce(self): return self.buf[240 + 3] >> 5
    def valid(self): return self.buf[240] >> 7
    def tpbc(self): return int.from_bytes(self.get(244, 4), "big")
    def out(self): return b"".join(b.raw[:n] for b, n in zip(self.dst_bufs, self.dst_caps))[: self.tpbc()]
    def crc(self): return int.from_bytes(self.get(256 + 388, 4), "little")
    def adler(self): return int.from_bytes(self.get(256 + 384, 4), "big")
    def w392(self): return int.from_bytes(self.get(256 + 392, 4), "big")
    def w396(self): return int.from_bytes(self.get(256 + 396, 4), "big")
    def spbc_decomp(self): return int.from_bytes(self.get(256 + 688, 4), "big")
    def spbc_comp(self, count): return int.from_bytes(self.get(256 + (1664 if count else 400), 4), "big")
