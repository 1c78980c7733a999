"""power-gzip_b200 — host-side mirror (ctypes) of the C-ABI in include/nxgpu.h.

The product is ``libnxgpu.so`` (hand-written sm_100a CUDA behind a C-ABI that
stands in for the POWER NX-GZIP engine under libnxz's unchanged host code,
reference lib/gzip_vas.c:281 ``nxu_run_job``).  This module only marshals
arguments; it contains no compression, decompression or checksum arithmetic and
has NO CPU fallback: if the shared library is missing, or no sm_100 GPU is
usable, every entry point raises.

Names follow the reference's zlib-compatible surface (libnxz.h:119-192):
``compress``/``uncompress``/``crc32``/``adler32`` plus the batch calls.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnxgpu.so")

MEM_HOST, MEM_DEVICE = 0, 1
WRAP_RAW, WRAP_ZLIB, WRAP_GZIP, WRAP_AUTO, WRAP_RAW_CONT = 0, 1, 2, 3, 5
STREAM_INDEPENDENT = 0x100
F_FINAL, F_FIXED, F_NO_JOINER = 1, 2, 4
E_NODEV, E_ARG, E_DATA, E_MEM, E_BUF = -100, -2, -3, -4, -5


class NxGpuError(RuntimeError):
    def __init__(self, rc: int, what: str):
        super().__init__(f"{what}: rc={rc} ({last_error()})")
        self.rc = rc


class CksumItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("len", C.c_uint64), ("crc_seed", C.c_uint32), ("adler_seed", C.c_uint32)]


class CksumResult(C.Structure):
    _fields_ = [("crc32", C.c_uint32), ("adler32", C.c_uint32)]


class DeflateItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_len", C.c_uint32), ("hist_len", C.c_uint32),
                ("dst", C.c_void_p), ("dst_cap", C.c_uint32), ("flags", C.c_uint32)]


class DeflateResult(C.Structure):
    _fields_ = [("rc", C.c_int32), ("out_len", C.c_uint32), ("tebc", C.c_uint32),
                ("crc32", C.c_uint32), ("adler32", C.c_uint32), ("n_tokens", C.c_uint32)]


class StreamResult(C.Structure):
    _fields_ = [("out_len", C.c_uint64), ("crc32", C.c_uint32), ("adler32", C.c_uint32),
                ("n_chunks", C.c_uint32), ("n_tokens", C.c_uint64)]


class TeamResult(C.Structure):
    _fields_ = [("out_len", C.c_uint64), ("crc32", C.c_uint32), ("adler32", C.c_uint32), ("src_len", C.c_uint64),
                ("my_offset", C.c_uint64), ("my_size", C.c_uint64), ("device_ms", C.c_float)]


class InflateItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_len", C.c_uint32), ("dst", C.c_void_p),
                ("dst_cap", C.c_uint32), ("wrap", C.c_uint32), ("hist_len", C.c_uint32)]


class InflateResult(C.Structure):
    _fields_ = [("rc", C.c_int32), ("out_len", C.c_uint32), ("in_used", C.c_uint32),
                ("crc32", C.c_uint32), ("adler32", C.c_uint32), ("flags", C.c_uint32)]


# every symbol include/nxgpu.h declares; tests check each one is exported
EXPORTS = [
    "tb_freq", "nx_function_begin", "nx_function_end", "nx_wait_ticks", "nxu_run_job", "__crc32_vpmsum",
    "nxgpu_open", "nxgpu_close", "nxgpu_last_error", "nxgpu_dev_alloc", "nxgpu_dev_free",
    "nxgpu_memcpy_h2d", "nxgpu_memcpy_d2h", "nxgpu_host_alloc", "nxgpu_host_free", "nxgpu_sync",
    "nxgpu_timer_start", "nxgpu_timer_stop", "nxgpu_launch_count", "nxgpu_kernel_time", "nxgpu_kernel_time_reset",
    "nxgpu_checksum_batch", "nxgpu_crc32", "nxgpu_adler32", "nxgpu_crc32_combine", "nxgpu_adler32_combine",
    "nxgpu_deflate_batch", "nxgpu_deflate_bound", "nxgpu_deflate_stream", "nxgpu_deflate_stream_bound",
    "nxgpu_inflate_batch", "nxgpu_inflate_stream", "nxgpu_makedata", "nxgpu_makedata_range", "nxgpu_job_stats",
    "nxgpu_dhtgen", "nxgpu_dhtgen_batch", "nxgpu_gunzip_concat",
    "nxgpu_team_open", "nxgpu_team_deflate", "nxgpu_team_dst", "nxgpu_team_close",
    "nxgpu_gzopen", "nxgpu_gzdopen", "nxgpu_gzread", "nxgpu_gzeof", "nxgpu_gzmembers", "nxgpu_gzclose",
]

_lib = None


def load_library() -> C.CDLL:
    """dlopen libnxgpu.so (no CUDA call is made until a context is opened)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {_HERE}/csrc` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u32, u64, i32, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t
    P = C.POINTER
    sig = {
        "nxgpu_open": (i32, [i32, P(vp)]),
        "nxgpu_close": (None, [vp]),
        "nxgpu_last_error": (C.c_char_p, []),
        "nxgpu_dev_alloc": (i32, [vp, sz, P(vp)]),
        "nxgpu_dev_free": (i32, [vp, vp]),
        "nxgpu_memcpy_h2d": (i32, [vp, vp, vp, sz]),
        "nxgpu_memcpy_d2h": (i32, [vp, vp, vp, sz]),
        "nxgpu_host_alloc": (i32, [sz, P(vp)]),
        "nxgpu_host_free": (i32, [vp]),
        "nxgpu_sync": (i32, [vp]),
        "nxgpu_timer_start": (i32, [vp]),
        "nxgpu_timer_stop": (i32, [vp, P(C.c_float)]),
        "nxgpu_launch_count": (u64, [vp]),
        "nxgpu_kernel_time": (i32, [vp, C.c_char_p, P(C.c_double), P(u64)]),
        "nxgpu_kernel_time_reset": (None, [vp]),
        "nxgpu_checksum_batch": (i32, [vp, P(CksumItem), sz, P(CksumResult), i32]),
        "nxgpu_crc32": (i32, [vp, u32, vp, u64, i32, P(u32)]),
        "nxgpu_adler32": (i32, [vp, u32, vp, u64, i32, P(u32)]),
        "nxgpu_crc32_combine": (u32, [u32, u32, u64]),
        "nxgpu_adler32_combine": (u32, [u32, u32, u64]),
        "nxgpu_deflate_batch": (i32, [vp, P(DeflateItem), sz, P(DeflateResult), i32, i32]),
        "nxgpu_deflate_bound": (u32, [u32]),
        "nxgpu_deflate_stream": (i32, [vp, vp, u64, vp, u64, i32, i32, u32, P(u64), P(StreamResult), i32]),
        "nxgpu_deflate_stream_bound": (u64, [u64, u32]),
        "nxgpu_inflate_batch": (i32, [vp, P(InflateItem), sz, P(InflateResult), i32]),
        "nxgpu_inflate_stream": (i32, [vp, vp, u64, vp, u64, i32, P(u64), u32, u32, P(StreamResult), i32]),
        "nxgpu_gunzip_concat": (i32, [vp, vp, u64, vp, u64, P(u64), P(u32), i32]),
        "nxgpu_makedata": (u64, [i32, i32, vp, u64, vp, u64]),
        "nxgpu_makedata_range": (u64, [i32, i32, vp, u64, u64, u64, vp]),
        "nxgpu_team_open": (i32, [vp, C.c_char_p, i32, i32, u64, i32, P(vp)]),
        "nxgpu_team_deflate": (i32, [vp, vp, u64, i32, i32, u32, i32, P(TeamResult)]),
        "nxgpu_team_dst": (vp, [vp]),
        "nxgpu_team_close": (None, [vp]),
        "nxgpu_gzopen": (vp, [vp, C.c_char_p, C.c_char_p]),
        "nxgpu_gzdopen": (vp, [vp, i32, C.c_char_p]),
        "nxgpu_gzread": (i32, [vp, vp, C.c_uint]),
        "nxgpu_gzeof": (i32, [vp]),
        "nxgpu_gzmembers": (u32, [vp]),
        "nxgpu_gzclose": (i32, [vp]),
        "nxgpu_job_stats": (None, [i32, P(u64), P(u64), P(u64)]),
        "nxgpu_dhtgen": (i32, [vp, P(u32), i32, P(u32), i32, C.c_char_p, P(i32), P(i32), i32]),
        "nxgpu_dhtgen_batch": (i32, [vp, P(u32), sz, vp, P(u32), i32]),
        "nx_function_begin": (i32, [i32, i32, vp]),
        "nx_function_end": (i32, [vp]),
        "nx_wait_ticks": (u64, [u64, u64, i32]),
        "nxu_run_job": (i32, [vp, vp]),
        "__crc32_vpmsum": (C.c_uint, [C.c_uint, vp, C.c_ulong]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def last_error() -> str:
    return load_library().nxgpu_last_error().decode(errors="replace")


def _addr(buf) -> Tuple[int, int, object]:
    """(address, nbytes, keepalive) of bytes / bytearray / memoryview / numpy array."""
    if isinstance(buf, (bytes, bytearray)):
        keep = (C.c_char * len(buf)).from_buffer_copy(buf) if isinstance(buf, bytes) else (C.c_char * len(buf)).from_buffer(buf)
        return C.addressof(keep), len(buf), keep
    mv = memoryview(buf)
    if not mv.contiguous:
        raise ValueError("buffer must be contiguous")
    if mv.readonly:
        keep = (C.c_char * mv.nbytes).from_buffer_copy(mv)
    else:
        keep = (C.c_char * mv.nbytes).from_buffer(mv)
    return C.addressof(keep), mv.nbytes, keep


class DeviceBuffer:
    """A raw device allocation owned by an Engine (nxgpu_dev_alloc)."""

    def __init__(self, eng: "Engine", nbytes: int):
        self.eng, self.nbytes = eng, nbytes
        p = C.c_void_p()
        eng._check(eng.lib.nxgpu_dev_alloc(eng.ctx, max(nbytes, 1), C.byref(p)), "nxgpu_dev_alloc")
        self.ptr = p.value

    def upload(self, data, offset: int = 0) -> None:
        a, n, keep = _addr(data)
        assert offset + n <= self.nbytes
        self.eng._check(self.eng.lib.nxgpu_memcpy_h2d(self.eng.ctx, self.ptr + offset, a, n), "h2d")

    def download(self, nbytes: Optional[int] = None, offset: int = 0) -> bytes:
        n = self.nbytes - offset if nbytes is None else nbytes
        out = (C.c_char * n)()
        if n:
            self.eng._check(self.eng.lib.nxgpu_memcpy_d2h(self.eng.ctx, C.addressof(out), self.ptr + offset, n), "d2h")
        return bytes(out)

    def free(self) -> None:
        if self.ptr:
            self.eng.lib.nxgpu_dev_free(self.eng.ctx, self.ptr)
            self.ptr = None


class Engine:
    """One GPU context: the GPU stand-in for an NX-GZIP engine handle
    (reference lib/nx_zlib.c:509-633 ``nx_open`` / ``nx_close``)."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        ctx = C.c_void_p()
        rc = self.lib.nxgpu_open(device, C.byref(ctx))
        if rc != 0:
            raise NxGpuError(rc, "nxgpu_open (a B200 / sm_100 GPU is required; no CPU fallback)")
        self.ctx = ctx

    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.lib.nxgpu_close(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise NxGpuError(rc, what)

    # ---- device memory ----
    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def sync(self) -> None:
        self._check(self.lib.nxgpu_sync(self.ctx), "sync")

    def launch_count(self) -> int:
        return int(self.lib.nxgpu_launch_count(self.ctx))

    def kernel_time(self, family: str) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_uint64()
        self._check(self.lib.nxgpu_kernel_time(self.ctx, family.encode(), C.byref(ms), C.byref(n)), "kernel_time")
        return ms.value, n.value

    def kernel_time_reset(self) -> None:
        self.lib.nxgpu_kernel_time_reset(self.ctx)

    def timer_start(self) -> None:
        self._check(self.lib.nxgpu_timer_start(self.ctx), "timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self.lib.nxgpu_timer_stop(self.ctx, C.byref(ms)), "timer_stop")
        return ms.value

    # ---- checksums (libnxz.h crc32 / adler32) ----
    def crc32(self, data, seed: int = 0) -> int:
        a, n, keep = _addr(data)
        out = C.c_uint32()
        self._check(self.lib.nxgpu_crc32(self.ctx, seed, a, n, MEM_HOST, C.byref(out)), "nxgpu_crc32")
        return out.value

    def adler32(self, data, seed: int = 1) -> int:
        a, n, keep = _addr(data)
        out = C.c_uint32()
        self._check(self.lib.nxgpu_adler32(self.ctx, seed, a, n, MEM_HOST, C.byref(out)), "nxgpu_adler32")
        return out.value

    def checksum_batch(self, items: Sequence[Tuple[int, int, int, int]], mem: int = MEM_DEVICE) -> List[Tuple[int, int]]:
        """items: (address, length, crc_seed, adler_seed) -> [(crc32, adler32)]"""
        n = len(items)
        arr = (CksumItem * n)(*[CksumItem(a, l, cs, ads) for a, l, cs, ads in items])
        res = (CksumResult * n)()
        self._check(self.lib.nxgpu_checksum_batch(self.ctx, arr, n, res, mem), "nxgpu_checksum_batch")
        return [(r.crc32, r.adler32) for r in res]

    def crc32_combine(self, c1: int, c2: int, len2: int) -> int:
        return int(self.lib.nxgpu_crc32_combine(c1, c2, len2))

    def adler32_combine(self, a1: int, a2: int, len2: int) -> int:
        return int(self.lib.nxgpu_adler32_combine(a1, a2, len2))

    # ---- deflate ----
    def deflate_bound(self, n: int, chunk: int = 0) -> int:
        return int(self.lib.nxgpu_deflate_stream_bound(n, chunk))

    def compress(self, data, level: int = 6, wrap: int = WRAP_ZLIB, chunk: int = 0,
                 with_index: bool = False):
        """libnxz.h ``compress2`` (lib/nx_compress.c:82): host bytes in, host bytes out."""
        a, n, keep = _addr(data)
        cap = self.deflate_bound(n, chunk)
        out = (C.c_char * cap)()
        res = StreamResult()
        nchunks = max(1, -(-n // (chunk or 262144)))
        idx = (C.c_uint64 * (nchunks + 1))() if with_index else None
        self._check(self.lib.nxgpu_deflate_stream(self.ctx, a, n, C.addressof(out), cap, level, wrap, chunk,
                                                  idx, C.byref(res), MEM_HOST), "nxgpu_deflate_stream")
        blob = bytes(memoryview(out)[: res.out_len])
        if with_index:
            return blob, list(idx), res
        return blob

    def deflate_stream_device(self, src_ptr: int, n: int, dst_ptr: int, cap: int, level: int = 6,
                              wrap: int = WRAP_GZIP, chunk: int = 0, index=None) -> StreamResult:
        res = StreamResult()
        self._check(self.lib.nxgpu_deflate_stream(self.ctx, src_ptr, n, dst_ptr, cap, level, wrap, chunk,
                                                  index, C.byref(res), MEM_DEVICE), "nxgpu_deflate_stream")
        return res

    def deflate_batch(self, items: Sequence[DeflateItem], level: int = 6, mem: int = MEM_HOST) -> List[DeflateResult]:
        n = len(items)
        arr = (DeflateItem * n)(*items)
        res = (DeflateResult * n)()
        self._check(self.lib.nxgpu_deflate_batch(self.ctx, arr, n, res, level, mem), "nxgpu_deflate_batch")
        return list(res)

    # ---- dynamic Huffman table generation (lib/nx_dhtgen.c:945) ----
    def dhtgen(self, lhist: Sequence[int], dhist: Sequence[int]) -> Tuple[bytes, int]:
        """286 lit/len + 30 distance counts -> (cpb.in_dht bytes, length in bits)."""
        la = (C.c_uint32 * len(lhist))(*lhist)
        da = (C.c_uint32 * len(dhist))(*dhist)
        out = C.create_string_buffer(320)
        nb, vb = C.c_int(), C.c_int()
        self._check(self.lib.nxgpu_dhtgen(self.ctx, la, len(lhist), da, len(dhist), out, C.byref(nb), C.byref(vb), 0), "nxgpu_dhtgen")
        return out.raw[: nb.value], 8 * nb.value - ((8 - vb.value) if vb.value else 0)

    def dhtgen_batch(self, counts: Sequence[Sequence[int]]) -> List[Tuple[bytes, int]]:
        n = len(counts)
        flat = (C.c_uint32 * (316 * n))(*[x for c in counts for x in c])
        out = C.create_string_buffer(288 * n)
        bits = (C.c_uint32 * n)()
        self._check(self.lib.nxgpu_dhtgen_batch(self.ctx, flat, n, out, bits, MEM_HOST), "nxgpu_dhtgen_batch")
        return [(out.raw[288 * i: 288 * i + (bits[i] + 7) // 8], bits[i]) for i in range(n)]

    # ---- inflate ----
    def inflate_batch(self, items: Sequence[InflateItem], mem: int = MEM_HOST) -> List[InflateResult]:
        n = len(items)
        arr = (InflateItem * n)(*items)
        res = (InflateResult * n)()
        self._check(self.lib.nxgpu_inflate_batch(self.ctx, arr, n, res, mem), "nxgpu_inflate_batch")
        return list(res)

    def inflate_stream(self, blob, out_len: int, index: Sequence[int], chunk: int = 0, wrap: int = WRAP_GZIP) -> bytes:
        """One member inflated as the segments of its sync-point index, all in one batch (host bytes)."""
        a, n, keep = _addr(blob)
        out = (C.c_char * max(out_len, 1))()
        idx = (C.c_uint64 * len(index))(*index)
        res = StreamResult()
        self._check(self.lib.nxgpu_inflate_stream(self.ctx, a, n, C.addressof(out), out_len, wrap, idx, len(index) - 1, chunk,
                                                  C.byref(res), MEM_HOST), "nxgpu_inflate_stream")
        return bytes(memoryview(out)[: res.out_len])

    def inflate_stream_device(self, src_ptr: int, n: int, dst_ptr: int, cap: int, index, n_chunks: int, chunk: int = 0,
                              wrap: int = WRAP_GZIP) -> StreamResult:
        res = StreamResult()
        self._check(self.lib.nxgpu_inflate_stream(self.ctx, src_ptr, n, dst_ptr, cap, wrap, index, n_chunks, chunk,
                                                  C.byref(res), MEM_DEVICE), "nxgpu_inflate_stream")
        return res

    def gunzip(self, blob, max_out: int) -> Tuple[bytes, int]:
        """All members of a concatenated gzip buffer, inflated as one batch -> (bytes, number of members)."""
        a, n, keep = _addr(blob)
        out = (C.c_char * max(max_out, 1))()
        total, members = C.c_uint64(), C.c_uint32()
        self._check(self.lib.nxgpu_gunzip_concat(self.ctx, a, n, C.addressof(out), max_out, C.byref(total), C.byref(members),
                                                 MEM_HOST), "nxgpu_gunzip_concat")
        return bytes(memoryview(out)[: total.value]), members.value

    def uncompress(self, blob, out_len: int, wrap: int = WRAP_AUTO) -> bytes:
        """libnxz.h ``uncompress`` (lib/nx_uncompr.c:91) for one member."""
        return self.uncompress_many([blob], [out_len], wrap)[0]

    def uncompress_many(self, blobs: Sequence[bytes], out_caps: Sequence[int], wrap: int = WRAP_AUTO) -> List[bytes]:
        keeps, items, outs = [], [], []
        for b, cap in zip(blobs, out_caps):
            a, n, k = _addr(b)
            o = (C.c_char * max(cap, 1))()
            keeps.append(k)
            outs.append(o)
            items.append(InflateItem(a, n, C.addressof(o), cap, wrap, 0))
        res = self.inflate_batch(items, MEM_HOST)
        result = []
        for r, o in zip(res, outs):
            if r.rc != 0:
                raise NxGpuError(r.rc, "inflate member")
            result.append(bytes(memoryview(o)[: r.out_len]))
        return result


class Team:
    """One member deflated by several GPUs of a box (include/nxgpu.h: nxgpu_team_*): rank `rank` of `nranks`,
    rendezvous through the shared segment `name`.  `deflate` is collective."""

    def __init__(self, eng: "Engine", name: str, rank: int, nranks: int, dst_cap: int, dst_mem: int = MEM_HOST):
        self.eng, self.rank, self.nranks, self.dst_mem = eng, rank, nranks, dst_mem
        h = C.c_void_p()
        eng._check(eng.lib.nxgpu_team_open(eng.ctx, name.encode(), rank, nranks, dst_cap, dst_mem, C.byref(h)), "nxgpu_team_open")
        self.h = h

    def deflate(self, src_ptr: int, n: int, level: int = 6, wrap: int = WRAP_GZIP, chunk: int = 0, src_mem: int = MEM_DEVICE) -> TeamResult:
        res = TeamResult()
        self.eng._check(self.eng.lib.nxgpu_team_deflate(self.h, src_ptr, n, level, wrap, chunk, src_mem, C.byref(res)), "nxgpu_team_deflate")
        return res

    def dst(self) -> int:
        return int(self.eng.lib.nxgpu_team_dst(self.h) or 0)

    def close(self) -> None:
        if self.h:
            self.eng.lib.nxgpu_team_close(self.h)
            self.h = None


def makedata(seed: int, log2size: int, seedfile: bytes) -> bytes:
    """Byte-for-byte ``makedata -s seed -b log2size < seedfile`` (reference samples/makedata.c:35-70)."""
    lib = load_library()
    cap = (1 << log2size) + (1 << log2size) // 10 + 16
    out = (C.c_char * cap)()
    a, n, keep = _addr(seedfile)
    got = lib.nxgpu_makedata(seed, log2size, a, n, C.addressof(out), cap)
    if got == 0:
        raise RuntimeError("nxgpu_makedata failed")
    return bytes(memoryview(out)[:got])
