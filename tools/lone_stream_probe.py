#!/usr/bin/env python3
"""Developer probe: ONE foreign zlib member, device-resident, serial path (one warp) against the parallel path
(csrc/inflate_par.cuh).  usage: lone_stream_probe.py [log2 bytes=26] [level=6] [seed=1]"""
import gzip, importlib.util, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
data = pg.makedata(seed, lg, alice)
eng = pg.Engine(0)
blob = zlib.compress(data, level)
dc = eng.alloc(len(blob)); dc.upload(blob)
do = eng.alloc(len(data))
for mode in ("0", "262144"):
    if mode == "0" and lg > 24 and not os.environ.get("PROBE_SERIAL"):
        continue
    os.environ["NXGPU_INFLATE_PAR_MIN"] = mode
    best, bw = 1e9, 1e9
    for it in range(3):
        eng.kernel_time_reset()
        t0 = time.perf_counter()
        r = eng.inflate_batch([pg.InflateItem(dc.ptr, len(blob), do.ptr, len(data), pg.WRAP_ZLIB, 0)], mem=pg.MEM_DEVICE)[0]
        bw = min(bw, time.perf_counter() - t0)
        kms, _ = eng.kernel_time("inflate")
        best = min(best, kms)
    ok = r.rc == 0 and r.out_len == len(data) and r.crc32 == zlib.crc32(data)
    print(f"one {len(data) >> 20} MiB zlib-{level} member ({len(blob) >> 10} KiB), {'parallel' if mode != '0' else 'one warp'}: device {best:.2f} ms = {len(data)/best/1e6:.2f} GB/s, "
          f"call {bw*1e3:.2f} ms = {len(data)/bw/1e9:.2f} GB/s, ok={ok} rc={r.rc} out={r.out_len}", flush=True)
