#!/usr/bin/env python3
"""Developer sweep: device-resident deflate throughput and ratio for LZ77 parameter / parser-mask
variants (NXGPU_LZ_PARAMS="depth,lazy,nice", NXGPU_PARSER_MASK).  Not the bench.
usage: gpu_deflate_sweep.py <log2size> <seed> "depth,lazy,nice[;mask]" ..."""
import gzip, importlib.util, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
lg, seed = int(sys.argv[1]), int(sys.argv[2])
data = alice if lg == 0 else pg.makedata(seed, lg, alice)
n = len(data)
z6 = len(zlib.compress(data[: 1 << 24], 6)) / min(n, 1 << 24)
eng = pg.Engine(0)
dsrc = eng.alloc(n); dsrc.upload(data)
cap = eng.deflate_bound(n); ddst = eng.alloc(cap)
print(f"data 2^{lg} seed {seed}: {n} B, zlib-6 ratio on first 16 MiB {1/z6:.3f}", flush=True)
for cfg in sys.argv[3:]:
    params, _, mask = cfg.partition(";")
    os.environ["NXGPU_LZ_PARAMS"] = params
    if mask:
        os.environ["NXGPU_PARSER_MASK"] = mask
    else:
        os.environ.pop("NXGPU_PARSER_MASK", None)
    best = 1e9
    for it in range(3):
        eng.kernel_time_reset()
        res = eng.deflate_stream_device(dsrc.ptr, n, ddst.ptr, cap, level=6, wrap=pg.WRAP_GZIP)
        kms, kn = eng.kernel_time("deflate")
        best = min(best, kms)
    blob = ddst.download(res.out_len)
    ok = gzip.decompress(blob) == data if n <= (1 << 27) else True
    print(f"{cfg:28s} kernel {best:8.3f} ms = {n/best/1e6:7.2f} GB/s  ratio {n/res.out_len:7.3f}  tokens {res.n_tokens}  roundtrip {'ok' if ok else 'FAIL'}", flush=True)
