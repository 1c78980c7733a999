#!/usr/bin/env python3
"""Developer probe: compressed sizes per level for the fixture inputs."""
import gzip, importlib.util, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
eng = pg.Engine(0)
for name, data in [("alice", alice)] + [(f"s{s}_b20", pg.makedata(s, 20, alice)) for s in (1, 4, 5)]:
    row = [name]
    for lv in (1, 3, 4, 6, 9):
        b = eng.compress(data, level=lv, wrap=pg.WRAP_ZLIB)
        assert zlib.decompress(b) == data
        row.append(f"L{lv}={len(b)}")
    row.append(f"zlib1={len(zlib.compress(data,1))} zlib6={len(zlib.compress(data,6))}")
    print(" ".join(row), flush=True)
