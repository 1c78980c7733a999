#!/usr/bin/env python3
"""Developer probe: output size relative to zlib at levels 1 / 6 / 9 on inputs where 3- and 4-byte matches matter
(the LZ77 stage only emits matches of 5 bytes and more): source code, binary records, short-period data, the 33-symbol
alphabet text of the reference's tests (test/test_utils.c:22-28), plus the makedata / alice29 fixtures."""
import gzip, importlib.util, os, random, struct, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
from ratio_inputs import ratio_inputs
eng = pg.Engine(0)
for name, data in ratio_inputs(pg):
    row = []
    for lvl in (1, 6, 9):
        got = len(eng.compress(data, level=lvl, wrap=pg.WRAP_ZLIB))
        row.append(f"L{lvl} {got / len(zlib.compress(data, lvl)):.3f}")
    print(f"{name:18s} {len(data):8d} B  " + "  ".join(row), flush=True)
