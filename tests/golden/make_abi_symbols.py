#!/usr/bin/env python3
"""Regenerates tests/golden/libnxz_abi_symbols.json from the reference's ABI description test/libnxz.abi (abidw XML;
test/test_abi compares it with abidiff, which this image lacks): every defined function symbol with its version node."""
import json
import os
import re

REF = os.environ.get("REF", "/root/reference")
out = []
for m in re.finditer(r"<elf-symbol ([^>]*)/>", open(os.path.join(REF, "test", "libnxz.abi")).read()):
    a = dict(re.findall(r"([a-z-]+)='([^']*)'", m.group(1)))
    if a.get("type") == "func-type" and a.get("is-defined") == "yes" and a.get("binding") == "global-binding":
        out.append([a["name"], a.get("version", "")])
out.sort()
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnxz_abi_symbols.json")
json.dump({"generator": "tests/golden/make_abi_symbols.py over the reference's test/libnxz.abi", "functions": out}, open(path, "w"), indent=0)
print(len(out), "functions ->", path)
