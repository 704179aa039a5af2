#!/usr/bin/env python
"""Runs the UNMODIFIED reference program (oracle/_ref/fluidsim_cpu = /root/reference/src/main.cpp + src/fluid.cpp, built by
oracle/Makefile) on the driver cases of tests/driver_cases.py and records the CRC-32 of every frame it writes in
tests/golden/driver_frames.json.  Needs /root/reference (to build the program); the JSON travels."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import driver_cases as dc  # noqa: E402

if not os.path.exists(dc.REF_CPU):
    sys.exit("oracle/_ref/fluidsim_cpu missing: run `make -C oracle` where /root/reference exists")
gold = {"generator": "scripts/make_driver_golden.py", "program": "reference src/main.cpp + src/fluid.cpp (NUM_JACOBI_ITERS 30), "
        "g++ -O2 -ffp-contract=off, libpng 1.6.56", "cases": {}}
for name in dc.CASES:
    with tempfile.TemporaryDirectory() as d:
        lines, crcs, out = dc.run_driver(dc.REF_CPU, name, d)
        gold["cases"][name] = {"frames_crc32": crcs, "stdout_head": lines[0]}
        print(name, crcs)
with open(dc.GOLDEN_JSON, "w") as f:
    json.dump(gold, f, indent=1)
