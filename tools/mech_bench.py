#!/usr/bin/env python
"""MECH-3D-n (test/tests/mechanics/mech3d.i at n^3) on the GPU: prints bench.mechanics_bench(n) as one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    print(json.dumps(bench.mechanics_bench(int(sys.argv[1]) if len(sys.argv) > 1 else 256)), flush=True)
