#!/usr/bin/env python3
"""Build kernel experiment variants (extra -D flags) and, on the GPU box, time each.
    python tools/variants.py build  name=D1,D2 name2=D3 ...
    python tools/variants.py run    name name2 ...      (under gpurun)
"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    mode, items = sys.argv[1], sys.argv[2:]
    if mode == "build":
        from astc_encoder_b200 import build as B
        from concurrent.futures import ThreadPoolExecutor
        def one(it):
            name, _, defs = it.partition("=")
            lib = B.build_variant(name, [d for d in defs.split(",") if d])
            return name, lib
        with ThreadPoolExecutor(4) as ex:
            for name, lib in ex.map(one, items):
                print("built", name, lib)
    else:
        for name in items:
            lib = ROOT / "astc_encoder_b200" / (f"libastc_b200_{name}.so" if name != "base" else "libastc_b200.so")
            env = dict(os.environ, ASTC_B200_LIB=str(lib))
            out = subprocess.run([sys.executable, str(ROOT / "tools" / "quick_bench.py"), "short"], env=env, capture_output=True, text=True)
            print(f"== {name}\n{out.stdout.strip()}\n{out.stderr.strip()[-400:]}", flush=True)


if __name__ == "__main__":
    main()
