#!/usr/bin/env python
"""Device-resident timing of the five drop-in blocks (dvbt2ll_work_device) -- the `per_block_device` section of bench.py
on its own, for A/B runs of library variants (DVBT2LL_LIB):  python tools/block_device_bench.py [config] [T2 frames] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-dvbt2ll_b200", "python"))


def main():
    import torch
    import bench
    import dvbt2ll_b200 as T
    from dvbt2ll_b200 import configs as K
    cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
    nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    r = bench.per_block_device_extras(torch, T, K, dev, stream, peak, src, cfg, nframes, steps)
    r["lib"] = os.environ.get("DVBT2LL_LIB", "default")
    print(json.dumps(r))
    short = {k: (round(v["ms_per_call"], 4), round(v["item_gbs"]), round(v.get("algorithmic", {}).get("frac_of_peak", 0), 3)) for k, v in r["blocks"].items()}
    print("#", r["lib"], cfg, nframes, short)


if __name__ == "__main__":
    main()
