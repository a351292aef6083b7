#!/usr/bin/env python
"""Generate tests/golden/chain_<cfg>.json from the compiled UNMODIFIED reference (oracle/_ref).

The reference ships no golden vectors (its qa_*.py / test_dvbt2ll.cc are empty templates, SURVEY.md
section 4), so the fixtures are outputs of the reference itself run in this container on the synthetic
transport stream of dvbt2ll_b200.configs.make_ts: SHA-256 of the packed BCH / LDPC codewords, of the
raw complex64 cells and frame-mapper output, and the first 64 baseband samples + RMS of each T2 frame,
two consecutive T2 frames per configuration (state carried across frames).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-dvbt2ll_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from dvbt2ll_b200 import configs as K  # noqa: E402
from common import sha  # noqa: E402


def main():
    ref.lib().ref_set_quiet(1)
    for name in ("c1", "c2", "c3", "c4"):
        cfg = K.resolve(name)
        ch = ref.Chain(cfg)
        n = ch.ts_bytes_per_t2_frame()
        ts = K.make_ts(2 * n + 1000)
        frames = []
        for fr in range(2):
            r = ch.run_frame(ts)
            s = r["samples"]
            frames.append(dict(
                ts_used=int(r["ts_used"]),
                bch_sha256=sha(np.packbits(r["bch"])), fec_sha256=sha(np.packbits(r["fec"])),
                cells_sha256=sha(r["cells"]), mapped_sha256=sha(r["mapped"]),
                samples_head=[float(v) for v in s[:64].view(np.float32)],
                samples_rms=float(np.sqrt(np.mean(np.abs(s.astype(np.complex128)) ** 2))),
                n_samples=int(s.size)))
        out = dict(config=name, params={k: int(v) for k, v in cfg.items()}, ts_seed=K.TS_SEED,
                   ts_head_sha256=sha(ts[:4096]), frames=frames,
                   generator="tools/make_golden.py (oracle/_ref = unmodified reference + oracle/shim)")
        path = os.path.join(ROOT, "tests", "golden", "chain_%s.json" % name)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        print("wrote", path)


if __name__ == "__main__":
    main()
