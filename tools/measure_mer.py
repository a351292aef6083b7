#!/usr/bin/env python
"""Print MER and max error / RMS of the GPU chain against the oracle for the four single-channel configs
(run on the GPU box): evidence for the baseband criterion MER >= 90 dB, max error <= 1e-5 RMS."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-dvbt2ll_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dvbt2ll_b200 as T  # noqa: E402
from dvbt2ll_b200 import configs as K  # noqa: E402
from common import mer_db, max_err_over_rms  # noqa: E402
from oracle import ref, t2oracle  # noqa: E402

for name in ("c1", "c2", "c3", "c4"):
    cfg = K.resolve(name)
    ch = T.Chain(cfg, max_frames=1)
    ts = K.make_ts(ch.ts_bytes_per_frame + 1000)
    got = ch.run_host(ts[:ch.ts_bytes_per_frame], 1, 1)[0]
    if ref.available():
        ref.lib().ref_set_quiet(1)
        want = ref.Chain(cfg).run_frame(ts)["samples"]
        src = "oracle/_ref (double-precision IFFT)"
    else:
        want = t2oracle.chain(cfg, ts, 1)["samples"]
        src = "oracle/t2oracle.py"
    print("%s: MER %.1f dB, max error / RMS %.3g  (vs %s)" % (name, mer_db(got, want), max_err_over_rms(got, want), src))
