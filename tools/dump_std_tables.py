#!/usr/bin/env python
"""Write oracle/std_tables.json: the EN 302 755 constant tables (as transcribed in the reference, read
from oracle/_ref) for the numpy oracle restatement oracle/t2oracle.py.  Committed; regenerate only if
the reference changes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

T = ref.tables()
out = {k: v.tolist() for k, v in T.items()}
with open(os.path.join(ROOT, "oracle", "std_tables.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print("tables:", len(out))

# cells per OFDM symbol (Tables 47-49): [fft 1K..32K][normal, extended][PP1..PP8] -> [C_data, N_FC, C_FC], PAPR off
FFT_ENUM = [3, 0, 2, 1, 4, 5]
cells = []
for fi, f in enumerate(FFT_ENUM):
    per = []
    for ext in (0, 1):
        row = []
        for pp in range(8):
            b = ref.pilotgen(ext, f, pp, 2, 10, 0, 0, 0, 0, 0, 4, 1024 << fi)
            row.append([b.get_int("C_DATA"), b.get_int("N_FC"), b.get_int("C_FC")])
        per.append(row)
    cells.append(per)
with open(os.path.join(ROOT, "oracle", "cell_counts.json"), "w") as f:
    json.dump(cells, f, separators=(",", ":"))
print("cell counts written")
