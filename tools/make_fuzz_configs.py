#!/usr/bin/env python
"""Draws random valid parameter sets for the chain parity sweep (tests/golden/fuzz_configs.json).

FFT size / guard interval / pilot pattern follow the allowed combinations of EN 302 755 (SISO table); the number of FEC
blocks is the largest the reference's own capacity check accepts without a warning (oracle/_ref, CPU).  Run here (needs
/root/reference through oracle/_ref); the GPU test only reads the JSON."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-dvbt2ll_b200", "python"))
from oracle import ref as R                      # noqa: E402
from dvbt2ll_b200 import configs as K            # noqa: E402

ALLOWED = {   # fft -> gi -> pilot patterns (0-based PP1..PP8)
    K.FFTSIZE_1K: {K.GI_1_16: (3, 4), K.GI_1_8: (1, 2), K.GI_1_4: (0,)},
    K.FFTSIZE_2K: {K.GI_1_32: (6, 3), K.GI_1_16: (3, 4), K.GI_1_8: (1, 2), K.GI_1_4: (0,)},
    K.FFTSIZE_4K: {K.GI_1_32: (6, 3), K.GI_1_16: (3, 4), K.GI_1_8: (1, 2), K.GI_1_4: (0,)},
    K.FFTSIZE_8K: {K.GI_1_128: (6,), K.GI_1_32: (6, 3), K.GI_1_16: (7, 3, 4), K.GI_19_256: (7, 3, 4),
                   K.GI_1_8: (1, 2, 7), K.GI_19_128: (1, 2, 7), K.GI_1_4: (0, 7)},
    K.FFTSIZE_16K: {K.GI_1_128: (6,), K.GI_1_32: (6, 3, 5), K.GI_1_16: (1, 7, 3, 4), K.GI_19_256: (1, 7, 3, 4),
                    K.GI_1_8: (1, 2, 7), K.GI_19_128: (1, 2, 7), K.GI_1_4: (0, 7)},
    K.FFTSIZE_32K: {K.GI_1_128: (6,), K.GI_1_32: (3, 5), K.GI_1_16: (1, 7, 3), K.GI_19_256: (1, 7, 3),
                    K.GI_1_8: (1, 7), K.GI_19_128: (1, 7)},
}
SHORT_RATES = (K.C1_3, K.C2_5, K.C3_5, K.C2_3, K.C4_5)          # short 1/2, 3/4, 5/6: the reference's dead-code LDPC overruns
NORMAL_RATES = (K.C1_2, K.C3_5, K.C2_3, K.C3_4, K.C4_5, K.C5_6)


def maps_consistent(cfg):
    """The reference's carrier maps hold exactly the number of data carriers its own C_P2 / C_DATA / N_FC tables promise.
    (For a few non-standard combinations they differ by one -- e.g. 16K extended, PP1, GI 1/4, reserved tones, MISO group
    2: 12535 data carriers in one symbol against C_DATA = 12534 -- and the reference then reads one cell past its
    input; the plan compiler refuses those, so they are not drawn.)"""
    from oracle import t2oracle as O
    pg = O.PilotGen(K.resolve(cfg))
    d = pg.d
    for l in range(d["L"]):
        want = d["c_p2"] if l < d["n_p2"] else d["n_fc"] if (d["n_fc"] and l == d["L"] - 1) else d["c_data"]
        if int((pg.carrier_map(l) == 1).sum()) != want:
            return False
    return True


def ok(cfg):
    try:
        return R.Chain(K.resolve(cfg)).fm.warnings == 0 and maps_consistent(cfg)
    except Exception:
        return False


WIDE = False      # also draw MISO / T2-Lite preambles, the MISO group, PAPR signalling modes, reserved-bias bits, bandwidth


def draw(rng):
    fft = rng.choice(list(ALLOWED))
    gi = rng.choice(list(ALLOWED[fft]))
    pp = rng.choice(ALLOWED[fft][gi])
    short = rng.random() < 0.7 or fft in (K.FFTSIZE_1K, K.FFTSIZE_2K)
    cfg = dict(K.CONFIGS["c1"], fftsize=fft, guardinterval=gi, pilotpattern=pp,
               framesize=K.FECFRAME_SHORT if short else K.FECFRAME_NORMAL,
               rate=rng.choice(SHORT_RATES if short else NORMAL_RATES),
               constellation=rng.choice((K.MOD_QPSK, K.MOD_16QAM, K.MOD_64QAM, K.MOD_256QAM)),
               rotation=rng.choice((0, 1)), carriermode=rng.choice((0, 1)) if fft >= K.FFTSIZE_8K or fft in (K.FFTSIZE_16K, K.FFTSIZE_32K) else 0,
               l1constellation=rng.choice((0, 1, 2, 3)), paprmode=rng.choice((K.PAPR_OFF, K.PAPR_TR)),
               version=rng.choice((K.VERSION_111, K.VERSION_131)), inband=rng.choice((0, 1)), inputmode=rng.choice((0, 1)),
               misogroup=0, preamble=K.PREAMBLE_T2_SISO, equalization=rng.choice((0, 1)), t2frames=rng.choice((2, 3, 4)),
               numdatasyms=rng.choice((3, 5, 8, 12)) if fft in (K.FFTSIZE_16K, K.FFTSIZE_32K) else rng.choice((6, 10, 20, 40)))
    if WIDE:
        cfg["preamble"] = rng.choice((K.PREAMBLE_T2_SISO, K.PREAMBLE_T2_MISO, K.PREAMBLE_T2_LITE_SISO, K.PREAMBLE_T2_LITE_MISO))
        if cfg["preamble"] in (K.PREAMBLE_T2_LITE_SISO, K.PREAMBLE_T2_LITE_MISO):
            cfg["version"] = K.VERSION_131
        cfg["misogroup"] = rng.choice((K.MISO_TX1, K.MISO_TX2))
        cfg["paprmode"] = rng.choice((K.PAPR_OFF, K.PAPR_ACE, K.PAPR_TR, K.PAPR_BOTH))
        cfg["reservedbiasbits"] = rng.choice((0, 1)) if cfg["version"] == K.VERSION_131 else 0
        cfg["bandwidth"] = rng.choice(range(6))
    if cfg["version"] == K.VERSION_111:
        cfg["l1scrambled"] = 0
    else:
        cfg["l1scrambled"] = rng.choice((0, 1))
    if fft in (K.FFTSIZE_1K, K.FFTSIZE_2K, K.FFTSIZE_4K):
        cfg["carriermode"] = 0
    best = None
    for fb in range(1, 60):
        c = dict(cfg, fecblocks=fb, tiblocks=0)
        if ok(c):
            best = fb
        elif best is not None:
            break
    if best is None:
        return None
    cfg["fecblocks"] = best if rng.random() < 0.5 else rng.randint(1, best)
    cfg["tiblocks"] = rng.choice([t for t in (0, 1, 2, 3) if t <= cfg["fecblocks"]])
    return cfg if ok(cfg) else None


def main():
    # usage: make_fuzz_configs.py [seed [count [output.json [wide]]]]
    global WIDE
    WIDE = len(sys.argv) > 4 and sys.argv[4] == "wide"
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 20261018
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "tests", "golden", "fuzz_configs.json")
    rng = random.Random(seed)
    out = []
    while len(out) < count:
        c = draw(rng)
        if c is not None:
            out.append(c)
            print(len(out), {k: c[k] for k in ("fftsize", "guardinterval", "pilotpattern", "framesize", "rate", "constellation", "fecblocks", "tiblocks")}, flush=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
