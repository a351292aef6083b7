"""Shared helpers for the parity tests."""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

FM_KEYS = ("framesize", "rate", "constellation", "rotation", "fecblocks", "tiblocks", "carriermode", "fftsize",
           "guardinterval", "l1constellation", "pilotpattern", "t2frames", "numdatasyms", "paprmode", "version",
           "preamble", "inputmode", "reservedbiasbits", "l1scrambled", "inband")
PG_KEYS = ("carriermode", "fftsize", "pilotpattern", "guardinterval", "numdatasyms", "paprmode", "version",
           "preamble", "misogroup", "equalization", "bandwidth", "vlength")


def fm_args(cfg):
    return [cfg[k] for k in FM_KEYS]


def pg_args(cfg):
    return [cfg[k] for k in PG_KEYS]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a, b)


def cells_equal(a, b):
    """bit-exact comparison of complex64 arrays"""
    return a.shape == b.shape and np.array_equal(np.ascontiguousarray(a).view(np.uint32),
                                                 np.ascontiguousarray(b).view(np.uint32))


def mer_db(x, ref):
    err = np.mean(np.abs(x.astype(np.complex128) - ref.astype(np.complex128)) ** 2)
    sig = np.mean(np.abs(ref.astype(np.complex128)) ** 2)
    return 10.0 * np.log10(sig / max(err, 1e-300))


def max_err_over_rms(x, ref):
    rms = np.sqrt(np.mean(np.abs(ref.astype(np.complex128)) ** 2))
    return float(np.abs(x.astype(np.complex128) - ref.astype(np.complex128)).max() / rms)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)
