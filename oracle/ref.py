"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle/_ref/libdvbt2ll_ref.so.

That library is the UNMODIFIED reference (gr-dvbt2ll lib/*_impl.cc) compiled against the
GNU Radio stand-in headers in oracle/shim/ (recipe: oracle/Makefile).  It is the strongest
checker available: the numpy restatement in oracle/t2oracle.py is pinned against it, golden
fixtures in tests/golden/ are generated from it (tools/make_golden.py), and bench.py times it
as the CPU baseline.  Only tests/, tools/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdvbt2ll_ref.so")

_lib = None

# (framesize, rate) pairs for which the reference's own (dead-code) LDPC encoder is unusable: its lookup
# table (LDPC_BF macro, lib/bbheaderbch_bb_impl.cc:533-561) reserves floor(edges*360/P)+2 slots per
# parity bit, but for the irregular short codes 1/2, 3/4 and 5/6 some check nodes have more neighbours
# than that, so ldpc_lookup_generate() overruns its rows (heap overflow already in the constructor) and
# ldpc_calculate() returns garbage.  The live path (gr-dtv dvb_ldpc_bb) has no such limit; for these
# codes the numpy oracle (scatter form straight from the address table, checked through H.c = 0) is the
# only checker, and tests do not construct the reference's bbheaderbch block for them.
REF_LDPC_BROKEN = {(0, 0), (0, 3), (0, 5)}


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_bbheaderbch_new.restype = C.c_void_p
        L.ref_bbheaderbch_new.argtypes = [C.c_int] * 6
        L.ref_interleavermod_new.restype = C.c_void_p
        L.ref_interleavermod_new.argtypes = [C.c_int] * 4
        L.ref_framemapper_new.restype = C.c_void_p
        L.ref_framemapper_new.argtypes = [C.c_int] * 20
        L.ref_pilotgen_new.restype = C.c_void_p
        L.ref_pilotgen_new.argtypes = [C.c_int] * 12
        L.ref_block_free.argtypes = [C.c_void_p]
        L.ref_block_output_multiple.argtypes = [C.c_void_p]
        L.ref_block_warnings.argtypes = [C.c_void_p]
        L.ref_block_forecast.argtypes = [C.c_void_p, C.c_int]
        L.ref_block_work.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                     C.POINTER(C.c_int)]
        L.ref_ldpc_calculate.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_get_int.restype = C.c_long
        L.ref_get_int.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_get_float.restype = C.c_float
        L.ref_get_float.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_get_int_array.restype = C.POINTER(C.c_int)
        L.ref_get_int_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.ref_get_complex_array.restype = C.POINTER(C.c_float)
        L.ref_get_complex_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.ref_framemapper_l1post.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_num_tables.restype = C.c_int
        L.ref_table_name.restype = C.c_char_p
        L.ref_table_name.argtypes = [C.c_int]
        L.ref_table.restype = C.POINTER(C.c_int)
        L.ref_table.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_byte_table.restype = C.POINTER(C.c_ubyte)
        L.ref_byte_table.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.ref_set_quiet.argtypes = [C.c_int]
        L.ref_set_fft_fast.argtypes = [C.c_int]
        L.ref_fft_seconds.restype = C.c_double
        L.ref_fft_seconds.argtypes = [C.c_int]
        _lib = L
    return _lib


class Block:
    """One reference block instance (bbheaderbch / interleavermod / framemapper / pilotgen)."""

    def __init__(self, handle, in_dtype, out_dtype):
        if not handle:
            raise RuntimeError("reference block construction failed")
        self.h = C.c_void_p(handle)
        self.in_dtype = np.dtype(in_dtype)
        self.out_dtype = np.dtype(out_dtype)

    def __del__(self):
        try:
            lib().ref_block_free(self.h)
        except Exception:
            pass

    @property
    def output_multiple(self):
        return lib().ref_block_output_multiple(self.h)

    @property
    def warnings(self):
        return lib().ref_block_warnings(self.h)

    def forecast(self, noutput):
        return lib().ref_block_forecast(self.h, noutput)

    def work(self, data, nframes):
        """Run nframes frames (one general_work call per frame). Returns (out, consumed)."""
        data = np.ascontiguousarray(data, dtype=self.in_dtype)
        nout = nframes * self.output_multiple
        out = np.zeros(nout, dtype=self.out_dtype)
        consumed = C.c_int(0)
        r = lib().ref_block_work(self.h, nout, data.ctypes.data, data.size, out.ctypes.data,
                                 C.byref(consumed))
        if r != nout:
            raise RuntimeError("reference produced %d of %d items" % (r, nout))
        return out, consumed.value

    def get_int(self, name):
        v = lib().ref_get_int(self.h, name.encode())
        if v == -999999999:
            raise KeyError(name)
        return int(v)

    def get_float(self, name):
        return float(lib().ref_get_float(self.h, name.encode()))

    def get_int_array(self, name):
        n = C.c_int(0)
        p = lib().ref_get_int_array(self.h, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def get_complex_array(self, name):
        n = C.c_int(0)
        p = lib().ref_get_complex_array(self.h, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(p, shape=(2 * n.value,)).copy().view(np.complex64)

    def ldpc(self, bch_bits, frame_size):
        """Reference's own LDPC (dead code in bbheaderbch_bb_impl.cc:625-646) on one frame."""
        buf = np.zeros(frame_size, dtype=np.uint8)
        buf[:bch_bits.size] = bch_bits
        lib().ref_ldpc_calculate(self.h, buf.ctypes.data)
        return buf

    def l1post(self, frame_idx):
        n = self.get_int("N_post") // self.get_int("eta_mod")
        out = np.zeros(n, dtype=np.complex64)
        lib().ref_framemapper_l1post(self.h, frame_idx, out.ctypes.data)
        return out


def bbheaderbch(framesize, rate, mode, inband, fecblocks, tsrate):
    if (framesize, rate) in REF_LDPC_BROKEN:
        raise NotImplementedError("reference bbheaderbch_bb overruns its LDPC lookup table for this code (see REF_LDPC_BROKEN)")
    return Block(lib().ref_bbheaderbch_new(framesize, rate, mode, inband, fecblocks, tsrate),
                 np.uint8, np.uint8)


def interleavermod(framesize, rate, constellation, rotation):
    return Block(lib().ref_interleavermod_new(framesize, rate, constellation, rotation),
                 np.uint8, np.complex64)


def framemapper(framesize, rate, constellation, rotation, fecblocks, tiblocks, carriermode,
                fftsize, guardinterval, l1constellation, pilotpattern, t2frames, numdatasyms,
                paprmode, version, preamble, inputmode, reservedbiasbits, l1scrambled, inband):
    return Block(lib().ref_framemapper_new(framesize, rate, constellation, rotation, fecblocks,
                                           tiblocks, carriermode, fftsize, guardinterval,
                                           l1constellation, pilotpattern, t2frames, numdatasyms,
                                           paprmode, version, preamble, inputmode,
                                           reservedbiasbits, l1scrambled, inband),
                 np.complex64, np.complex64)


def pilotgen(carriermode, fftsize, pilotpattern, guardinterval, numdatasyms, paprmode, version,
             preamble, misogroup, equalization, bandwidth, vlength):
    return Block(lib().ref_pilotgen_new(carriermode, fftsize, pilotpattern, guardinterval,
                                        numdatasyms, paprmode, version, preamble, misogroup,
                                        equalization, bandwidth, vlength),
                 np.complex64, np.complex64)


def tables():
    """All constant tables of the standard as transcribed in the reference: name -> ndarray."""
    L = lib()
    out = {}
    for i in range(L.ref_num_tables()):
        r, c = C.c_int(0), C.c_int(0)
        p = L.ref_table(i, C.byref(r), C.byref(c))
        a = np.ctypeslib.as_array(p, shape=(r.value, c.value)).copy()
        out[L.ref_table_name(i).decode()] = a if r.value > 1 else a[0]
    for name in ("pn_sequence_table", "s1_modulation_patterns", "s2_modulation_patterns"):
        n = C.c_int(0)
        p = L.ref_byte_table(name.encode(), C.byref(n))
        out[name] = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    return out


class Chain:
    """The reference flowgraph order: bbheaderbch -> LDPC -> interleavermod -> framemapper -> pilotgen
    (apps/vv009-4kshort.grc), driven one T2 frame at a time with state carried across frames."""

    def __init__(self, cfg):
        self.cfg = cfg = dict(cfg)   # fully resolved dict (dvbt2ll_b200.configs.resolve)
        self.bb = bbheaderbch(cfg["framesize"], cfg["rate"], cfg["inputmode"], cfg["inband"],
                              cfg["fecblocks"], cfg["tsrate"])
        self.im = interleavermod(cfg["framesize"], cfg["rate"], cfg["constellation"], cfg["rotation"])
        self.fm = framemapper(cfg["framesize"], cfg["rate"], cfg["constellation"], cfg["rotation"],
                              cfg["fecblocks"], cfg["tiblocks"], cfg["carriermode"], cfg["fftsize"],
                              cfg["guardinterval"], cfg["l1constellation"], cfg["pilotpattern"],
                              cfg["t2frames"], cfg["numdatasyms"], cfg["paprmode"], cfg["version"],
                              cfg["preamble"], cfg["inputmode"], cfg["reservedbiasbits"],
                              cfg["l1scrambled"], cfg["inband"])
        self.pg = pilotgen(cfg["carriermode"], cfg["fftsize"], cfg["pilotpattern"],
                           cfg["guardinterval"], cfg["numdatasyms"], cfg["paprmode"], cfg["version"],
                           cfg["preamble"], cfg["misogroup"], cfg["equalization"], cfg["bandwidth"],
                           cfg["vlength"])
        self.frame_size = 64800 if cfg["framesize"] == 1 else 16200
        self.nbch = self.bb.get_int("nbch")
        self.kbch = self.bb.get_int("kbch")
        self.ts_pos = 0

    def ts_bytes_per_t2_frame(self):
        # NORMAL input mode without in-band: (kbch-80)/8 TS bytes per FECFRAME
        return self.cfg["fecblocks"] * ((self.kbch - 80) // 8)

    def run_frame(self, ts, stages=None, timers=None):
        """Consume TS for ONE T2 frame starting at self.ts_pos. Returns dict of stage outputs."""
        import time
        F = self.cfg["fecblocks"]
        need = sum(self.bb.forecast(self.nbch) for _ in range(F)) + 400
        t0 = time.perf_counter()
        bch, used = self.bb.work(ts[self.ts_pos:self.ts_pos + need], F)
        self.ts_pos += used
        t1 = time.perf_counter()
        fec = np.zeros(F * self.frame_size, dtype=np.uint8)
        for f in range(F):
            fec[f * self.frame_size:(f + 1) * self.frame_size] = self.bb.ldpc(
                bch[f * self.nbch:(f + 1) * self.nbch], self.frame_size)
        t2 = time.perf_counter()
        cells, _ = self.im.work(fec, F)
        t3 = time.perf_counter()
        mapped, _ = self.fm.work(cells, 1)
        t4 = time.perf_counter()
        samples, _ = self.pg.work(mapped, 1)
        t5 = time.perf_counter()
        if timers is not None:
            for k, v in zip(("bbheaderbch", "ldpc", "interleavermod", "framemapper", "pilotgen"),
                            (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                timers[k] = timers.get(k, 0.0) + v
        return {"bch": bch, "fec": fec, "cells": cells, "mapped": mapped, "samples": samples,
                "ts_used": used}
