"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the gr-dvbt2ll modulator hot path.

Every function restates, in the reference's own direct form (scatter LDPC, byte-wise copies turned
into index arithmetic, per-symbol carrier maps), what the cited reference code computes; citations are
relative to the gr-dvbt2ll tree.  It is the checker for the CUDA path, never the thing measured or
shipped: only tests/, tools/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

PARITY PINNING.  The reference ships no golden vectors and its tests are empty templates (SURVEY.md
section 4), so this restatement is pinned against the reference ITSELF: tests/test_oracle.py compares
every stage below with oracle/_ref (the unmodified reference sources compiled against oracle/shim) on
the five configurations and on parameter sweeps, bit-exactly for bits / cells / frame-mapper output and
to < 1e-6 of RMS for the baseband; tests/golden/*.json holds outputs of the reference generated here
(tools/make_golden.py) for boxes where oracle/_ref is not available.  The LDPC stage (GNU Radio gr-dtv
dvb_ldpc_bb, not vendored in the reference tree, version only pinned as GNU Radio >= 3.7.2 in
CMakeLists.txt:143) follows the reference's own dead-code restatement lib/bbheaderbch_bb_impl.cc:533-646
and is additionally checked through H.c = 0.  The IFFT (FFTW3f in the reference) is a double-precision
numpy transform here; the baseband criterion is tolerance based (MER >= 90 dB, max error <= 1e-5 RMS).

Constant tables of EN 302 755 come from oracle/std_tables.json (tools/dump_std_tables.py).
"""
import json
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, "std_tables.json")) as _f:
    TAB = {k: np.array(v) for k, v in json.load(_f).items()}

# enum values, include/dvbt2ll/dvbt2ll_config.h:60-202
C1_2, C3_5, C2_3, C3_4, C4_5, C5_6, C1_3, C2_5 = range(8)
QPSK, QAM16, QAM64, QAM256 = range(4)
FFT_2K, FFT_8K, FFT_4K, FFT_1K, FFT_16K, FFT_32K, FFT_8K_T2GI, FFT_32K_T2GI = range(8)
FFT_16K_T2GI = 11
PAPR_TR, PAPR_BOTH = 2, 3
VERSION_111, VERSION_131 = 0, 2


# --------------------------------------------------------------------------------------------------
# parameters: lib/bbheaderbch_bb_impl.cc:51-165
# --------------------------------------------------------------------------------------------------
_NORMAL = {C1_2: (32208, 32400, 90, 192), C3_5: (38688, 38880, 72, 192), C2_3: (43040, 43200, 60, 160),
           C3_4: (48408, 48600, 45, 192), C4_5: (51648, 51840, 36, 192), C5_6: (53840, 54000, 30, 160)}
_SHORT = {C1_3: (5232, 5400, 30, 168), C2_5: (6312, 6480, 27, 168), C1_2: (7032, 7200, 25, 168),
          C3_5: (9552, 9720, 18, 168), C2_3: (10632, 10800, 15, 168), C3_4: (11712, 11880, 12, 168),
          C4_5: (12432, 12600, 10, 168), C5_6: (13152, 13320, 8, 168)}
_LDPC_TAB = {(1, C1_2): "ldpc_tab_1_2N", (1, C3_5): "ldpc_tab_3_5N", (1, C2_3): "ldpc_tab_2_3N_DVBT2",
             (1, C3_4): "ldpc_tab_3_4N", (1, C4_5): "ldpc_tab_4_5N", (1, C5_6): "ldpc_tab_5_6N",
             (0, C1_3): "ldpc_tab_1_3S", (0, C2_5): "ldpc_tab_2_5S", (0, C1_2): "ldpc_tab_1_2S",
             (0, C3_5): "ldpc_tab_3_5S_DVBT2", (0, C2_3): "ldpc_tab_2_3S", (0, C3_4): "ldpc_tab_3_4S",
             (0, C4_5): "ldpc_tab_4_5S", (0, C5_6): "ldpc_tab_5_6S"}


def fec_params(framesize, rate):
    kbch, nbch, q, r = (_NORMAL if framesize else _SHORT)[rate]
    return dict(kbch=kbch, nbch=nbch, q=q, r=r, nldpc=64800 if framesize else 16200)


def bb_prbs(n):
    """BB scrambler sequence, lib/bbheaderbch_bb_impl.cc:357-369."""
    out = np.zeros(n, dtype=np.uint8)
    sr = 0x4A80
    for i in range(n):
        b = (sr ^ (sr >> 1)) & 1
        out[i] = b
        sr >>= 1
        if b:
            sr |= 0x4000
    return out


def crc8_table():
    """lib/bbheaderbch_bb_impl.cc:222-240 (poly 0xD5, MSB first)."""
    t = np.zeros(256, dtype=np.uint8)
    for i in range(256):
        crc = 0
        for j in range(7, -1, -1):
            if ((i >> j) & 1) ^ ((crc >> 7) & 1):
                crc = ((crc << 1) ^ 0xD5) & 0xFF
            else:
                crc = (crc << 1) & 0xFF
        t[i] = crc
    return t


def _bits(value, n):
    return [(value >> i) & 1 for i in range(n - 1, -1, -1)]


# --------------------------------------------------------------------------------------------------
# BCH: lib/bbheaderbch_bb_impl.cc:375-531
# --------------------------------------------------------------------------------------------------
_POLY_N = [[0, 2, 3, 5, 16], [0, 1, 4, 5, 6, 8, 16], [0, 2, 3, 4, 5, 7, 8, 9, 10, 11, 16], [0, 2, 4, 6, 9, 11, 12, 14, 16],
           [0, 1, 2, 3, 5, 8, 9, 10, 11, 12, 16], [0, 2, 4, 5, 7, 8, 9, 10, 12, 13, 14, 15, 16],
           [0, 2, 5, 6, 8, 9, 10, 11, 13, 15, 16], [0, 1, 2, 5, 6, 8, 9, 12, 13, 14, 16], [0, 5, 7, 9, 10, 11, 16],
           [0, 1, 2, 5, 7, 8, 10, 12, 13, 14, 16], [0, 2, 3, 5, 9, 11, 12, 13, 16], [0, 1, 5, 6, 7, 9, 11, 12, 16]]
_POLY_S = [[0, 1, 3, 5, 14], [0, 6, 8, 11, 14], [0, 1, 2, 6, 9, 10, 14], [0, 4, 7, 8, 10, 12, 14],
           [0, 2, 4, 6, 8, 9, 11, 13, 14], [0, 3, 7, 8, 9, 13, 14], [0, 2, 5, 6, 7, 10, 11, 13, 14], [0, 5, 8, 9, 10, 11, 14],
           [0, 1, 2, 3, 9, 10, 14], [0, 3, 6, 9, 11, 12, 14], [0, 4, 11, 12, 14], [0, 1, 2, 3, 5, 6, 7, 8, 10, 13, 14]]


def bch_generator(r):
    """Product of the first r/16 normal or the 12 short minimal polynomials (:424-502); g[i] = coeff of x^i."""
    polys = _POLY_S if r == 168 else _POLY_N
    nf = 12 if r == 168 else r // 16
    g = np.array([1], dtype=np.uint8)
    for k in range(nf):
        f = np.zeros(polys[k][-1] + 1, dtype=np.uint8)
        f[polys[k]] = 1
        g = (np.convolve(g.astype(np.int64), f.astype(np.int64)) & 1).astype(np.uint8)
    return g


def bch_parity(msgs, r):
    """Systematic BCH parity of a batch of messages [F, k] (bit 0 = highest order), parity highest order first.
    Bit-serial LFSR, vectorised over the F frames (the reference does it byte-wise with a table, :504-531)."""
    g = bch_generator(r)
    taps = g[:r][::-1].copy()          # position j <-> x^(r-1-j)
    F, k = msgs.shape
    reg = np.zeros((F, r), dtype=np.uint8)
    for i in range(k):
        fb = reg[:, 0] ^ msgs[:, i]
        reg[:, :-1] = reg[:, 1:]
        reg[:, -1] = 0
        reg ^= fb[:, None] & taps[None, :]
    return reg


# --------------------------------------------------------------------------------------------------
# block 1: lib/bbheaderbch_bb_impl.cc:648-742
# --------------------------------------------------------------------------------------------------
class BbHeaderBch(object):
    def __init__(self, framesize, rate, mode, inband, fecblocks, tsrate):
        self.p = fec_params(framesize, rate)
        self.mode, self.inband, self.fecblocks, self.tsrate = mode, inband, fecblocks, tsrate
        self.count, self.crc, self.fec_block = 0, 0, 0
        self.crc_tab = crc8_table()
        self.prbs = bb_prbs(self.p["kbch"])
        self.warnings = 0

    def _header(self, count, padding):
        """add_bbheader :272-325 and add_crc8_bits :247-270 (bit-serial, reflected register)."""
        b = [1, 1, 1, 1, 0, 0, 0, 0] + [0] * 8
        b += _bits(0 if self.mode else 188 * 8, 16) + _bits(self.p["kbch"] - 80 - padding, 16)
        b += _bits(0 if self.mode else 0x47, 8) + _bits((188 - count) * 8 if count else 0, 16)
        crc = 0
        for bit in b:
            x = bit ^ (crc & 1)
            crc >>= 1
            if x:
                crc ^= 0xAB
        if self.mode:
            crc ^= 0x80
        return b + [(crc >> n) & 1 for n in range(8)]

    def _inband_b(self):
        """add_inband_type_b :327-355"""
        return [0, 1] + [0] * 65 + _bits(self.tsrate, 27) + [0] * 10

    def work(self, ts, nframes):
        kbch, nbch = self.p["kbch"], self.p["nbch"]
        out = np.zeros((nframes, nbch), dtype=np.uint8)
        pos = 0
        for f in range(nframes):
            padding = 104 if (self.fec_block == 0 and self.inband) else 0
            frame = list(self._header(self.count, padding))
            nbytes = (kbch - 80 - padding) // 8
            j = 0
            while j < nbytes:
                byte = int(ts[pos])
                if self.count == 0:
                    if byte != 0x47:
                        self.warnings += 1
                    if self.mode:                       # HIEFF :674-680: sync byte dropped
                        pos += 1
                        self.count = (self.count + 1) % 188
                        continue
                    byte, self.crc = self.crc, 0        # NORMAL :702-708: replaced by CRC-8 of previous packet
                elif not self.mode:
                    self.crc = int(self.crc_tab[byte ^ self.crc])
                pos += 1
                self.count = (self.count + 1) % 188
                frame += _bits(byte, 8)
                j += 1
            if padding:
                frame += self._inband_b()
            fr = np.array(frame, dtype=np.uint8) ^ self.prbs
            out[f, :kbch] = fr
            if self.inband:
                self.fec_block = (self.fec_block + 1) % self.fecblocks
        out[:, kbch:] = bch_parity(out[:, :kbch], self.p["r"])
        return out.reshape(-1), pos


# --------------------------------------------------------------------------------------------------
# LDPC: lib/bbheaderbch_bb_impl.cc:533-646 (scatter form, then accumulator)
# --------------------------------------------------------------------------------------------------
def ldpc_edges(table, q, nbch, nldpc):
    P = nldpc - nbch
    tab = TAB[table]
    info_idx, par_idx = [], []
    n = np.arange(360)
    for g in range(tab.shape[0]):
        for col in range(1, int(tab[g, 0]) + 1):
            info_idx.append(360 * g + n)
            par_idx.append((int(tab[g, col]) + n * q) % P)
    return np.concatenate(info_idx), np.concatenate(par_idx)


_EDGE_CACHE = {}


def ldpc_encode(frames_bits, framesize, rate, table=None, q=None, nbch=None, nldpc=None):
    """frames_bits [F, nbch] -> [F, nldpc] (info then parity in natural order)."""
    if table is None:
        p = fec_params(framesize, rate)
        table, q, nbch, nldpc = _LDPC_TAB[(1 if framesize else 0, rate)], p["q"], p["nbch"], p["nldpc"]
    key = (table, q, nbch, nldpc)
    if key not in _EDGE_CACHE:
        _EDGE_CACHE[key] = ldpc_edges(table, q, nbch, nldpc)
    ii, pi = _EDGE_CACHE[key]
    F = frames_bits.shape[0]
    P = nldpc - nbch
    par = np.zeros((F, P), dtype=np.int32)
    for f in range(F):
        par[f] = np.bincount(pi, weights=frames_bits[f, ii], minlength=P).astype(np.int64) & 1
    par = np.bitwise_xor.accumulate(par.astype(np.uint8), axis=1)     # p[j] ^= p[j-1]  (:643-645)
    return np.concatenate([frames_bits, par], axis=1)


# --------------------------------------------------------------------------------------------------
# block 3: lib/interleavermod_bc_impl.cc
# --------------------------------------------------------------------------------------------------
def constellation(con, rotation):
    """LUT construction :169-253 with the reference's arithmetic (double -> float, float complex rotate)."""
    if con == QPSK:
        n = math.sqrt(2.0)
        lut = np.array([complex(1 / n, 1 / n), complex(1 / n, -1 / n), complex(-1 / n, 1 / n), complex(-1 / n, -1 / n)])
        angle = 29.0
    else:
        lev, n, angle = {QAM16: ([3, 1, -3, -1], math.sqrt(10.0), 16.8), QAM64: ([7, 5, 1, 3, -7, -5, -1, -3], math.sqrt(42.0), 8.6),
                         QAM256: ([15, 13, 9, 11, 1, 3, 7, 5, -15, -13, -9, -11, -1, -3, -7, -5], math.sqrt(170.0), 3.576334375)}[con]
        mod = 2 * (con + 1)
        lut = np.zeros(1 << mod, dtype=np.complex128)
        for i in range(1 << mod):
            ri = ii = 0
            for b in range(mod // 2):
                ri = (ri << 1) | ((i >> (mod - 1 - 2 * b)) & 1)
                ii = (ii << 1) | ((i >> (mod - 2 - 2 * b)) & 1)
            lut[i] = complex(lev[ri] / n, lev[ii] / n)
    lut = lut.astype(np.complex64)
    if rotation:
        a = (2.0 * math.pi * angle) / 360.0
        c, s = np.float32(math.cos(a)), np.float32(math.sin(a))
        re = lut.real * c - lut.imag * s
        im = lut.real * s + lut.imag * c
        lut = (re + 1j * im).astype(np.complex64)
    return lut


def interleavermod(fec, framesize, rate, con, rotation):
    """fec [F, nldpc] bits -> [F, cell_size] complex64.  Parity interleave :549-557, column twist :558-568,
    row read :569-587, demux :588-598, map + cyclic Q delay :599-613; QPSK branch :289-330."""
    p = fec_params(framesize, rate)
    N, nbch, q = p["nldpc"], p["nbch"], p["q"]
    F = fec.shape[0]
    mod = 2 * (con + 1)
    lut = constellation(con, rotation)

    def parity_interleave(x):
        u = x.copy()
        u[:, nbch:] = x[:, nbch:].reshape(F, 360, q).transpose(0, 2, 1).reshape(F, -1)   # u[nbch+360t+s] = in[nbch+q s+t]
        return u

    if con == QPSK:
        u = parity_interleave(fec) if rate in (C1_3, C2_5) else fec
        words = (u[:, 0::2].astype(np.int64) << 1) | u[:, 1::2]
    else:
        u = parity_interleave(fec)
        normal = bool(framesize)
        if con == QAM16:
            twist = TAB["twist16n" if normal else "twist16s"]
            mux = TAB["mux16_35" if (rate == C3_5 and normal) else "mux16_13" if (rate == C1_3 and not normal)
                      else "mux16_25" if (rate == C2_5 and not normal) else "mux16"]
            ncol = 8
        elif con == QAM64:
            twist = TAB["twist64n" if normal else "twist64s"]
            mux = TAB["mux64_35" if (rate == C3_5 and normal) else "mux64_13" if (rate == C1_3 and not normal)
                      else "mux64_25" if (rate == C2_5 and not normal) else "mux64"]
            ncol = 12
        elif normal:
            twist = TAB["twist256n"]
            mux = TAB["mux256_35" if rate == C3_5 else "mux256_23" if rate == C2_3 else "mux256"]
            ncol = 16
        else:
            twist = TAB["twist256s"]
            mux = TAB["mux256s_13" if rate == C1_3 else "mux256s_25" if rate == C2_5 else "mux256s"]
            ncol = 8
        rows = N // ncol
        cols = u.reshape(F, ncol, rows)
        v = np.empty_like(cols)
        for c in range(ncol):
            v[:, c, :] = np.roll(cols[:, c, :], int(twist[c]), axis=1)     # v[(twist + r) % rows] = u[r]
        w = v.transpose(0, 2, 1)                                           # [F, rows, ncol] row read-out
        pack = np.zeros((F, rows), dtype=np.int64)
        for e in range(ncol):
            pack |= w[:, :, e].astype(np.int64) << (ncol - 1 - int(mux[e]))
        if ncol == 2 * mod:
            words = np.stack([pack >> mod, pack & ((1 << mod) - 1)], axis=2).reshape(F, -1)
        else:
            words = pack
    cells = lut[words]
    if rotation:
        cells = (cells.real + 1j * np.roll(cells.imag, 1, axis=1)).astype(np.complex64)
    return cells


# --------------------------------------------------------------------------------------------------
# OFDM dimensions: lib/framemapperfint_cc_impl.cc:290-915 == lib/pilotgenp1insert_cc_impl.cc:56-666
# --------------------------------------------------------------------------------------------------
_FFT_INDEX = {FFT_1K: 0, FFT_2K: 1, FFT_4K: 2, FFT_8K: 3, FFT_8K_T2GI: 3, FFT_16K: 4, FFT_16K_T2GI: 4, FFT_32K: 5, FFT_32K_T2GI: 5}
_CELLS = json.load(open(os.path.join(_HERE, "cell_counts.json")))


def ofdm_dims(carriermode, fftsize, pp, gi, numdatasyms, paprmode, preamble):
    fi = _FFT_INDEX[fftsize]
    siso = preamble in (0, 3)
    d = dict(fi=fi, N=1024 << fi, miso=not siso, n_p2=[16, 8, 4, 2, 1, 1][fi])
    d["c_p2"] = ([558, 1118, 2236, 4472, 8944, 22432] if siso else [546, 1098, 2198, 4398, 8814, 17612])[fi]
    ext = carriermode == 1
    d["c_ps"] = ([853, 1705, 3409, 6913, 13921, 27841] if ext else [853, 1705, 3409, 6817, 13633, 27265])[fi]
    kx = [0, 0, 0, 48, 144, 288][fi]
    d["k_ext"], d["k_offset"] = (kx, 0) if ext else (0, kx)
    c_data, n_fc, c_fc = _CELLS[fi][1 if ext else 0][pp]
    if paprmode in (PAPR_TR, PAPR_BOTH):
        t = [10, 18, 36, 72, 144, 288][fi]
        c_data, n_fc, c_fc = [x - t if x else 0 for x in (c_data, n_fc, c_fc)]
    if siso and (gi, pp) in ((4, 6), (0, 3), (1, 1), (6, 1)):      # :898-915
        n_fc = c_fc = 0
    d.update(c_data=c_data, n_fc=n_fc, c_fc=c_fc, dx=[3, 6, 6, 12, 12, 24, 24, 6][pp], dy=[4, 2, 4, 2, 4, 2, 4, 16][pp])
    N = d["N"]
    d["gi"] = [N // 32, N // 16, N // 8, N // 4, N // 128, N * 19 // 128, N * 19 // 256][gi]
    d["L"] = numdatasyms + d["n_p2"]
    d["n_data_syms"] = numdatasyms if n_fc == 0 else numdatasyms - 1
    d["active"] = d["n_p2"] * d["c_p2"] + d["n_data_syms"] * c_data + n_fc
    return d


# --------------------------------------------------------------------------------------------------
# L1 signalling: lib/framemapperfint_cc_impl.cc:114-250, :1205-1224, :1366-1910
# --------------------------------------------------------------------------------------------------
def crc32_bits(bits):
    crc = 0xFFFFFFFF
    for b in bits:
        x = b ^ ((crc >> 31) & 1)
        crc = (crc << 1) & 0xFFFFFFFF
        if x:
            crc ^= 0x04C11DB7
    return _bits(crc, 32)


def _l1_fec(kbits, nbch, table, q):
    msg = np.array(kbits, dtype=np.uint8)[None, :]
    bch = np.concatenate([msg, bch_parity(msg, 168)], axis=1)
    return ldpc_encode(bch, 0, 0, table=table, q=q, nbch=nbch, nldpc=16200)[0]


def plp_blocks(c):
    """FEC blocks per T2 frame of every PLP.  The reference is single-PLP (lib/framemapperfint_cc_impl.cc:152); a config
    with "plp_fecblocks" describes several type-1 PLPs of the same parameters (EN 302 755 7.2.3: one configurable and one
    dynamic entry per PLP in L1-post, the PLPs one after the other in the frame) -- restated here from the standard's
    field list, not from the reference."""
    p = c.get("plp_fecblocks")
    return list(p) if p and len(p) > 1 else [c["fecblocks"]]


def l1post_sig_bits(c):
    """K_sig of L1-post: configurable 35 + 35 + 89 P + 32, dynamic 71 + 48 P + 8, CRC-32 (350 for one PLP, :978)."""
    return 213 + 137 * len(plp_blocks(c))


def l1pre_cells(c, l1_post_size):
    """add_l1pre :1366-1534 -> 1840 BPSK cells."""
    v131 = c["version"] == VERSION_131
    b = (_bits(0, 8) + [c["carriermode"]] + _bits(c["preamble"], 3) + _bits(c["fftsize"] & 7, 3) + [0] + [0] +
         _bits(c["guardinterval"], 3) + _bits(c["paprmode"], 4) + _bits(c["l1constellation"], 4) + _bits(0, 2) + _bits(0, 2) +
         _bits(l1_post_size, 18) + _bits(l1post_sig_bits(c) - 32, 18) + _bits(c["pilotpattern"], 4) + _bits(0, 8) + _bits(0, 16) +
         _bits(0x3085, 16) + _bits(0x8001, 16) + _bits(c["t2frames"], 8) + _bits(c["numdatasyms"], 12) + _bits(0, 3) + [0] +
         _bits(1, 3) + _bits(0, 3) + _bits(c["version"], 4) + [c["l1scrambled"] if v131 else 0] + [0] +
         _bits(0xF if (c["reservedbiasbits"] and v131) else 0, 4))
    b += crc32_bits(b)
    k = b + [0] * (3072 - len(b))
    cw = _l1_fec(k, 3240, "ldpc_tab_1_4S_L1", 36)
    keep = np.ones(16200, dtype=bool)
    pp = TAB["pre_puncture"]
    for cidx in range(32):                                   # :1508-1520
        n = 360 if cidx < 31 else 328
        keep[3240 + np.arange(n) * 36 + int(pp[cidx])] = False
    bits = np.concatenate([cw[:200], cw[3072:3240], cw[3240:][keep[3240:]]])
    return (1.0 - 2.0 * bits).astype(np.complex64)


def l1post_cells(c, frame_idx, n_post, n_punc):
    """add_l1post :1536-1910.  plp_id_dynamic is never initialised by the reference (:240 sets plp_id twice);
    it is 0 on a fresh heap and pinned to 0 in oracle/ref_driver.cc."""
    v131 = c["version"] == VERSION_131
    bias = bool(c["reservedbiasbits"]) and v131
    plp_cod = {C1_3: 6, C2_5: 7, C1_2: 0, C3_5: 1, C2_3: 2, C3_4: 3, C4_5: 4, C5_6: 5}[c["rate"]]
    blocks = plp_blocks(c)
    cell_size = (64800 if c["framesize"] else 16200) // (2 * (c["constellation"] + 1))
    r1 = 0x7FF if bias else 0
    ff = 0xFF if bias else 0
    b = _bits(1, 15) + _bits(len(blocks), 8) + _bits(0, 4) + _bits(0, 8) + _bits(0, 3) + _bits(729833333, 32)
    for pid, nb in enumerate(blocks):                 # configurable part, one entry per PLP (:1582-1672 for the single one)
        b += (_bits(pid, 8) + _bits(1, 3) + _bits(3, 5) + [0] + _bits(0, 3) + _bits(0, 8) + _bits(1, 8) + _bits(plp_cod, 3) +
              _bits(c["constellation"], 3) + [c["rotation"]] + _bits(c["framesize"], 2) + _bits(nb, 10) + _bits(1, 8) +
              _bits(c["tiblocks"], 8) + [0, 0] + [1 if (c["inband"] and v131) else 0] + _bits(r1, 11) +
              _bits(0 if c["version"] == VERSION_111 else c["inputmode"] + 1, 2) + [0, 0])
    b += _bits(0, 2) + _bits(0x3FFFFFFF if bias else 0, 30)
    b += _bits(frame_idx, 8) + _bits(0, 22) + _bits(0, 22) + _bits(0, 8) + _bits(0, 3) + _bits(ff, 8)
    start = 0
    for pid, nb in enumerate(blocks):                 # dynamic part: PLP_ID, PLP_START (cell address), PLP_NUM_BLOCKS, reserved
        b += _bits(pid, 8) + _bits(start, 22) + _bits(nb, 10) + _bits(ff, 8)
        start += nb * cell_size
    b += _bits(ff, 8)
    b += crc32_bits(b)
    sig = np.array(b, dtype=np.uint8)
    if v131 and c["l1scrambled"]:
        sig ^= bb_prbs(sig.size)
    l1mod = c["l1constellation"]
    suffix = {0: "bqpsk", 1: "bqpsk", 2: "16qam", 3: "64qam"}[l1mod]
    pad, punct = TAB["post_padding_" + suffix], TAB["post_puncture_" + suffix]
    ksig = sig.size
    is_pad = np.zeros(7032, dtype=bool)
    if ksig <= 360:
        m, last = 19, 360 - ksig
    else:
        m = (7032 - ksig) // 360
        last = 7032 - ksig - 360 * m
    for n in range(m):
        g = int(pad[n])
        is_pad[g * 360: g * 360 + (192 if g == 19 else 360)] = True
    g = int(pad[m])
    end = g * 360 + (192 if g == 19 else 360)
    is_pad[end - last:end] = True
    k = np.zeros(7032, dtype=np.uint8)
    k[~is_pad] = sig
    cw = _l1_fec(list(k), 7200, "ldpc_tab_1_2S_L1", 25)
    keep = np.ones(16200, dtype=bool)
    full = n_punc // 360
    for cidx in range(full):
        keep[7200 + np.arange(360) * 25 + int(punct[cidx])] = False
    keep[7200 + np.arange(n_punc - full * 360) * 25 + int(punct[full])] = False
    tx = np.concatenate([cw[:7032][~is_pad], cw[7032:7200], cw[7200:][keep[7200:]]])
    assert tx.size == n_post
    if l1mod == 0:
        return (1.0 - 2.0 * tx).astype(np.complex64)
    eta = [1, 2, 4, 6][l1mod]
    lut = constellation(l1mod - 1, 0)
    if l1mod == 1:
        return lut[(tx[0::2].astype(int) << 1) | tx[1::2]]
    ncol = 2 * eta
    rows = n_post // ncol
    il = tx.reshape(ncol, rows).T                       # :1832-1852 column write, row read, no twist
    mux = TAB["l1_mux16" if eta == 4 else "l1_mux64"]
    pack = np.zeros(rows, dtype=np.int64)
    for e in range(ncol):                               # :1879-1907 the map is the SOURCE index here
        pack = (pack << 1) | il[:, int(mux[e])]
    return lut[np.stack([pack >> eta, pack & ((1 << eta) - 1)], axis=1).reshape(-1)]


# --------------------------------------------------------------------------------------------------
# block 4: lib/framemapperfint_cc_impl.cc
# --------------------------------------------------------------------------------------------------
def _lfsr_sequence(nbits_reg, taps, total, top_shift):
    """Shared shape of the two PRBS address generators (:916-931 and :1087-1103): returns the raw register
    value for i = 0..total-1 (before toggle bit / wire permutation)."""
    out = np.zeros(total, dtype=np.int64)
    reg = 0
    for i in range(total):
        if i < 2:
            reg = 0
        elif i == 2:
            reg = 1
        else:
            fb = 0
            for t in taps:
                fb ^= (reg >> t) & 1
            reg = ((reg & ((1 << nbits_reg) - 1)) >> 1) | (fb << top_shift)
        out[i] = reg
    return out


def cell_permutation(framesize, con, cell_size):
    """:999-1107"""
    deg = ([15, 14, 14, 13] if framesize else [13, 12, 12, 11])[con]
    taps = {11: [0, 3], 12: [0, 2], 13: [0, 1, 4, 6], 14: [0, 1, 4, 5, 9, 11], 15: [0, 1, 2, 12]}[deg]
    perm = []
    reg = 0
    for i in range(1 << deg):
        if i < 2:
            reg = 0
        elif i == 2:
            reg = 1
        else:
            fb = 0
            for t in taps:
                fb ^= (reg >> t) & 1
            reg = ((reg & ((1 << (deg - 1)) - 1)) >> 1) | (fb << (deg - 2))
        reg |= (i & 1) << (deg - 1)
        if reg < cell_size:
            perm.append(reg)
    return np.array(perm, dtype=np.int64), deg


def freq_interleaver_H(fi, limit, odd):
    """:916-960"""
    taps = [[0, 4], [0, 3], [0, 2], [0, 1, 4, 6], [0, 1, 4, 5, 9, 11], [0, 1, 2, 12]][fi]
    name = "bitperm32k" if fi == 5 else "bitperm%s%s" % (["1k", "2k", "4k", "8k", "16k"][fi], "odd" if odd else "even")
    perm = TAB[name]
    nbits = 9 + fi
    mmax = 1024 << fi
    regs = _lfsr_sequence(nbits, taps, mmax, nbits - 1)
    v = np.zeros(mmax, dtype=np.int64)
    for n in range(nbits):
        v |= ((regs >> n) & 1) << int(perm[n])
    v += (np.arange(mmax) & 1) * (mmax // 2)
    return v[v < limit]


class FrameMapper(object):
    def __init__(self, c):
        self.c = c
        self.d = d = ofdm_dims(c["carriermode"], c["fftsize"], c["pilotpattern"], c["guardinterval"], c["numdatasyms"],
                               c["paprmode"], c["preamble"])
        self.cell_size = (64800 if c["framesize"] else 16200) // (2 * (c["constellation"] + 1))
        self.eta = [1, 2, 4, 6][c["l1constellation"]]
        ksig = l1post_sig_bits(c)
        npt = (6 * (7032 - ksig)) // 5
        nposttmp = ksig + 168 + 9000 - npt
        if d["n_p2"] == 1:                                              # :978-987 (float ceil in the reference)
            self.n_post = int(math.ceil(np.float32(nposttmp) / np.float32(2 * self.eta))) * 2 * self.eta
        else:
            self.n_post = int(math.ceil(np.float32(nposttmp) / np.float32(self.eta * d["n_p2"]))) * self.eta * d["n_p2"]
        self.n_punc = npt - (self.n_post - nposttmp)
        self.l1pre = l1pre_cells(c, self.n_post // self.eta)
        self.perm, self.deg = cell_permutation(c["framesize"], c["constellation"], self.cell_size)
        self.stream_items = self.cell_size * c["fecblocks"]
        self.mapped_items = d["active"]
        self.dummy = self.mapped_items - self.stream_items - 1840 - self.n_post // self.eta - (d["n_fc"] - d["c_fc"])
        self.dummy_cells = (1.0 - 2.0 * bb_prbs(max(self.dummy, 0))).astype(np.complex64)    # :1912-1926
        fi = d["fi"]
        H = {}
        for kind, lim in (("data", d["c_data"]), ("p2", d["c_p2"]), ("fc", d["n_fc"])):
            H[kind, 0] = freq_interleaver_H(fi, lim, False)
            H[kind, 1] = freq_interleaver_H(fi, lim, True)
            if fi == 5:                                                  # :961-977
                inv = np.zeros_like(H[kind, 1])
                inv[H[kind, 1]] = np.arange(H[kind, 1].size)
                H[kind, 0] = inv
        self.H = H
        self.t2_frame_num = 0

    def work(self, cells):
        c, d, Nc, F = self.c, self.d, self.cell_size, self.c["fecblocks"]
        T = c["tiblocks"]
        blocks = []                                                      # TI blocks, PLP after PLP
        for Fp in plp_blocks(c):
            if T == 0:
                blocks += [1] * Fp
            else:
                small, big = Fp // T, -(-Fp // T)
                nbig = Fp % T
                blocks += [small] * (T - nbig) + [big] * nbig
        ti = np.empty(F * Nc, dtype=np.complex64)
        r_glob = 0
        for k in blocks:                                                 # cell interleaver :1973-1998
            n = 0
            for r in range(k):
                shift = Nc
                while shift >= Nc:
                    t, sh = n, 0
                    for _ in range(self.deg):
                        sh |= t & 1
                        sh <<= 1
                        t >>= 1
                    shift = sh
                    n += 1
                base = r_glob * Nc
                ti[base + (self.perm + shift) % Nc] = cells[base: base + Nc]
                r_glob += 1
        if T:                                                            # time interleaver :1999-2028
            out, o = np.empty_like(ti), 0
            rows = Nc // 5
            for k in blocks:
                cols = 5 * k
                out[o:o + rows * cols] = ti[o:o + rows * cols].reshape(cols, rows).T.reshape(-1)
                o += rows * cols
            ti = out
        l1post = l1post_cells(c, self.t2_frame_num, self.n_post, self.n_punc)
        self.t2_frame_num = (self.t2_frame_num + 1) % c["t2frames"]
        linear = np.concatenate([self.l1pre, l1post, ti, self.dummy_cells,
                                 np.zeros(d["n_fc"] - d["c_fc"], dtype=np.complex64)])
        NP, CP = d["n_p2"], d["c_p2"]
        if NP == 1:
            framed = linear
        else:                                                            # zig-zag :2047-2103
            framed = np.empty_like(linear)
            pre, post = 1840 // NP, l1post.size // NP
            rest = CP - pre - post
            read = 1840 + l1post.size
            for n in range(NP):
                framed[n * CP: n * CP + pre] = linear[n: 1840: NP][:pre]
                framed[n * CP + pre: n * CP + pre + post] = linear[1840 + n: 1840 + l1post.size: NP][:post]
                framed[n * CP + pre + post: (n + 1) * CP] = linear[read: read + rest]
                read += rest
            framed[NP * CP:] = linear[read:]
        out, off, sym = np.empty_like(framed), 0, 0                      # frequency interleaver :2104-2142
        for kind, count, n in (("p2", NP, CP), ("data", d["n_data_syms"], d["c_data"]), ("fc", 1 if d["n_fc"] else 0, d["n_fc"])):
            for _ in range(count):
                out[off: off + n] = framed[off + self.H[kind, sym & 1]]
                off += n
                sym += 1
        return out


# --------------------------------------------------------------------------------------------------
# block 5: lib/pilotgenp1insert_cc_impl.cc
# --------------------------------------------------------------------------------------------------
DATA, P2PILOT, P2PAPR, TRPAPR, SCATTERED, CONTINUAL, P2PILOT_INV, SCATTERED_INV, CONTINUAL_INV = range(1, 10)
# (FFT index, pilot pattern index) pairs whose continual-pilot loops have no MISO branch in init_pilots (:1292-2705)
_CP_NO_MISO = {0: (1, 3, 4, 6), 1: (1, 6), 2: (1, 6), 3: (1, 6), 4: (1, 5, 6), 5: (0, 2, 4, 6)}


class PilotGen(object):
    def __init__(self, c):
        self.c = c
        self.d = d = ofdm_dims(c["carriermode"], c["fftsize"], c["pilotpattern"], c["guardinterval"], c["numdatasyms"],
                               c["paprmode"], c["preamble"])
        fi, cps, kext, dx = d["fi"], d["c_ps"], d["k_ext"], d["dx"]
        self.tx2 = d["miso"] and c["misogroup"] == 1
        self.tr = c["paprmode"] in (PAPR_TR, PAPR_BOTH)
        sz = ["1k", "2k", "4k", "8k", "16k", "32k"][fi]
        self.p2res, self.trres = TAB["p2_papr_map_" + sz], TAB["tr_papr_map_" + sz]
        roff = kext if fi >= 3 else 0
        sr, prbs = 0x7FF, np.zeros(27841, dtype=np.int64)                # init_prbs :1245-1266
        for i in range(27841):
            b = (sr ^ (sr >> 2)) & 1
            prbs[i] = sr & 1
            sr >>= 1
            if b:
                sr |= 0x400
        self.prbs = prbs
        self.pn = np.unpackbits(TAB["pn_sequence_table"].astype(np.uint8))
        k = np.arange(cps)
        # P2 map :667-926
        p2 = np.full(cps, DATA)
        step = 6 if (fi == 5 and not d["miso"]) else 3

        def p2t(i):
            return np.where(self.tx2 & ((i // 3) % 2 == 1) & (i % 3 == 0), P2PILOT_INV, P2PILOT)
        p2[::step] = p2t(k[::step])
        if c["carriermode"] == 1:
            e = np.arange(kext)
            p2[e] = p2t(e)
            p2[e + cps - kext] = p2t(e + cps - kext)
        if d["miso"]:
            p2[[kext + 1, kext + 2, cps - kext - 2, cps - kext - 3]] = P2PILOT
        p2[self.p2res + roff] = P2PAPR
        if d["miso"]:
            n = self.p2res.size
            for i in range(n):
                ki = int(self.p2res[i]) + kext
                if ki % 3 == 1 and (i == n - 1 or ki + 1 != int(self.p2res[i + 1]) + kext):
                    p2[ki + 1] = P2PILOT
                if ki % 3 == 2 and (i == 0 or ki - 1 != int(self.p2res[i - 1]) + kext):
                    p2[ki - 1] = P2PILOT
        self.p2map = p2
        # frame closing map :993-1070
        fc = np.full(cps, DATA)
        fc[::dx] = np.where(self.tx2 & ((k[::dx] // dx) % 2 == 1), SCATTERED_INV, SCATTERED)
        pp = c["pilotpattern"]
        if (fi == 0 and pp in (3, 4)) or (fi == 1 and pp == 6):
            fc[cps - 2] = SCATTERED
        edge = SCATTERED_INV if (self.tx2 and (c["numdatasyms"] + d["n_p2"] - 1) % 2) else SCATTERED
        fc[0] = fc[cps - 1] = edge
        if self.tr:
            fc[self.p2res + roff] = TRPAPR
        self.fcmap = fc
        # continual pilots :1292-2705
        cp = np.full(cps, DATA)
        cp_inv = self.tx2 and pp not in _CP_NO_MISO[fi]
        mod = [1632, 1632, 3264, 6528, 13056, 0][fi]
        names = ["pp%d_cp%d" % (pp + 1, g) for g in range(1, fi + 2)]
        if c["carriermode"] == 1 and fi >= 3:
            names.append("pp%d_%s" % (pp + 1, sz))
        for nm in names:
            if nm not in TAB:
                continue
            kk = np.atleast_1d(TAB[nm])
            if mod and "_cp" in nm:
                kk = kk % mod
            kk = kk[kk < cps]
            cp[kk] = np.where(cp_inv & ((kk // dx) % 2 == 1) & (kk % dx == 0), CONTINUAL_INV, CONTINUAL)
        self.cpmap = cp
        a_p2 = math.sqrt(37.0) / 5.0 if (fi == 5 and not d["miso"]) else math.sqrt(31.0) / 5.0      # :1083-1094
        a_cp = 4.0 / 3.0 if fi <= 1 else (4.0 * math.sqrt(2.0)) / 3.0 if fi == 2 else 8.0 / 3.0
        a_sp = 4.0 / 3.0 if pp <= 1 else 7.0 / 4.0 if pp <= 3 else 7.0 / 3.0
        self.amp = {P2PILOT: a_p2, P2PILOT_INV: -a_p2, SCATTERED: a_sp, SCATTERED_INV: -a_sp, CONTINUAL: a_cp, CONTINUAL_INV: -a_cp}
        self.norm = np.float32(5.0 / math.sqrt(27.0 * cps))             # :1095
        self.left = (d["N"] - cps) // 2 + 1
        self.p1 = self._p1()
        self.inv_sinc = self._inv_sinc() if c["equalization"] else None

    def carrier_map(self, l):
        d = self.d
        if l < d["n_p2"]:
            return self.p2map
        if d["n_fc"] and l == d["L"] - 1:
            return self.fcmap
        m = self.cpmap.copy()                                            # init_pilots tail :2706-2781
        k = np.arange(d["c_ps"])
        dx, dy, kext = d["dx"], d["dy"], d["k_ext"]
        sp = ((k - kext) % (dx * dy)) == dx * (l % dy)
        m[sp] = np.where(self.tx2 & ((k[sp] // dx) % 2 == 1), SCATTERED_INV, SCATTERED)
        m[0] = m[-1] = SCATTERED_INV if (self.tx2 and l % 2) else SCATTERED
        if self.tr:
            shift = dx * ((l + kext // dx) % dy) if self.c["carriermode"] == 1 else dx * (l % dy)
            m[self.trres + shift] = TRPAPR
        return m

    def _p1(self):
        """P1 :1119-1178, :2802-2810 (double-precision IFFT here)."""
        sr, rnd = 0x4E46, np.zeros(384)
        for i in range(384):
            b = (sr ^ (sr >> 1)) & 1
            rnd[i] = -1.0 if b else 1.0
            sr >>= 1
            if b:
                sr |= 0x4000
        s1 = TAB["s1_modulation_patterns"].reshape(8, 8)[self.c["preamble"] & 7].astype(np.uint8)
        s2 = TAB["s2_modulation_patterns"].reshape(16, 32)[((self.c["fftsize"] & 7) << 1) & 15].astype(np.uint8)
        bits = np.concatenate([np.unpackbits(s1), np.unpackbits(s2), np.unpackbits(s1)])
        seq = np.ones(385)
        for i in range(1, 385):
            seq[i] = -seq[i - 1] if bits[i - 1] else seq[i - 1]
        freq = np.zeros(1024)
        freq[TAB["p1_active_carriers"] + 86] = seq[1:] * rnd
        scale = np.float32(math.sqrt(384.0))
        t = (np.fft.ifft(np.fft.ifftshift(freq)) * 1024).astype(np.complex64) / scale
        ts = (np.fft.ifft(np.fft.ifftshift(np.roll(freq, 1))) * 1024).astype(np.complex64) / scale
        return np.concatenate([ts[:542], t, ts[542:]]).astype(np.complex64)

    def _inv_sinc(self):
        """:1179-1219"""
        N = self.d["N"]
        fs = {0: 131.0 * 1e6 / 71.0, 1: 5.0 * 8e6 / 7.0, 2: 6.0 * 8e6 / 7.0, 3: 7.0 * 8e6 / 7.0, 4: 8.0 * 8e6 / 7.0,
              5: 10.0 * 8e6 / 7.0}.get(self.c["bandwidth"], 1.0)
        fstep = fs / N
        inv = np.zeros(N, dtype=np.float32)
        f = rms = 0.0
        for i in range(N // 2):
            x = math.pi * f / fs
            s = 1.0 if i == 0 else math.sin(x) / x
            rms += s * s
            inv[i + N // 2] = inv[N // 2 - i - 1] = np.float32(1.0 / s)
            f += fstep
        return inv * np.float32(math.sqrt(rms / (N // 2)))

    def work(self, cells):
        """general_work :2784-2907: carrier fill, fftshift, backward unnormalised FFT, scale, CP, P1 first."""
        d = self.d
        N, cps, gi = d["N"], d["c_ps"], d["gi"]
        out = [self.p1]
        pos = 0
        for l in range(d["L"]):
            m = self.carrier_map(l)
            X = np.zeros(cps, dtype=np.complex64)
            sign = 1.0 - 2.0 * (self.prbs[np.arange(cps) + d["k_offset"]] ^ int(self.pn[l]))
            for t, a in self.amp.items():
                sel = m == t
                X[sel] = (np.float32(a) * sign[sel]).astype(np.float32)
            sel = m == DATA
            n = int(sel.sum())
            X[sel] = cells[pos: pos + n]
            pos += n
            full = np.zeros(N, dtype=np.complex64)
            full[self.left: self.left + cps] = X
            if self.inv_sinc is not None:
                full = (full * self.inv_sinc).astype(np.complex64)
            x = np.fft.ifft(np.fft.ifftshift(full.astype(np.complex128))) * N
            x = (x * np.float64(self.norm)).astype(np.complex64)
            out += [x[N - gi:], x]
        assert pos == d["active"]
        return np.concatenate(out)


# --------------------------------------------------------------------------------------------------
def chain(cfg, ts, nframes):
    """The shipped flowgraph order (apps/vv009-4kshort.grc) for nframes T2 frames of one channel.  With several PLPs
    (cfg["plp_fecblocks"]) `ts` is a list of transport streams, one per PLP, each with its own BB framing."""
    c = cfg
    p = fec_params(c["framesize"], c["rate"])
    pb = plp_blocks(c)
    tss = list(ts) if len(pb) > 1 else [ts]
    bbs = [BbHeaderBch(c["framesize"], c["rate"], c["inputmode"], c["inband"], nb, c["tsrate"]) for nb in pb]
    fm, pg = FrameMapper(c), PilotGen(c)
    res = dict(bch=[], fec=[], cells=[], mapped=[], samples=[])
    pos = [0] * len(pb)
    for _ in range(nframes):
        parts = []
        for i, nb in enumerate(pb):
            part, used = bbs[i].work(tss[i][pos[i]:], nb)
            pos[i] += used
            parts.append(part)
        bch = np.concatenate(parts)
        F = sum(pb)
        fec = ldpc_encode(bch.reshape(F, p["nbch"]), c["framesize"], c["rate"])
        cells = interleavermod(fec, c["framesize"], c["rate"], c["constellation"], c["rotation"]).reshape(-1)
        mapped = fm.work(cells)
        samples = pg.work(mapped)
        for k, v in (("bch", bch), ("fec", fec.reshape(-1)), ("cells", cells), ("mapped", mapped), ("samples", samples)):
            res[k].append(v)
    out = {k: np.concatenate(v) for k, v in res.items()}
    out["ts_used"] = pos[0] if len(pb) == 1 else pos
    return out
