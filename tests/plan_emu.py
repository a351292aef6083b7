"""Test-side numpy interpreters of the host plan tables (dvbt2ll_plan_get).

They apply the SAME tables the CUDA kernels consume, with the same algorithms, on the CPU, so the
plan compiler can be checked against the oracle in the CPU-only test tier (no GPU here).  They are test
infrastructure: nothing in the product imports this file and the product has no CPU path.
"""
import numpy as np


def _words_be(b):
    """uint8 array (multiple of 4) -> big-endian uint32 words."""
    return b.reshape(-1, 4).astype(np.uint32) @ np.array([1 << 24, 1 << 16, 1 << 8, 1], dtype=np.uint32)


# ---------------------------------------------------------------------------------------------------
# block 1: BB framing + scrambler + BCH  (mirrors k_bb_bch)
# ---------------------------------------------------------------------------------------------------
class BbEmu(object):
    def __init__(self, blk, mode=0, inband=0, fecblocks=1):
        d = blk.plan("bb.dims", np.int32)
        self.kbch, self.nbch, self.q, self.r, self.D, self.chunk, self.lead, self.nldpc = [int(x) for x in d]
        self.scr = blk.plan("bb.scramble", np.uint8)
        self.crc8 = blk.plan("bb.crc8", np.uint8)
        tabs = blk.plan("bb.bch_tab", np.uint32).reshape(2, 256, 6)
        self.tab, self.tab1 = tabs[0], tabs[1]          # b x^r mod g, b x^(r+8) mod g
        self.cols = blk.plan("bb.bch_cols", np.uint32).reshape(6, 32, 6)
        self.ib = blk.plan("bb.inband", np.uint8)
        self.mode, self.inband, self.fecblocks = mode, inband, fecblocks
        self.count, self.fec_block = 0, 0
        self.hist = np.zeros(187, dtype=np.uint8)

    def _crc(self, data, init=0):
        c = init
        for v in data:
            c = int(self.crc8[c ^ int(v)])
        return c

    def _bch(self, msg):
        """per-lane chunk remainders + Horner combine with the ballot column matrix (as in the kernel)."""
        regs = []
        for lane in range(32):
            s = lane * self.chunk - self.lead
            e = s + self.chunk
            s = max(s, 0)
            r = [0] * 6
            i = s
            while i + 2 <= e:                          # two bytes per step, as in the kernel (slicing by 2)
                A = self.tab1[(r[0] >> 24) ^ int(msg[i])]
                B = self.tab[((r[0] >> 16) & 0xFF) ^ int(msg[i + 1])]
                r = [(((r[w] << 16) & 0xFFFFFFFF) | (r[w + 1] >> 16 if w < 5 else 0)) ^ int(A[w]) ^ int(B[w]) for w in range(6)]
                i += 2
            while i < e:
                T = self.tab[(r[0] >> 24) ^ int(msg[i])]
                r = [(((r[w] << 8) & 0xFFFFFFFF) | (r[w + 1] >> 24 if w < 5 else 0)) ^ int(T[w]) for w in range(6)]
                i += 1
            regs.append(r)
        acc = regs[0]
        for i in range(1, 32):
            new = []
            for w in range(6):
                word = 0
                for lane in range(32):
                    col = self.cols[w, lane]
                    x = 0
                    for c in range(6):
                        x ^= acc[c] & int(col[c])
                    word |= (bin(x).count("1") & 1) << lane      # ballot: lane -> bit lane
                new.append(word)
            acc = [new[w] ^ regs[i][w] for w in range(6)]
        out = []
        for b in range(self.r // 8):
            out.append((acc[b >> 2] >> (24 - 8 * (b & 3))) & 0xFF)
        return np.array(out, dtype=np.uint8)

    def work(self, ts, nframes):
        """NORMAL input mode. Returns (bits [nframes*nbch], consumed)."""
        assert self.mode == 0
        stream = np.concatenate([self.hist, np.asarray(ts, dtype=np.uint8)])
        base = 187
        out = np.zeros((nframes, self.nbch // 8), dtype=np.uint8)
        pos = 0
        for f in range(nframes):
            ib = bool(self.inband) and self.fec_block == 0
            Dj = self.D - (13 if ib else 0)
            hdr = np.zeros(10, dtype=np.uint8)
            hdr[0] = 0xF0
            upl, dfl = 188 * 8, self.kbch - 80 - (104 if ib else 0)
            syncd = (188 - self.count) * 8 if self.count else 0
            hdr[2], hdr[3], hdr[4], hdr[5], hdr[6] = upl >> 8, upl & 255, dfl >> 8, dfl & 255, 0x47
            hdr[7], hdr[8] = syncd >> 8, syncd & 255
            hdr[9] = self._crc(hdr[:9])
            pay = stream[base + pos: base + pos + Dj].copy()
            i0 = (188 - self.count) % 188
            for si in range(i0, Dj, 188):
                t = base + pos + si
                pay[si] = self._crc(stream[t - 187:t])
            frame = np.concatenate([hdr, pay, self.ib if ib else np.zeros(0, np.uint8)])
            assert frame.size == self.kbch // 8
            frame ^= self.scr
            out[f, :self.kbch // 8] = frame
            out[f, self.kbch // 8:] = self._bch(frame)
            pos += Dj
            self.count = (self.count + Dj) % 188
            if self.inband:
                self.fec_block = (self.fec_block + 1) % self.fecblocks
        joined = stream[:base + pos]
        self.hist = joined[-187:].copy()
        return np.unpackbits(out, axis=1).reshape(-1), pos


# ---------------------------------------------------------------------------------------------------
# LDPC in rotation form (mirrors k_ldpc)
# ---------------------------------------------------------------------------------------------------
def ldpc_emu(blk, bch_bits, nbch, nldpc, q):
    """bch_bits: [nbch] 0/1. Returns the codeword in natural order [nldpc]."""
    row_ptr = blk.plan("ldpc.row_ptr", np.uint16).astype(int)
    entries = blk.plan("ldpc.entries", np.uint32)
    groups = bch_bits.reshape(-1, 360)
    rows = np.zeros((q, 360), dtype=np.uint8)
    for t in range(q):
        for e in range(row_ptr[t], row_ptr[t + 1]):
            g, s = int(entries[e] & 0xFFFF), int(entries[e] >> 16)
            rows[t] ^= np.roll(groups[g], s)      # out[(n + s) % 360] = in[n]
    T = np.bitwise_xor.accumulate(rows, axis=0)
    C = T[q - 1]
    E = np.concatenate([[0], np.bitwise_xor.accumulate(C)[:-1]]).astype(np.uint8)
    P = T ^ E[None, :]                            # P[t][s] = p[q*s + t]
    return np.concatenate([bch_bits, P.T.reshape(-1)])


# ---------------------------------------------------------------------------------------------------
# block 3: bit interleaver + demux + mapper (mirrors k_map)
# ---------------------------------------------------------------------------------------------------
def to_u_order(bits, nbch, q):
    """natural codeword -> 'u' order (info, then q rows of 360 parity bits)."""
    u = bits.copy()
    u[nbch:] = bits[nbch:].reshape(360, q).T.reshape(-1)
    return u


def map_emu(blk, fec_bits, nbch, q, mod, rotation):
    src = blk.plan("map.bit_src", np.uint16)
    lut = blk.plan("map.lut", np.complex64)
    u = to_u_order(fec_bits, nbch, q)
    cb = u[src].reshape(-1, mod)
    w = np.zeros(cb.shape[0], dtype=np.int64)
    for b in range(mod):
        w = (w << 1) | cb[:, b]
    c = lut[w]
    if rotation:
        c = (c.real + 1j * np.roll(c.imag, 1)).astype(np.complex64)
    return c


# ---------------------------------------------------------------------------------------------------
# blocks 4 and 5: gather by code table (mirrors fetch_cell / k_gather / carrier fill of k_ofdm)
# ---------------------------------------------------------------------------------------------------
def gather_emu(code, pool, l1b, l1c, l1v, cells_in, frame_idx):
    out = np.empty(code.size, np.complex64)
    pos = code >= 0
    out[pos] = cells_in[code[pos]]
    idx = -(code[~pos] + 1)
    isl1 = (idx >= l1b) & (idx < l1b + l1c)
    idx = idx + isl1 * (frame_idx % l1v) * l1c
    out[~pos] = pool[idx]
    return out


def frame_emu(blk, cells_in, frame_idx):
    info = blk.plan("frame.info", np.int32)
    return gather_emu(blk.plan("frame.code", np.int32), blk.plan("frame.pool", np.complex64),
                      int(info[7]), int(info[8]), int(info[9]), cells_in, frame_idx)


def ofdm_emu(blk, cells_in, double=True):
    """One T2 frame of block 5 from the plan tables, IFFT by numpy (double precision)."""
    d = blk.plan("ofdm.dims", np.int32)
    N, cps, gi, L = int(d[0]), int(d[8]), int(d[13]), int(d[14])
    info = blk.plan("ofdm.info", np.int32)
    left = int(info[0])
    norm = info[2:3].view(np.float32)[0]
    code = blk.plan("ofdm.code", np.int32).reshape(L, cps)
    pool = blk.plan("ofdm.pool", np.complex64)
    p1 = blk.plan("ofdm.p1", np.complex64)
    sinc = blk.plan("ofdm.inv_sinc", np.float32)
    out = [p1]
    for l in range(L):
        carriers = gather_emu(code[l], pool, 0, 0, 1, cells_in, 0)
        X = np.zeros(N, dtype=np.complex128)
        X[left:left + cps] = carriers
        if sinc.size:
            X = (X.astype(np.complex64) * sinc).astype(np.complex128)
        x = np.fft.ifft(np.fft.ifftshift(X)) * N
        x = (x * np.float64(norm)).astype(np.complex64)
        out.append(x[N - gi:])
        out.append(x)
    return np.concatenate(out)
