"""Enum values (ABI of the reference's include/dvbt2ll/dvbt2ll_config.h:60-202), the five
BASELINE.json configurations (parameterised as in SURVEY.md section 8(d)) and the synthetic
transport-stream generator shared by the tests, bench.py and the golden-vector generator."""
import numpy as np

# dvbt2_code_rate_t
C1_2, C3_5, C2_3, C3_4, C4_5, C5_6, C1_3, C2_5 = range(8)
# dvbt2_constellation_t
MOD_QPSK, MOD_16QAM, MOD_64QAM, MOD_256QAM = range(4)
ROTATION_OFF, ROTATION_ON = 0, 1
FECFRAME_SHORT, FECFRAME_NORMAL = 0, 1
INPUTMODE_NORMAL, INPUTMODE_HIEFF = 0, 1
CARRIERS_NORMAL, CARRIERS_EXTENDED = 0, 1
PREAMBLE_T2_SISO, PREAMBLE_T2_MISO, PREAMBLE_NON_T2, PREAMBLE_T2_LITE_SISO, PREAMBLE_T2_LITE_MISO = range(5)
FFTSIZE_2K, FFTSIZE_8K, FFTSIZE_4K, FFTSIZE_1K, FFTSIZE_16K, FFTSIZE_32K, FFTSIZE_8K_T2GI, FFTSIZE_32K_T2GI = range(8)
FFTSIZE_16K_T2GI = 11
GI_1_32, GI_1_16, GI_1_8, GI_1_4, GI_1_128, GI_19_128, GI_19_256 = range(7)
PAPR_OFF, PAPR_ACE, PAPR_TR, PAPR_BOTH = range(4)
L1_MOD_BPSK, L1_MOD_QPSK, L1_MOD_16QAM, L1_MOD_64QAM = range(4)
PILOT_PP1, PILOT_PP2, PILOT_PP3, PILOT_PP4, PILOT_PP5, PILOT_PP6, PILOT_PP7, PILOT_PP8 = range(8)
VERSION_111, VERSION_121, VERSION_131 = range(3)
RESERVED_OFF, RESERVED_ON = 0, 1
L1_SCRAMBLED_OFF, L1_SCRAMBLED_ON = 0, 1
MISO_TX1, MISO_TX2 = 0, 1
INBAND_OFF, INBAND_ON = 0, 1
EQUALIZATION_OFF, EQUALIZATION_ON = 0, 1
BANDWIDTH_1_7_MHZ, BANDWIDTH_5_0_MHZ, BANDWIDTH_6_0_MHZ, BANDWIDTH_7_0_MHZ, BANDWIDTH_8_0_MHZ, BANDWIDTH_10_0_MHZ = range(6)

# GRC derives vlength from the FFT size option (grc/dvbt2ll_pilotgenp1insert_cc.xml:43-91)
VLENGTH = {FFTSIZE_1K: 1024, FFTSIZE_2K: 2048, FFTSIZE_4K: 4096, FFTSIZE_8K: 8192,
           FFTSIZE_8K_T2GI: 8192, FFTSIZE_16K: 16384, FFTSIZE_16K_T2GI: 16384,
           FFTSIZE_32K: 32768, FFTSIZE_32K_T2GI: 32768}

REALTIME_MSPS = 64.0 / 7.0   # apps/vv009-4kshort.grc: samp_rate = 8e6 * 8 / 7

_COMMON = dict(inputmode=INPUTMODE_NORMAL, inband=INBAND_OFF, paprmode=PAPR_OFF, version=VERSION_111,
               preamble=PREAMBLE_T2_SISO, misogroup=MISO_TX1, equalization=EQUALIZATION_OFF,
               l1constellation=L1_MOD_64QAM, t2frames=2, reservedbiasbits=RESERVED_OFF,
               l1scrambled=L1_SCRAMBLED_OFF, bandwidth=BANDWIDTH_8_0_MHZ, tsrate=4000000, tiblocks=3)

CONFIGS = {
    # apps/vv009-4kshort.grc as shipped
    "c1": dict(_COMMON, framesize=FECFRAME_SHORT, rate=C4_5, constellation=MOD_256QAM, rotation=ROTATION_ON,
               fecblocks=8, carriermode=CARRIERS_NORMAL, fftsize=FFTSIZE_4K, guardinterval=GI_1_32,
               pilotpattern=PILOT_PP7, numdatasyms=3),
    "c2": dict(_COMMON, framesize=FECFRAME_NORMAL, rate=C1_2, constellation=MOD_QPSK, rotation=ROTATION_OFF,
               fecblocks=19, carriermode=CARRIERS_NORMAL, fftsize=FFTSIZE_8K, guardinterval=GI_1_4,
               pilotpattern=PILOT_PP1, numdatasyms=100),
    "c3": dict(_COMMON, framesize=FECFRAME_NORMAL, rate=C2_3, constellation=MOD_256QAM, rotation=ROTATION_ON,
               fecblocks=202, carriermode=CARRIERS_EXTENDED, fftsize=FFTSIZE_32K, guardinterval=GI_1_128,
               pilotpattern=PILOT_PP7, numdatasyms=59),
    "c4": dict(_COMMON, framesize=FECFRAME_NORMAL, rate=C3_5, constellation=MOD_64QAM, rotation=ROTATION_ON,
               fecblocks=120, carriermode=CARRIERS_NORMAL, fftsize=FFTSIZE_16K, guardinterval=GI_1_16,
               pilotpattern=PILOT_PP4, numdatasyms=100),
}
CONFIGS["c5"] = dict(CONFIGS["c3"], channels=64)   # 64 independent c3 channels, seeds 0x12345678+ch

TS_SEED = 0x12345678


def resolve(cfg):
    """Accept a config name or dict; fill derived fields (vlength, channels)."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    cfg = dict(cfg)
    cfg.setdefault("vlength", VLENGTH[cfg["fftsize"]])
    cfg.setdefault("channels", 1)
    return cfg


def make_ts(nbytes, seed=TS_SEED):
    """Synthetic TS: xorshift32 advanced on EVERY byte index; byte i is 0x47 when
    i % 188 == 0, else the low 8 bits of the state."""
    # xorshift32 is linear over GF(2): vectorise by running 188-byte-aligned lanes is not possible
    # (sequential state), so generate in blocks with a tight Python-free loop via numpy object-free math.
    out = np.empty(nbytes, dtype=np.uint8)
    x = seed & 0xFFFFFFFF
    # chunked pure-python loop is slow for MBs; use the jump-free vectorised trick: state after each
    # step is a linear map; we simply iterate in C-speed using numpy on 2^k lanes via matrix powers.
    M = _xorshift_matrix()
    lanes = 4096
    if nbytes <= 65536:
        for i in range(nbytes):
            x ^= (x << 13) & 0xFFFFFFFF
            x ^= x >> 17
            x ^= (x << 5) & 0xFFFFFFFF
            out[i] = x & 0xFF
    else:
        # lane l handles bytes l*chunk .. (l+1)*chunk-1 ; its start state = M^(l*chunk) * seed
        chunk = (nbytes + lanes - 1) // lanes
        Mc = _mat_pow(M, chunk)
        states = np.empty(lanes, dtype=np.uint32)
        s = x
        for l in range(lanes):
            states[l] = s
            s = _mat_apply(Mc, s)
        buf = np.empty((lanes, chunk), dtype=np.uint8)
        st = states.copy()
        for i in range(chunk):
            st ^= st << np.uint32(13)
            st ^= st >> np.uint32(17)
            st ^= st << np.uint32(5)
            buf[:, i] = st.astype(np.uint8)
        out[:] = buf.reshape(-1)[:nbytes]
    out[::188] = 0x47
    return out


def _xorshift_matrix():
    """32x32 GF(2) matrix (as 32 column images) of one xorshift32 step."""
    cols = []
    for b in range(32):
        x = 1 << b
        x ^= (x << 13) & 0xFFFFFFFF
        x ^= x >> 17
        x ^= (x << 5) & 0xFFFFFFFF
        cols.append(x)
    return cols


def _mat_apply(M, v):
    r = 0
    b = 0
    while v:
        if v & 1:
            r ^= M[b]
        v >>= 1
        b += 1
    return r


def _mat_mul(A, B):
    return [_mat_apply(A, B[b]) for b in range(32)]


def _mat_pow(M, e):
    R = [1 << b for b in range(32)]
    P = M
    while e:
        if e & 1:
            R = _mat_mul(P, R)
        P = _mat_mul(P, P)
        e >>= 1
    return R


def fnv1a64(buf):
    """FNV-1a 64-bit over the raw bytes of an ndarray (the survey's stage checksums)."""
    b = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    # vectorised FNV is not associative; do it in chunks with python ints (fast enough for fixtures)
    h = 0xcbf29ce484222325
    prime = 0x100000001b3
    mask = 0xFFFFFFFFFFFFFFFF
    for v in b.tobytes():
        h = ((h ^ v) * prime) & mask
    return "%016x" % h
