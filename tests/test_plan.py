"""CPU tier: the host plan compiler of libdvbt2ll_cuda.so (dvbt2ll_plan_get tables) interpreted with numpy
(tests/plan_emu.py mirrors the kernels' algorithms) against the oracle.  No GPU compute is involved;
what runs on the GPU is covered by tests/test_gpu_parity.py."""
import numpy as np
import pytest

import dvbt2ll_b200 as T
from dvbt2ll_b200 import configs as K
from oracle import t2oracle as O
import plan_emu as E
from common import bits_equal, cells_equal, max_err_over_rms, fm_args, pg_args


@pytest.mark.parametrize("name,nframes", [("c1", 8), ("c3", 2), ("c2", 1)])
def test_bb_plan(name, nframes):
    """BB header, CRC-8 sync substitution, scrambler, BCH byte table, chunking and the Horner column matrix."""
    cfg = K.resolve(name)
    bb = T.bbheaderbch_bb(cfg["framesize"], cfg["rate"], 0, 0, cfg["fecblocks"], cfg["tsrate"])
    ob = O.BbHeaderBch(cfg["framesize"], cfg["rate"], 0, 0, cfg["fecblocks"], cfg["tsrate"])
    emu = E.BbEmu(bb)
    ts = K.make_ts(3 * nframes * bb.forecast(bb.output_multiple) + 500)
    pa = pb = 0
    for _ in range(2):                      # two calls: streaming state (count, CRC history) carried
        a, ua = emu.work(ts[pa:], nframes)
        b, ub = ob.work(ts[pb:], nframes)
        assert ua == ub and bits_equal(a, b)
        pa += ua
        pb += ub


def test_bb_plan_inband():
    fs, rate = 0, K.C2_3
    bb = T.bbheaderbch_bb(fs, rate, 0, 1, 3, 4000000)
    ob = O.BbHeaderBch(fs, rate, 0, 1, 3, 4000000)
    emu = E.BbEmu(bb, inband=1, fecblocks=3)
    ts = K.make_ts(30000)
    a, ua = emu.work(ts, 7)
    b, ub = ob.work(ts, 7)
    assert ua == ub and bits_equal(a, b)


@pytest.mark.parametrize("fs,rate", [(1, K.C1_2), (1, K.C3_5), (1, K.C2_3), (1, K.C3_4), (1, K.C4_5), (1, K.C5_6),
                                     (0, K.C1_3), (0, K.C2_5), (0, K.C1_2), (0, K.C3_5), (0, K.C2_3), (0, K.C3_4),
                                     (0, K.C4_5), (0, K.C5_6)])
def test_ldpc_plan_all_codes(fs, rate):
    """Rotation form + closed-form accumulator == scatter form of the address table, all 14 T2 codes."""
    p = O.fec_params(fs, rate)
    ld = T.ldpc_bb(fs, rate)
    rng = np.random.default_rng(fs * 10 + rate)
    info = rng.integers(0, 2, p["nbch"], dtype=np.uint8)
    cw = E.ldpc_emu(ld, info, p["nbch"], p["nldpc"], p["q"])
    assert bits_equal(cw, O.ldpc_encode(info[None, :], fs, rate)[0])


def test_map_plan_all_modes():
    """bit_src + LUT for every (frame size, rate, constellation, rotation): 112 combinations, bit-exact cells."""
    rng = np.random.default_rng(1)
    for fs in (0, 1):
        for rate in ([0, 1, 2, 3, 4, 5] if fs else [6, 7, 0, 1, 2, 3, 4, 5]):
            p = O.fec_params(fs, rate)
            bits = rng.integers(0, 2, p["nldpc"], dtype=np.uint8)
            for con in range(4):
                for rot in (0, 1):
                    im = T.interleavermod_bc(fs, rate, con, rot)
                    got = E.map_emu(im, bits, p["nbch"], p["q"], 2 * (con + 1), rot)
                    want = O.interleavermod(bits[None, :], fs, rate, con, rot)[0]
                    assert cells_equal(got, want), (fs, rate, con, rot)


def _variants():
    base = K.resolve("c1")
    return [("c1", base), ("c2", K.resolve("c2")), ("c3", K.resolve("c3")), ("c4", K.resolve("c4")),
            ("l1-bpsk", dict(base, l1constellation=0, fecblocks=6)), ("l1-qpsk", dict(base, l1constellation=1, fecblocks=7)),
            ("l1-16qam-ti0", dict(base, l1constellation=2, fecblocks=7, tiblocks=0)),
            ("v131", dict(base, version=2, l1scrambled=1, reservedbiasbits=1, inband=1, inputmode=1)),
            ("1k", dict(base, fftsize=K.FFTSIZE_1K, pilotpattern=K.PILOT_PP1, guardinterval=K.GI_1_8, numdatasyms=40,
                        fecblocks=12, l1constellation=0, vlength=1024)),
            ("2k-tr", dict(base, fftsize=K.FFTSIZE_2K, pilotpattern=K.PILOT_PP2, guardinterval=K.GI_1_8, numdatasyms=30,
                           fecblocks=18, paprmode=2, vlength=2048)),
            ("8k-miso-tx2", dict(K.resolve("c2"), preamble=K.PREAMBLE_T2_MISO, misogroup=1, fecblocks=18, tiblocks=5, equalization=1)),
            ("16k-t2gi-ext", dict(K.resolve("c4"), fftsize=K.FFTSIZE_16K_T2GI, guardinterval=K.GI_19_128, carriermode=1, fecblocks=118))]


@pytest.mark.parametrize("name,cfg", _variants(), ids=[v[0] for v in _variants()])
def test_frame_and_ofdm_plans(name, cfg):
    """Frame-mapper gather table (cell int + TI + L1 + frame + zig-zag + freq int) bit-exact over t2frames+1
    frames; per-symbol carrier table + pilots + P1 + inverse sinc through a numpy IFFT within 1e-6 of RMS."""
    rng = np.random.default_rng(2)
    fm = T.framemapperfint_cc(*fm_args(cfg))
    ofm = O.FrameMapper(cfg)
    assert fm.output_multiple == ofm.mapped_items and fm.forecast(ofm.mapped_items) == ofm.stream_items
    y = None
    for fr in range(cfg["t2frames"] + 1):
        x = (rng.standard_normal(ofm.stream_items) + 1j * rng.standard_normal(ofm.stream_items)).astype(np.complex64)
        y = ofm.work(x)
        assert cells_equal(E.frame_emu(fm, x, fr), y), (name, fr)
    pg = T.pilotgenp1insert_cc(*pg_args(cfg))
    opg = O.PilotGen(cfg)
    want = opg.work(y)
    assert pg.output_multiple == want.size
    L, cps = opg.d["L"], opg.d["c_ps"]
    ct = pg.plan("ofdm.carrier_type", np.uint8).reshape(L, cps)
    for l in range(L):
        assert np.array_equal(ct[l], opg.carrier_map(l)), (name, l)
    assert max_err_over_rms(E.ofdm_emu(pg, y), want) < 1e-6


@pytest.mark.parametrize("name", ["c1", "c2", "c3"])
def test_chain16_tables(name):
    """Chain mode: the bulk-copy list (16-byte aligned spans of the frame's 16-bit cell memory, landed at 16-byte
    aligned staging offsets) + per-carrier slots reproduce exactly the composition frequency interleaver o frame o time
    interleaver o cell interleaver of the drop-in tables; copies never overlap in the staging area."""
    cfg = K.resolve(name)
    ch = T.Chain(cfg, max_frames=1)
    oc, fc = ch.plan("ofdm.code", np.int32), ch.plan("frame.code", np.int32)
    cc = ch.plan("chain.code", np.int32)
    ci_dst = ch.plan("frame.ci_dst", np.int32)
    run_desc = ch.plan("chain.run_desc", np.int32).reshape(-1, 2)
    run_ptr = ch.plan("chain.run_ptr", np.int32)
    run_cnt = ch.plan("chain.run_cnt", np.int32)
    assert np.all(run_ptr[:-1] % 2 == 0)                    # every list starts 16-byte aligned (fetched by a bulk copy)
    stage_bytes = ch.plan("chain.stage_bytes", np.int32)
    starts = ch.plan("ofdm.sym_data_start", np.int32)
    dims = ch.plan("ofdm.dims", np.int32)
    cps, L = int(dims[8]), int(dims[14])
    assert np.array_equal(np.sort(ci_dst), np.arange(ci_dst.size))
    rng = np.random.default_rng(3)
    cells16 = rng.integers(0, 65536, ci_dst.size).astype(np.int64)       # cell-interleaved memory of one T2 frame
    oc, cc = oc.reshape(L, cps), cc.reshape(L, cps)
    pad = (-ci_dst.size) % 8
    mem = np.concatenate([cells16, np.zeros(pad + 8, dtype=np.int64)])
    for l in range(L):
        runs = run_desc[run_ptr[l]:run_ptr[l] + run_cnt[l]].astype(np.int64)
        assert run_ptr[l] + run_cnt[l] <= run_ptr[l + 1] <= run_ptr[l] + run_cnt[l] + 1
        src_u, dst_u, n_u = runs[:, 0], (runs[:, 1] >> 16) & 0xFFFF, runs[:, 1] & 0xFFFF
        assert int(16 * n_u.sum()) == int(stage_bytes[l])
        stage = np.full(int(8 * (dst_u + n_u).max()) if len(runs) else 0, -1, dtype=np.int64)
        for s_, d_, n_ in zip(src_u, dst_u, n_u):                          # 8 cells per 16-byte unit
            assert np.all(stage[8 * d_:8 * (d_ + n_)] == -1), "staging copies overlap"
            assert 8 * (s_ + n_) <= mem.size
            stage[8 * d_:8 * (d_ + n_)] = mem[8 * s_:8 * (s_ + n_)]
        data = oc[l] >= 0
        assert np.array_equal(cc[l][~data], oc[l][~data])                  # pilots / nulls untouched
        f = fc[oc[l][data]]                                                # drop-in frame mapper code of each data carrier
        got = cc[l][data]
        from_cells = f >= 0
        assert np.all(got[from_cells] >= 0) and np.all(got[~from_cells] < 0)
        assert np.array_equal(stage[got[from_cells]], cells16[ci_dst[f[from_cells]]])
    assert ch.ts_bytes_per_frame == {"c1": 12352, "c2": 76304, "c3": 1084740}[name]


def test_single_table_constellation_identity():
    """Im lut[w] == Re lut[w~] bit for bit for every constellation, rotated or not (MapPlan::im_from_re): the kernels
    rely on it to look both parts of a cell up in one table."""
    for con in range(4):
        for rot in (0, 1):
            im = T.interleavermod_bc(K.FECFRAME_SHORT, K.C1_2, con, rot)
            flag, mi, mq, fl = [int(x) & 0xFFFFFFFF for x in im.plan("map.im_from_re", np.int32)]
            assert flag == 1, (con, rot)
            lut = im.plan("map.lut", np.complex64)
            w = np.arange(lut.size)
            wt = (((w << 1) & mi) | ((w >> 1) & mq)) ^ fl
            assert np.array_equal(lut[wt].real.view(np.uint32), lut.imag.copy().view(np.uint32))
