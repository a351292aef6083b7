"""The C-ABI library loads without a GPU, exports every symbol include/dvbt2ll_cuda.h declares, builds
plans on the host, and FAILS LOUDLY (no CPU fallback) when asked to compute without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dvbt2ll_b200 as T
from dvbt2ll_b200 import configs as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dvbt2ll_cuda.h")).read()
    return sorted(set(re.findall(r"DVBT2LL_API_EXPORT[^;]*?\b(dvbt2ll_\w+)\s*\(", text, flags=re.S)))


def test_header_symbols_exported():
    L = T.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "symbol %s declared in include/dvbt2ll_cuda.h is not exported" % n
    assert sorted(T.EXPORTED_SYMBOLS) == names


def test_factories_validate_parameters():
    with pytest.raises(ValueError):
        T.bbheaderbch_bb(K.FECFRAME_NORMAL, K.C1_3, 0, 0, 1, 0)          # 1/3 exists only for short frames
    with pytest.raises(ValueError):
        T.interleavermod_bc(K.FECFRAME_SHORT, K.C1_2, 7, 0)              # unknown constellation
    with pytest.raises(ValueError):
        T.pilotgenp1insert_cc(0, K.FFTSIZE_1K, K.PILOT_PP8, K.GI_1_8, 10, 0, 0, 0, 0, 0, 4, 1024)   # PP8 not defined for 1K
    with pytest.raises(ValueError):
        T.pilotgenp1insert_cc(0, K.FFTSIZE_8K, K.PILOT_PP1, K.GI_1_8, 10, 0, 0, 0, 0, 0, 4, 4096)   # vlength != FFT size
    cfg = K.resolve("c1")
    with pytest.raises(ValueError) as e:
        T.blocks_for(dict(cfg, fecblocks=9))                             # reference: "too many FEC blocks in T2 frame"
    assert "too many FEC blocks" in str(e.value)


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4"])
def test_scheduling_contract_matches_reference_formulas(name):
    """set_output_multiple / forecast values (reference: bbheaderbch :195,:207-216; interleavermod :254,:264-268;
    framemapper :1135,:1942-1946; pilotgen :1228,:1239-1243) and the sizes of SURVEY.md 8(d)."""
    cfg = K.resolve(name)
    B = T.blocks_for(cfg)
    want = {"c1": (12600, 2025, 18866, 31616, 16200), "c2": (32400, 32400, 628080, 1046528, 615600),
            "c3": (43200, 8100, 1639268, 1983488, 1636200), "c4": (38880, 10800, 1308638, 1760256, 1296000)}[name]
    assert B["bb"].output_multiple == want[0]
    assert B["im"].output_multiple == want[1]
    assert B["fm"].output_multiple == want[2]
    assert B["pg"].output_multiple == want[3]
    assert B["fm"].forecast(want[2]) == want[4]
    assert B["pg"].forecast(want[3]) == want[2]
    kbch = {"c1": 12432, "c2": 32208, "c3": 43040, "c4": 38688}[name]
    assert B["bb"].forecast(want[0]) == (kbch - 80) // 8
    N = 64800 if cfg["framesize"] else 16200
    assert B["ldpc"].output_multiple == N and B["ldpc"].forecast(N) == want[0]
    assert B["im"].forecast(want[1]) == N


@pytest.mark.skipif(T.device_available(), reason="a CUDA device is present")
def test_no_cpu_fallback():
    cfg = K.resolve("c1")
    bb = T.bbheaderbch_bb(cfg["framesize"], cfg["rate"], 0, 0, 8, 0)
    ts = K.make_ts(4000)
    with pytest.raises(RuntimeError) as e:
        bb.work(ts, 1)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    ch = T.Chain(cfg, max_frames=1)
    with pytest.raises(RuntimeError):
        ch.run_host(K.make_ts(ch.ts_bytes_per_frame), 1, 1)


def test_overfull_policy_is_opt_in():
    """Reference lib/framemapperfint_cc_impl.cc:1138-1141 warns and keeps going; here that is an explicit opt-in."""
    from common import fm_args
    cfg = K.resolve(dict(K.CONFIGS["c1"], fecblocks=9))
    T.set_overfull_policy(False)
    with pytest.raises(ValueError):
        T.framemapperfint_cc(*fm_args(cfg))
    T.set_overfull_policy(True)
    try:
        fm = T.framemapperfint_cc(*fm_args(cfg))
        assert fm.warnings == 1
        ok = T.framemapperfint_cc(*fm_args(K.resolve("c1")))
        assert ok.warnings == 0
        assert fm.output_multiple == ok.output_multiple          # the physical frame does not grow
        code = fm.plan("frame.code", np.int32)
        info = fm.plan("frame.info", np.int32)
        assert code.size == fm.output_multiple and info[6] == 0   # no dummy cells
        # every input cell index that survives the truncation is unique, and some cells are dropped
        data = code[code >= 0]
        assert np.unique(data).size == data.size and data.size < fm.forecast(fm.output_multiple)
    finally:
        T.set_overfull_policy(False)


def test_link_rejects_ts_consumers_and_mismatched_items():
    cfg = K.resolve("c1")
    B = T.blocks_for(cfg)
    B["bb"].link_to(B["ldpc"])
    B["im"].link_to(B["fm"])
    with pytest.raises(ValueError):
        B["ldpc"].link_to(B["fm"])           # bytes -> complex cells
    with pytest.raises(ValueError):
        B["ldpc"].link_to(B["bb"])           # the TS consumer keeps stream history in front of its input


def test_ts_ingest_helpers():
    """dvbt2ll_ts_sync / dvbt2ll_ts_fill: sync acquisition on a stream that starts inside a packet, and completion of a
    stream that has run dry with null packets (PID 0x1FFF) on the packet grid."""
    ts = K.make_ts(188 * 40)
    junk = np.full(77, 0x47, np.uint8)
    junk[::3] = 0x12                                  # stray 0x47 bytes that do not repeat at the packet stride
    stream = np.concatenate([junk, ts[:188 * 20 + 50]])
    assert T.ts_sync(stream) == 77
    assert T.ts_sync(np.zeros(1000, np.uint8)) == -1
    aligned = stream[77:]
    out, nulls = T.ts_fill(aligned, 188 * 18 - 5, 188 * 5 + 5)
    whole = 188 * 20                                  # the trailing partial packet (50 bytes) is dropped
    assert np.array_equal(out[:188 * 2 + 5], aligned[188 * 18 - 5:whole])
    rest = out[188 * 2 + 5:]
    assert nulls == rest.size == 188 * 3
    pk = rest.reshape(3, 188)
    assert np.all(pk[:, 0] == 0x47) and np.all(pk[:, 1] == 0x1F) and np.all(pk[:, 2] == 0xFF) and np.all(pk[:, 3] == 0x10)
    assert np.all(pk[:, 4:] == 0xFF)
    hist, n0 = T.ts_fill(aligned, -187, 200)
    assert n0 == 0 and not hist[:187].any() and np.array_equal(hist[187:], aligned[:13])
