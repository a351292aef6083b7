#!/usr/bin/env python
"""Regenerate the five-configuration table of README.md from bench.py JSON lines.

usage: tools/bench_table.py <log with the N=1 line> [<log with an N>1 line> ...]

The N = 1 line carries config 5 (the headline: 64 x c3 channels) and `per_config` (c1, c2, c4 = BASELINE.json
configs[0], [1], [3]; config 3 is the per-channel configuration of config 5).  Lines of multi-GPU runs add the
scaling rows.  The table replaces the text between the BENCH-TABLE markers of README.md."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"c1": "1: apps/vv009-4kshort.grc (4K, short FECFRAME, 256QAM rot, CR 4/5)",
         "c2": "2: 8K, normal FECFRAME, QPSK CR 1/2, GI 1/4, PP1",
         "c4": "4: 16K, 64QAM rot CR 3/5, GI 1/16, PP4, TI depth 3",
         "c5": "3 / 5: 64 x (32K ext, 256QAM rot CR 2/3, GI 1/128, PP7)"}


def last_json(path):
    line = None
    for ln in open(path):
        if ln.startswith("{"):
            line = ln
    return json.loads(line) if line else None


def main():
    lines = [last_json(p) for p in sys.argv[1:]]
    lines = [d for d in lines if d]
    one = [d for d in lines if d.get("n_gpus") == 1][0]
    rows = ["| BASELINE.json config | channels x T2 frames / step | device Msamples/s | x real time | FECFRAMEs/s | stage ms (bb_bch, ldpc, map, ofdm) | k_ofdm roofline (of measured HBM copy) | reference, 1 host core Msamples/s |",
            "|---|---|---|---|---|---|---|---|"]

    def stage(st):
        return ", ".join("%.3f" % st[k] for k in ("bb_bch", "ldpc", "map", "ldpc_map", "ofdm") if k in st)

    for name in ("c1", "c2", "c4"):
        e = one.get("per_config", {}).get(name)
        if not e or "value" not in e:
            continue
        rows.append("| %s | %s | %.0f | %.0f | %.3g | %s | %.2f | %s |" % (
            NAMES[name], e["workload"].split(" channels")[0].replace(" x " + name, "") + " x " + e["workload"].split(" x ")[-1].split(" T2")[0],
            e["value"], e["x_realtime"], e["fecframes_per_s"], stage(e["roofline"]["stage_ms"]), e["roofline"]["frac"],
            "%.1f" % e["cpu_reference_1core"]["value"] if "cpu_reference_1core" in e else "-"))
    rows.append("| %s | %d x %d | %.0f | %.0f | %.3g | %s | %.2f | %s |" % (
        NAMES["c5"], one["config"]["channels"], one["config"]["t2_frames_per_channel_per_step"], one["value"], one["x_realtime"],
        one["fecframes_per_s"], stage(one["roofline"]["stage_ms"]), one["roofline"]["frac"],
        "%.1f" % one["cpu_baseline"]["value"] if "cpu_baseline" in one else "-"))
    out = ["\n".join(rows), ""]
    out.append("End to end through host buffers (`e2e`, N = 1): %.0f Msamples/s (%.1f GB/s over PCIe, %.0f %% of the box's measured "
               "device-to-host ceiling); with the fused int16 sink %.0f Msamples/s." % (
                   one["e2e"]["value"], one["e2e"]["pcie_gbs_achieved"], 100 * one["e2e"]["frac_of_ceiling"], one["e2e_int16_sink"]["value"]))
    d = one.get("dropin_e2e", {})
    if "pageable" in d:
        out.append("Drop-in per-block path (five `dvbt2ll_work` handles, one c3 T2 frame per round): %.0f Msamples/s on pageable buffers, "
                   "%.0f with the buffers registered on first sight, %.0f with the device-resident hand-off as well%s." % (
                       d["pageable"]["value"], d["host_register"]["value"], d["host_register+link"]["value"],
                       ", %.0f with the host copies of the linked edges left out" % d["host_register+link_lazy_host"]["value"]
                       if "host_register+link_lazy_host" in d else ""))
        k1, k2 = "host_register+link, thread per block", "host_register+link_lazy_host, thread per block"
        if k1 in d and k2 in d:
            out.append("The same handles driven by one thread per block over three-slot rings (GNU Radio's scheduler model): %.0f Msamples/s "
                       "linked, %.0f with lazy host output." % (d[k1]["value"], d[k2]["value"]))
    pb = one.get("per_block_device", {})
    if "blocks" in pb:
        parts = []
        for name, e in pb["blocks"].items():
            t = "`%s` %.3f ms (%.0f GB/s of items" % (name, e["ms_per_call"], e["item_gbs"])
            if "algorithmic" in e:
                t += "; SURVEY 8(d) bytes at %.2f of the copy rate" % e["algorithmic"]["frac_of_peak"]
            parts.append(t + ")")
        out.append("The five drop-in blocks on device-resident items (`dvbt2ll_work_device`, %s): %s." % (
            pb["workload"].split(",")[0], ", ".join(parts)))
    multi = sorted([x for x in lines if x.get("n_gpus", 1) > 1], key=lambda x: x["n_gpus"])
    if multi:
        out.append("")
        out.append("| GPUs (64 channels in total, 64 / N each) | Msamples/s with the ordered reassembly on GPU 0 (`value`) | ms / step | NVLink ingest of GPU 0, GB/s (bound 770) | same step without the reassembly | reassembly with the int16 sink | reassembly with root-weighted shares (channels on GPU 0) | e2e host buffers Msamples/s (PCIe GB/s of ceiling) |")
        out.append("|---|---|---|---|---|---|---|---|")
        out.append("| 1 | %.0f | %.3f | - | %.0f | - | - | %.0f (%.0f of %.0f) |" % (
            one["value"], one["ms_per_step"], one["value"], one["e2e"]["value"], one["e2e"]["pcie_gbs_achieved"], one["e2e"]["pcie_ceiling_gbs"]))
        for m in multi:
            mg = m["multi_gpu"]
            rw = mg.get("reassembly_root_weighted")
            out.append("| %d | %.0f | %.3f | %.0f | %.0f | %.0f | %s | %.0f (%.0f of %.0f) |" % (
                m["n_gpus"], m["value"], m["ms_per_step"], mg["reassembly"]["ingest_gbs_achieved"], mg["compute_only"]["value"],
                mg["reassembly_int16_sink"]["value"], "%.0f (%d)" % (rw["value"], rw["channels_per_gpu"][0]) if rw else "-",
                m["e2e"]["value"], m["e2e"]["pcie_gbs_achieved"], m["e2e"]["pcie_ceiling_gbs"]))
    text = "\n".join(out)
    p = os.path.join(ROOT, "README.md")
    s = open(p).read()
    s2, n = re.subn(r"(<!-- BENCH-TABLE:BEGIN[^>]*-->\n).*?(<!-- BENCH-TABLE:END -->)", lambda m: m.group(1) + text + "\n" + m.group(2), s, flags=re.S)
    if n != 1:
        raise SystemExit("README.md has no BENCH-TABLE markers")
    open(p, "w").write(s2)
    print(text)


if __name__ == "__main__":
    main()
