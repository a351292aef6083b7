"""Host-side pieces of bench.py that need no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_watchdog_reports_stage_and_exits():
    """A run that stops making progress prints one JSON line naming the stage it was in (rank 0) and exits non-zero."""
    code = ("import sys, time; sys.path.insert(0, %r); import bench; bench.stage('unit test stage'); "
            "bench.start_watchdog(1, 0, 2); time.sleep(30)" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=25)
    assert r.returncode == 4
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 0.0 and line["n_gpus"] == 2 and "unit test stage" in line["error"]
    # other ranks exit silently
    code2 = code.replace("start_watchdog(1, 0, 2)", "start_watchdog(1, 1, 2)")
    r2 = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, timeout=25)
    assert r2.returncode == 4 and r2.stdout.strip() == ""
