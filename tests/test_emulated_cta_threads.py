"""The device pivot source as a CTA of REAL threads: nsx_core.cuh compiled with NSX_HOST_MT (tests/emu), NSX_SYNC = a thread
barrier, NSX_PAR_FOR strided over the threads, NSX_SINGLE = thread 0.  What the serial emulation cannot show and a GPU run
only shows as a hang or a wrong answer: a barrier that part of the CTA skips (deadlock here) and shared state touched
without a barrier in between (ThreadSanitizer report here).  Covers the code common to device and host builds - driver,
pivot bookkeeping, flow / tree update, candidate-list scan, structure-rule and loop-Devex scans, warm-start paths; the warp-level
device sections (cycle walk, ratio test reduction, potential wavefront) have host stand-ins run by one thread.
Runs in a subprocess so that the sanitizer runtime can be preloaded."""

import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SELECT = ("emulated and (assignment_6 or shortest_path_20 or max_flow_24 or netgen_64 or transport_32 or small_ or "
          "uncap_64_costs or overpush or transport_16 or candidate)")
FILES = ["tests/test_next_special_pivots.py", "tests/test_next_devex_loop.py", "tests/test_next_warm_start.py",
         "tests/test_next_candidate_list.py"]


def run_suite(lib: Path, env_extra: dict, tmp_path: Path):
    env = {**os.environ, "NSX_EMU_LIB": str(lib), "NSX_EMU_THREADS": "4", **env_extra}
    return subprocess.run([sys.executable, "-m", "pytest", *FILES, "-x", "-q", "-m", "not gpu", "-k", SELECT, "-p", "no:cacheprovider"],
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)


@pytest.mark.slow
@pytest.mark.parametrize("lazy", ["0", "1"])
def test_cta_of_real_threads_neither_deadlocks_nor_races(tmp_path, lazy):
    from emu import emu

    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not tsan or not Path(tsan).exists():
        pytest.skip("libtsan not available")
    lib = emu.build_mt(tmp_path / "libnsx_emu_tsan.so", sanitize=True)
    log = tmp_path / "tsan.log"
    proc = run_suite(lib, {"LD_PRELOAD": tsan, "TSAN_OPTIONS": f"report_signal_unsafe=0 exitcode=0 log_path={log}",
                           "NSX_EMU_BLOCKED": lazy, "NSX_EMU_BLK_LG": "2" if lazy == "1" else "", "NSX_EMU_BLK_NB": "6" if lazy == "1" else ""}, tmp_path)
    if "unexpected memory mapping" in proc.stderr or "unexpected memory mapping" in proc.stdout:
        pytest.skip("ThreadSanitizer cannot map its shadow memory on this kernel (ASLR layout)")
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]  # a deadlock ends in the timeout above
    assert " passed" in proc.stdout and "failed" not in proc.stdout
    reports = [p.read_text() for p in tmp_path.glob("tsan.log*")]
    assert not any("WARNING: ThreadSanitizer" in r for r in reports), reports[0][:3000]
