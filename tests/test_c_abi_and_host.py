"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the ctypes structs have the C layout, and the host half of solve_min_cost_flow (option
resolution, canonical arrays, error behaviour) mirrors the reference - no compute call is made."""

import ctypes
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402
from network_flow_solver_b200 import (  # noqa: E402
    SolverConfigurationError, SolverOptions, build_problem, solve_min_cost_flow, _capi)
from network_flow_solver_b200 import generators as gen  # noqa: E402
from network_flow_solver_b200.canonical import PERTURB_GROWTH, initial_block_size  # noqa: E402
from network_flow_solver_b200.exceptions import InvalidProblemError  # noqa: E402
from network_flow_solver_b200.solver import prepare  # noqa: E402


def test_library_exports_every_declared_symbol():
    lib_path = entry.build_engine()
    lib = ctypes.CDLL(str(lib_path))
    symbols = entry.declared_symbols()
    assert {"nsx_solve", "nsx_solve_warm", "nsx_solve_resident", "nsx_solve_batch", "nsx_solve_sharded", "nsx_sweep_probe",
            "nsx_mailbox_create", "nsx_mailbox_open", "nsx_mailbox_bytes", "nsx_last_error", "nsx_version"} <= set(symbols)
    for name in symbols:
        assert hasattr(lib, name), name
    abi, arch = ctypes.c_int32(), ctypes.c_int32()
    lib.nsx_version(ctypes.byref(abi), ctypes.byref(arch))
    assert (abi.value, arch.value) == (_capi.ABI_VERSION, 100) == (4, 100)


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of the three ABI structs, as the C compiler sees them."""
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "nsx_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(nsx_problem), sizeof(nsx_options), sizeof(nsx_result),'
        ' offsetof(nsx_result, status), offsetof(nsx_result, solve_ms), offsetof(nsx_result, phase_cycles),'
        ' offsetof(nsx_result, handshake_ns));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    R = _capi.NsxResult
    want = [ctypes.sizeof(_capi.NsxProblem), ctypes.sizeof(_capi.NsxOptions), ctypes.sizeof(R),
            R.status.offset, R.solve_ms.offset, R.phase_cycles.offset, R.handshake_ns.offset]
    assert got == want


def test_no_cpu_fallback_when_the_library_is_missing(monkeypatch, tmp_path):
    monkeypatch.setenv("NSX_B200_LIB", str(tmp_path / "absent.so"))
    monkeypatch.setattr(_capi, "_lib", None)
    from network_flow_solver_b200.exceptions import DeviceEngineError

    with pytest.raises(DeviceEngineError):
        _capi.load_library()
    monkeypatch.setattr(_capi, "_lib", None)


def small_problem():
    nodes = [{"id": "a", "supply": 4}, {"id": "b", "supply": 0}, {"id": "c", "supply": -4}]
    arcs = [{"tail": "a", "head": "b", "capacity": 5, "cost": 2.0}, {"tail": "b", "head": "c", "capacity": 5, "cost": 1.0},
            {"tail": "a", "head": "c", "capacity": 2, "cost": 5.0, "lower": 1}]
    return build_problem(nodes, arcs, directed=True, tolerance=1e-6)


def test_canonical_arrays_follow_the_reference_index_space():
    cp, plan, _ = prepare(small_problem(), SolverOptions(pricing_strategy="dantzig", explicit_pricing_strategy=True, auto_scale=False))
    assert cp.node_ids == ["__root__", "a", "b", "c"] or cp.node_ids[1:] == ["a", "b", "c"]
    assert cp.arc_keys == [("a", "b"), ("a", "c"), ("b", "c")]           # stable (tail, head) order, simplex.py:392-395
    np.testing.assert_array_equal(cp.upper, [5.0, 1.0, 5.0])             # lower bound shifted out, simplex.py:416-428
    np.testing.assert_array_equal(cp.supply[1:], [3.0, 0.0, -3.0])
    growth = np.cumprod([1.0, PERTURB_GROWTH, PERTURB_GROWTH])
    np.testing.assert_array_equal(cp.pert_cost, np.array([2.0, 5.0, 1.0]) + 1e-10 * growth)  # simplex.py:1431-1440
    assert cp.penalty == 5.0 * (3 + 2)                                   # max|c| * (N + 2), simplex.py:161-163
    assert plan.engine.pricing == _capi.PRICING_DANTZIG and plan.engine.max_iterations == max(100, 20 * (3 + 3))


def test_option_resolution_matches_the_reference_rules():
    arrays = gen.transportation(6, 7, cost_max=20, seed=1)
    problem = gen.to_network_problem(arrays)
    cp, plan, _ = prepare(problem, SolverOptions(pricing_strategy="devex", explicit_pricing_strategy=True, auto_scale=False))
    assert plan.engine.row_scan_first and plan.engine.pricing == _capi.PRICING_DEVEX   # override, simplex.py:1060-1064
    assert plan.engine.block_size == initial_block_size(cp.n_arcs) and plan.engine.auto_block
    goto = gen.to_network_problem(gen.goto_like(16, seed=1))
    _, plan2, _ = prepare(goto, SolverOptions(pricing_strategy="devex", auto_scale=False))          # not explicit
    assert plan2.strategy == "dantzig"                                                               # simplex.py:358-374


def test_paths_outside_the_accelerated_scope_fail_loudly():
    options = SolverOptions(explicit_pricing_strategy=True, auto_scale=False)
    object.__setattr__(options, "pricing_strategy", "steepest")  # past the constructor's own check (data.py:163)
    with pytest.raises(SolverConfigurationError):
        prepare(small_problem(), options)


def test_loop_based_devex_has_its_own_engine_rule():
    _, plan, _ = prepare(small_problem(), SolverOptions(pricing_strategy="devex", explicit_pricing_strategy=True, auto_scale=False,
                                                         use_vectorized_pricing=False))
    assert plan.engine.pricing == _capi.PRICING_DEVEX_LOOP


def test_unbalanced_problem_is_rejected_like_the_reference():
    nodes = [{"id": "a", "supply": 4}, {"id": "b", "supply": -3}]
    arcs = [{"tail": "a", "head": "b", "capacity": 5, "cost": 2.0}]
    with pytest.raises(InvalidProblemError):
        solve_min_cost_flow(build_problem(nodes, arcs, directed=True, tolerance=1e-6),
                            SolverOptions(pricing_strategy="dantzig", explicit_pricing_strategy=True, auto_scale=False))


def test_library_is_sm100a_with_tma_and_mbarrier_in_the_resident_kernels():
    """What the GPU box will load: an sm_100a cubin whose resident kernels move tiles with TMA bulk copies (UBLKCP) counted
    on mbarriers (SYNCS) - not a PTX-only or an LDG-loop build - and no statically linked CUDA runtime entry points."""
    import shutil

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    lib = entry.build_engine()
    elf = subprocess.run([cuobjdump, "-lelf", str(lib)], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "nsx_resident_kernel", str(lib)], capture_output=True, text=True).stdout
    assert sass.count("UBLKCP.S.G") >= 8 and sass.count("SYNCS.PHASECHK") >= 8 and "SYNCS.ARRIVE" in sass
    assert "DSETP" in sass and "DADD" in sass  # float64 arithmetic of the reference, not a reduced-precision path
    # -fmad=false: the reference's `(c + pi[t]) - pi[h]` is never contracted into a fused multiply-add.  The only DFMAs in
    # the kernel are the Newton steps of the correctly rounded divisions (__ddiv_rn: Devex merit rc^2 / w), i.e. they sit
    # next to a MUFU.RCP64H
    lines = sass.splitlines()
    rcp = [i for i, ln in enumerate(lines) if "MUFU.RCP64H" in ln]
    stray = [i for i, ln in enumerate(lines) if "DFMA" in ln and not any(abs(i - r) <= 120 for r in rcp)]
    assert rcp and not stray, f"{len(stray)} DFMA instructions outside division sequences"
    nm = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True).stdout
    assert "cudaMemcpy" not in nm and "cudaLaunchKernel" not in nm
