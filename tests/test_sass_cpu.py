"""Static checks on the SASS of the built library (no GPU needed: cuobjdump reads the sm_100a cubin).

They pin what DESIGN.md claims about the code the hot kernels execute: the FP64 instruction count
per point*mode (W_exec = D + 5 + DEG + NC, DEG = polynomial degree), the instruction diet of the mode loop, the TMA bulk copy /
mbarrier staging of the mode records, and the FP64 tensor instruction of the grid and kriging GEMMs."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_budget  # noqa: E402

LIB = os.path.join(ROOT, "gstools-core_b200", "gstools_core", "libgsfield.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                                reason="needs cuobjdump and the built library")


# (D, NC, P) of the variants the five BASELINE configs run, with the unroll factor of gsf_unroll()
@pytest.mark.parametrize("deg", [5, 6])               # throughput degree / high degree (gsf_kernels.cuh)
@pytest.mark.parametrize("d,nc,p,unroll", [(3, 1, 3, 16), (2, 1, 3, 4), (3, 3, 3, 16), (3, 1, 1, 16), (2, 1, 1, 8)])
def test_mode_loop_instruction_budget(d, nc, p, unroll, deg):
    r = sass_budget.analyze(LIB, d, nc, p, deg)
    pm = unroll * p                                   # point*modes per trip of the unrolled loop
    mix = r["mix"]
    assert r["fp64"] == (d + 5 + deg + nc) * pm       # W_exec, DESIGN.md section 2
    assert mix.get("DMUL", 0) == pm                   # s = r*r
    assert mix.get("DADD", 0) == 4 * pm               # 3 (range reduction) + 1 (monic Horner start)
    assert mix.get("IMAD", 0) <= pm + 2 * unroll      # the sign, one integer op per point*mode
    rec_bytes = 8 * ((d + 1 + nc + 1) & ~1)
    assert mix.get("LDS", 0) <= unroll * ((rec_bytes + 15) // 16)   # one pass over each mode record per trip
    for banned in ("LDC", "LDG", "LDL", "STL", "MUFU", "BAR"):      # no constant re-loads, spills or calls
        assert banned not in mix, (banned, mix)
    # non-FP64 issue slots: below 10 % of the cycles with P = 3 (LDS and loop amortised over the points)
    assert r["other"] <= (0.2 if p > 1 else 0.34) * r["fp64"]
    assert r["pipe_frac"] >= (0.82 if p > 1 else 0.74)   # static model; ncu measures 3-4 points more


@pytest.mark.parametrize("deg", [5, 6])
def test_polynomial_has_no_three_register_instruction(deg):
    """The monic form keeps one constant per polynomial instruction: every
    DFMA/DADD of the Horner chain and the double-angle step reads at most two distinct registers."""
    body = sass_budget.hottest_loop(sass_budget.kernel_sass(LIB, 3, 1, 3, deg))
    n_const = 0
    for _, text in body:
        if sass_budget.opcode(text) in ("DFMA", "DADD") and ("UR" in text.split(None, 1)[1] or "c[" in text):
            regs = {s[0] for s in sass_budget.sources(text) if s is not None}
            assert len(regs) <= 2, text
            n_const += 1
    assert n_const == (deg + 1) * 48                  # DADD + (deg-1) DFMA + double angle, 48 point*modes per trip (unroll 16, P = 3)


def test_mode_records_are_staged_by_tma_bulk_copies():
    ops = [sass_budget.opcode(t) for _, t in sass_budget.kernel_sass(LIB, 3, 1, 3)]
    assert "UBLKCP" in ops                            # cp.async.bulk.shared::cluster.global (TMA engine)
    assert "SYNCS" in ops                             # mbarrier expect-tx / try_wait


@pytest.mark.parametrize("name", ["gsf_grid_gemm", "gsf_krige_gemm"])
def test_fp64_gemms_use_the_tensor_instruction(name):
    ops = [t for _, t in sass_budget.function_sass(LIB, name)]
    assert ops, name
    assert any(t.startswith("DMMA") for t in ops)
