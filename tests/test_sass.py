"""The built library contains what DESIGN.md says it does: sm_100a code only, 1-D bulk async copies (UBLKCP) completing on
mbarriers (SYNCS) in the SpMM kernels that claim them, and fp64 arithmetic without fused multiply-add in the units compiled
with -fmad=false (the bit-exactness argument rests on that).  Needs cuobjdump (CUDA toolkit), no GPU."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def sass_by_function(obj):
    out = subprocess.run([CUOBJDUMP, "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    return out, {k: "\n".join(v) for k, v in funcs.items()}


@pytest.fixture(scope="module")
def apply_sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    from icebin_b200 import build
    build.build()
    outs, funcs = "", {}
    for unit in ("apply.o", "apply_stream.o", "apply_group.o"):
        obj = os.path.join(ROOT, "icebin_b200", "build", unit)
        if not os.path.exists(obj):
            pytest.skip("object files not kept")
        o, f = sass_by_function(obj)
        outs += o
        funcs.update(f)
    return outs, funcs


def test_only_sm100a_code(apply_sass):
    out, _ = apply_sass
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("kernel", ["spmm_gc_kernel", "spmm_tma_kernel", "spmm_stream_kernel", "smooth_group_kernel"])
def test_bulk_async_copies_and_mbarriers(apply_sass, kernel):
    _, funcs = apply_sass
    mine = {k: v for k, v in funcs.items() if kernel in k}
    assert mine, kernel
    for name, body in mine.items():
        assert "UBLKCP" in body, name + ": no bulk async copy (cp.async.bulk)"
        assert "SYNCS" in body, name + ": no mbarrier operations"


def test_no_fused_multiply_add_in_the_spmm_kernels(apply_sass):
    """Rows are accumulated as acc = acc + v*x with separate roundings (Eigen's order on the CPU): no DFMA."""
    _, funcs = apply_sass
    for name, body in funcs.items():
        if "spmm" in name:
            assert "DFMA" not in body, name
            if "fixup" not in name:
                assert "DMUL" in body and "DADD" in body, name


def test_group_kernel_runs_on_the_fp64_tensor_path(apply_sass):
    """smooth_group_kernel multiplies its 8 x 4 blocks with DMMA (unpredicated: a predicated-off DMMA still occupies the
    pipe), hands the producer warpgroup's registers to the consumers and does not spill."""
    _, funcs = apply_sass
    mine = {k: v for k, v in funcs.items() if "smooth_group_kernel" in k}
    assert mine
    for name, body in mine.items():
        assert "DMMA.8x8x4" in body, name
        assert not re.search(r"@!?U?P\d+\s+DMMA", body), name + ": predicated DMMA"
        assert "USETMAXREG.DEALLOC" in body and "USETMAXREG.TRY_ALLOC" in body, name
        assert "STL" not in body and "LDL" not in body, name + ": register spills"
