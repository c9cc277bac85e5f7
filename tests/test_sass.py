"""The built library is Blackwell-native code, checked on the artefact itself: ``cuobjdump -sass`` of ``libpcb200.so`` must show
only ``sm_100a`` cubins and the instructions the design claims — ``UTCHMMA`` (tcgen05.mma), ``LDTM`` (tcgen05.ld from tensor
memory), ``UTCBAR`` (tcgen05.commit -> mbarrier), ``UTMALDG`` (TMA bulk tensor loads of the stencil bricks), ``LDGSTS``
(cp.async staging of GEMM operands), ``SYNCS`` (mbarrier) — and no Hopper ``HGMMA`` / legacy ``HMMA`` path."""

import re
import shutil
import subprocess

import pytest

from pytorch_connectomics_b200 import _lib


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_library_sass_is_sm100a_tcgen05_and_tma():
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    arches = set(re.findall(r"arch = (sm_\w+)", sass))
    assert arches == {"sm_100a"}, arches
    count = {m: len(re.findall(rf"\b{m}\b", sass)) for m in ("UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "LDGSTS", "SYNCS", "HGMMA")}
    print(count)
    assert count["UTCHMMA"] >= 600 and count["LDTM"] >= 80 and count["UTCBAR"] >= 80      # tcgen05 GEMM kernels, TMEM epilogues
    assert count["UTMALDG"] >= 16 and count["LDGSTS"] >= 90 and count["SYNCS"] >= 500     # TMA bricks, cp.async staging, mbarriers
    assert count["HGMMA"] == 0 and not re.search(r"\bHMMA\.", sass)                       # no wgmma / mma.sync fallbacks
    kernels = set(re.findall(r"Function : (\w+)", sass))
    for needle in ("mlp_fused_kernel", "mlp_bwd_ws_kernel", "mlp_bwd_ws2_kernel", "gemm_ws_kernel", "tn_gemm_ws_kernel",
                   "dwconv_same_tiled_kernel", "dw_wgrad_same_tiled_kernel", "conv_igemm_kernel", "tta_fold_kernel", "adamw_kernel"):
        assert any(needle in k for k in kernels), needle
