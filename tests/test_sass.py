"""Static checks on the built library's SASS (cuobjdump, no GPU): the tensor-core kernels really are tcgen05 / TMEM / bulk-copy
kernels (the mnemonics /opt/skills/guides/B200_PROFILING.md names as proof), and the code-generation properties the last
kernel commits of round 2 were measured on stay in place - they are easy to lose with an innocent-looking edit:

* the forward kernel's cycle counters are compiled out of the default build (only the spin-loop watchdogs read the clock);
* the warp index is warp-uniform for the compiler: no per-lane loop (BRA.U.ANY) around a bulk-store issue in the forward
  kernel's epilogue / generator warps;
* dgrad / wgrad address shared memory as shared memory (no generic LD.E / ST.E).
"""
import os
import re
import shutil
import subprocess

import pytest

from durf_b200 import _lib

FWD = "_ZN4durf17mlp_tc_fwd_kernelILi256ELb0EEEvNS_8TcParamsE"
FWD_SAVE = "_ZN4durf17mlp_tc_fwd_kernelILi256ELb1EEEvNS_8TcParamsE"
DGRAD = "_ZN4durf19mlp_tc_dgrad_kernelILi256EEEvNS_8DgParamsE"
WGRAD = "_ZN4durf19mlp_tc_wgrad_kernelENS_11WgradParamsE"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump (CUDA toolkit) not on PATH")


def _sass(function):
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    out = subprocess.run(["cuobjdump", "-sass", "-fun", function, _lib.LIB_PATH], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
    assert len(lines) > 1000, f"{function}: not found in {_lib.LIB_PATH} (sm_100a cubin missing?)"
    return lines


def _count(lines, pattern):
    rx = re.compile(pattern)
    return sum(1 for l in lines if rx.search(l))


@pytest.mark.parametrize("function", [FWD, FWD_SAVE, DGRAD, WGRAD])
def test_tensor_core_kernels_are_tcgen05_tmem_and_bulk_copy_kernels(function):
    s = _sass(function)
    assert _count(s, r"\bUTCHMMA\b") >= 8, "no tcgen05.mma (UTCHMMA) in the kernel"
    assert _count(s, r"\bUTCBAR\b") >= 1, "no tcgen05.commit (UTCBAR)"
    assert _count(s, r"\bUBLKCP\b") >= 1, "no cp.async.bulk (UBLKCP)"
    assert _count(s, r"\bLDTM\b") >= 1, "no tcgen05.ld (LDTM): accumulators are not read from tensor memory"
    assert _count(s, r"\bHMMA\b|\bWGMMA\b") == 0, "legacy mma.sync / wgmma instructions in a tcgen05 kernel"


def test_forward_activations_stay_in_tensor_memory_and_weights_are_multicast():
    s = _sass(FWD)
    assert _count(s, r"\bSTTM\b") >= 1, "no tcgen05.st (STTM): the next layer's A operand does not go to tensor memory"
    assert _count(s, r"UBLKCP\.S\.G\.MULTICAST") >= 1, "the weight ring is not multicast to the CTA pair"
    assert _count(s, r"UTCBAR\.MULTICAST") >= 1, "ring stages are not released with a multicast commit"


def test_default_build_has_no_cycle_counters_in_the_forward_kernel():
    # what remains are the watchdogs of the barrier spin loops (slow path) and two one-off reads at kernel start; a build
    # with -DDURF_TRACE=1 has several dozen, most of them inside the epilogue loop
    for f in (FWD, FWD_SAVE):
        assert _count(_sass(f), r"SR_CLOCKLO|SR_CLOCKHI") <= 10, \
            "clock reads in the default build: DURF_TRACE must be 0 (even switched off at run time they cost 5 %)"


def _store_issues_in_a_lane_loop(lines):
    """Bulk STORES (shared -> global: the saved activations, the dZ records, the generated tiles) whose issue sits in a
    per-lane loop: the compiler emits `UBLKCP ...; @P BRA.U.ANY back` when it cannot prove the address warp-uniform."""
    n = 0
    for i, l in enumerate(lines):
        if re.search(r"UBLKCP\.G\.S", l) and any("BRA.U.ANY" in x for x in lines[i + 1:i + 4]):
            n += 1
    return n


def test_warp_uniform_code_generation():
    # the epilogue / generator warps' bulk stores are issued from uniform registers (the weight producer's loads, issued by
    # one lane with slack to spare, may keep their loop)
    for f in (FWD, FWD_SAVE):
        s = _sass(f)
        assert _count(s, r"UBLKCP\.G\.S") >= 1
        assert _store_issues_in_a_lane_loop(s) == 0, f"per-lane loop around a bulk store in {f}"
    for f in (DGRAD, WGRAD):
        assert _count(_sass(f), r"\bLD\.E\b|\bST\.E\b|\bATOM\.E\b") == 0, f"generic memory accesses in {f}"
