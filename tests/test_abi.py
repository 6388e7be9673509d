"""CPU checks of the drop-in boundary: libdurf_b200.so loads, exports every symbol include/durf_b200.h declares, and the
ctypes binding covers exactly that set.  No compute calls (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "durf_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    return sorted(set(re.findall(r"\b(durf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared_functions()
    for must in ("durf_obb_frontend_fwd", "durf_obb_frontend_bwd", "durf_raymarch_fwd", "durf_raymarch_bwd", "durf_viewdir_enc_fwd",
                 "durf_mlp_fwd", "durf_mlp_bwd", "durf_composite_fwd", "durf_composite_bwd", "durf_resample_fwd", "durf_losses_prepare",
                 "durf_losses_fwd_bwd", "durf_grad_sanitize", "durf_adam_step", "durf_version", "durf_last_error"):
        assert must in names, f"{must} missing from include/durf_b200.h"


def test_library_exports_every_declared_symbol_and_binding_matches():
    from durf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_functions()
    for name in declared:
        assert hasattr(lib, name), f"libdurf_b200.so does not export {name}"
    assert sorted(_lib.SIGNATURES) == declared, (
        f"ctypes binding and header disagree: {sorted(set(_lib.SIGNATURES) ^ set(declared))}")
    bound = _lib.load()
    assert bound.durf_version().decode().startswith("durf_b200")
    assert bound.durf_last_error() is not None


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing on the host."""
    import torch
    from durf_b200 import _lib, ops
    with pytest.raises(_lib.DurfError):
        ops.viewdir_enc(torch.zeros(4, 3))
    with pytest.raises(_lib.DurfError):
        ops.composite(torch.zeros(2, 8, 3), torch.zeros(2, 8), torch.zeros(2, 9), torch.zeros(2, 3))


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under durf_b200/ may import it."""
    pkg = os.path.join(ROOT, "durf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_size_queries_are_host_functions_and_consistent():
    """durf_mlp_*_bytes are pure host arithmetic (callable without a GPU).  Tensor-core training keeps per ray-level every
    layer's bf16 activations as 16 KB block images plus the trunk layers' 1-bit ReLU masks (4 bytes per row and 32-column
    group); an unsupported topology reports 0 (the caller must then use the fp32 path)."""
    import ctypes as C
    from durf_b200 import _lib
    lib = _lib.load()
    bg = _lib.MlpTopology(60, 256, 8, 4, 27, 128)
    obj = _lib.MlpTopology(63, 128, 8, 4, 27, 128)
    for t, width in ((bg, 256), (obj, 128)):
        blocks = (8 + 1) * (width // 64) + 128 // 64
        for M in (0, 1, 300):
            want = M * blocks * 16384 + M * (8 + 1) * (width // 32) * 128 * 4     # 8 trunk layers + the condition layer
            assert lib.durf_mlp_saved_bytes(C.byref(t), _lib.PREC_BF16, M, 128) == want
            # backward workspace = the dZ tile records, same shape as the saved activations
            assert lib.durf_mlp_workspace_bytes(C.byref(t), _lib.PREC_BF16, M, 128, 1) == M * blocks * 16384
            # the tensor-core forward needs no workspace (the per-tile view bias is formed inside the kernel)
            assert lib.durf_mlp_workspace_bytes(C.byref(t), _lib.PREC_BF16, M, 128, 0) == 0
        assert lib.durf_mlp_packed_bytes(C.byref(t)) > 0
    odd = _lib.MlpTopology(60, 192, 8, 4, 27, 128)          # width the tensor-core path does not implement
    assert lib.durf_mlp_saved_bytes(C.byref(odd), _lib.PREC_BF16, 4, 128) == 0
    assert lib.durf_mlp_saved_bytes(C.byref(odd), _lib.PREC_FP32, 4, 128) > 0
