"""The C-ABI shared library loads on a CPU-only box and exports exactly the symbols include/l4p_b200.h declares
(no compute calls: there is no GPU here). Also checks the documented error behaviour of the entry points that can
be exercised without a device."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "l4p_b200.h"


def header_symbols():
    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(l4p_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from l4p_b200 import build, lib as L

    build.build()  # nvcc cross-compiles for sm_100a without a GPU
    return L.load()


def test_header_and_binding_tables_agree():
    from l4p_b200 import lib as L

    assert header_symbols() == L.exported_symbols()


def test_library_exports_every_declared_symbol(lib):
    from l4p_b200 import lib as L

    out = subprocess.run(["nm", "-D", "--defined-only", str(L.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (l4p_[a-z0-9_]+)\b", out))
    for sym in header_symbols():
        assert sym in exported, f"{sym} declared in the header but not exported"
        assert getattr(lib, sym) is not None


def test_gemm_desc_struct_matches_header():
    """Field order/count of the ctypes mirror == the C struct (guards against silent ABI drift)."""
    from l4p_b200 import lib as L

    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    body = re.search(r"typedef struct l4p_gemm_desc \{(.*?)\} l4p_gemm_desc;", txt, flags=re.S).group(1)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
    assert names == [f[0] for f in L.GemmDesc._fields_]


def test_error_reporting_without_gpu(lib):
    assert lib.l4p_version() >= 100
    rc = lib.l4p_init(0, 1)  # no CUDA device in the build container
    import torch

    if not torch.cuda.is_available():
        assert rc < 0
        assert b"CUDA" in lib.l4p_last_error() or b"device" in lib.l4p_last_error()
    # argument validation happens before any device work
    assert lib.l4p_gemm(None, None) == -1
    assert b"null descriptor" in lib.l4p_last_error()
    assert lib.l4p_layernorm(None, None, None, None, None, 4, 8, ctypes.c_float(1e-5), 0, None) == -1


def test_ops_refuse_cpu_tensors():
    """The product path has no CPU fallback: CPU tensors raise instead of silently computing elsewhere."""
    import torch

    from l4p_b200 import lib as L, ops

    with pytest.raises(L.L4PError):
        ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8), 1e-5, out32=torch.zeros(4, 8))
    from l4p_b200.models.videomae import VideoMAEEncoder

    enc = VideoMAEEncoder(img_size=28, patch_size=14, embed_dim=16, depth=1, num_heads=2, all_frames=2, device="meta")
    with pytest.raises(L.L4PError):
        enc(torch.zeros(1, 3, 2, 28, 28))


def test_shape_validation_precedes_device_work(lib):
    """Every entry point checks its arguments before touching the device: bad shapes return L4P_ERR_SHAPE (-2) with a message
    naming the entry, null pointers L4P_ERR_ARG (-1). Dummy non-null pointers are never dereferenced on these paths."""
    P = ctypes.c_void_p(4096)
    f = ctypes.c_float
    cases = [
        # fused attention: N must be a multiple of 256, head_dim a multiple of 8 <= 96, pad 96
        (lambda: lib.l4p_attention(P, P, P, P, 1, 16, 2000, 88, 96, f(0.1), 0, None, None), -2, b"l4p_attention"),
        (lambda: lib.l4p_attention(P, P, P, P, 1, 16, 2048, 90, 96, f(0.1), 0, None, None), -2, b"head_dim"),
        (lambda: lib.l4p_attention(P, P, P, P, 1, 16, 2048, 88, 128, f(0.1), 0, None, None), -2, b"pad"),
        (lambda: lib.l4p_attention(None, P, P, P, 1, 16, 2048, 88, 96, f(0.1), 0, None, None), -1, b"null"),
        # LayerNorm kernels
        (lambda: lib.l4p_layernorm(P, P, P, P, None, 4, 1407, f(1e-6), 0, None), -2, b"l4p_layernorm"),
        (lambda: lib.l4p_layernorm16(P, P, P, P, 4, 4096, f(1e-6), 0, 0, None), -2, b"l4p_layernorm16"),
        # gathers
        (lambda: lib.l4p_patchify(P, P, 1, 3, 16, 224, 225, 2, 14, 14, 0, None), -2, b"l4p_patchify"),
        (lambda: lib.l4p_cast16(P, P, 6, 0, None), -2, b"multiple of 4"),
        (lambda: lib.l4p_upsample3d(P, P, None, 1, 4, 4, 4, 8, 8, 8, 12, 1, 0, None), -2, b"l4p_upsample3d"),
        # track-head attentions and read-out
        (lambda: lib.l4p_image_attention(P, P, P, P, 2, 256, 6, 8, 64, f(0.1), 0, None), -2, b"head_dim"),
        (lambda: lib.l4p_token_attention(P, P, P, P, 2, 6, 2048, 8, 90, 2048, f(0.1), 0, None), -2, b"l4p_token_attention"),
        (lambda: lib.l4p_track_readout(P, P, None, None, 2, 3, 16, 64, 64, 224, 224, None), -1, b"missing output"),
    ]
    for call, code, needle in cases:
        rc = call()
        msg = lib.l4p_last_error()
        assert rc == code, (rc, code, msg)
        assert needle in msg, (needle, msg)
