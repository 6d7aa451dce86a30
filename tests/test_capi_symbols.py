"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/astc_b200.h declares; compute entry points fail loudly without a GPU."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import ROOT, _has_cuda


def _declared():
    text = (ROOT / "include" / "astc_b200.h").read_text()
    return sorted(set(re.findall(r"ASTC_B200_API[^;]*?\b(astc_b200_\w+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("astc_b200_encode_device", "astc_b200_encode_host", "astc_b200_batch_encode",
                 "astc_b200_save_astc", "astc_b200_load_image", "astc_b200_band"):
        assert must in names


def test_every_declared_symbol_is_exported(native):
    raw = C.CDLL(str(ROOT / "astc_encoder_b200" / "libastc_b200.so"))
    missing = [n for n in _declared() if not hasattr(raw, n)]
    assert not missing, missing


def test_python_binding_covers_header(native):
    import astc_encoder_b200 as A
    assert sorted(A._SIGNATURES) == _declared()


def test_no_torch_types_in_header():
    text = (ROOT / "include" / "astc_b200.h").read_text()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)              # declarations only
    assert "torch" not in code and "at::" not in code and "cuda" not in code.replace("cuda_stream", "").replace("last_cuda_error", "")
    assert re.findall(r"#include\s*[<\"]([^>\"]+)", code) == ["stddef.h", "stdint.h"]


def test_option_struct_layout(native):
    import astc_encoder_b200 as A
    assert C.sizeof(A._Option) == 8
    o = A._Option()
    native.lib().astc_b200_option_default(C.byref(o))
    assert (o.is4x4, o.is6x6, o.is_normal_map, o.has_alpha, o.srgb) == (1, 0, 0, 0, 0)   # astc_encode.h:21-27


def test_strerror_and_version(native):
    L = native.lib()
    assert L.astc_b200_strerror(0) == b"ok"
    assert b"invalid" in L.astc_b200_strerror(-1)
    assert b"sm_100a" in L.astc_b200_version()


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(native):
    """No CPU fallback: on a box without a device the host entry point returns an error."""
    img = np.zeros((8, 8, 4), np.uint8)
    with pytest.raises(native.AstcError) as e:
        native.encode_astc_host(img, native.encode_option())
    assert e.value.status in (-2, -3)


def test_argument_validation(native):
    L = native.lib()
    o = native.encode_option()._abi()
    assert L.astc_b200_encode_device(None, 16, 16, 64, C.byref(o), None, None) == -1
    assert L.astc_b200_encode_device(None, -1, 16, 64, C.byref(o), None, None) == -1
    assert L.astc_b200_encode_device(None, 0, 0, 0, C.byref(o), None, None) == 0          # empty image: nothing to do
    assert L.astc_b200_encode_host(None, 16, 16, 64, C.byref(o), None) == -1
    assert L.astc_b200_encode_host(None, 0, 16, 0, C.byref(o), None) == 0
    assert L.astc_b200_decode_device(None, 4, 4, 5, None, 16, None) == -1
    assert L.astc_b200_mufu_device(2, None, None, 4, None) == -1                       # unknown op
    assert L.astc_b200_mufu_device(0, None, None, 4, None) == -1 and L.astc_b200_mufu_device(1, None, None, 0, None) == 0
    assert L.astc_b200_downsample2x2_device(None, 8, 8, 32, None, 16, None) == -1      # null pointers
    assert L.astc_b200_downsample2x2_device(None, -1, 8, 32, None, 16, None) == -1
    assert L.astc_b200_downsample2x2_device(None, 0, 8, 32, None, 16, None) == 0       # empty image: nothing to do
    assert L.astc_b200_bise_encode_device(None, 65, 0, 1, None, None) == -1


def test_mip_chain_layout(native):
    """astc_b200_mip_chain_layout (host only): levels down to 1x1, 256-byte aligned offsets, a scratch tail."""
    offs, ws, hs, total = native.mip_chain_layout(2048, 2048)
    assert ws == hs == [1024 >> i for i in range(11)]
    assert all(o % 256 == 0 for o in offs) and offs[0] == 0 and offs[1] == 1024 * 1024 * 4
    assert total == offs[-1] + 256 + 256
    assert native.mip_chain_layout(1, 1) == ([], [], [], 256)
    offs, ws, hs, _ = native.mip_chain_layout(5, 3)
    assert list(zip(ws, hs)) == [(2, 1), (1, 1)]
    offs, ws, hs, _ = native.mip_chain_layout(64, 128)
    assert list(zip(ws, hs)) == [(32, 64), (16, 32), (8, 16), (4, 8), (2, 4), (1, 2), (1, 1)]
    n = __import__("ctypes").c_int()
    assert native.lib().astc_b200_mip_chain_layout(0, 4, n, None, None, None, None) == -1
