"""astc_encoder_b200 -- B200-native ASTC block encoder (sm_100a).

Python host side over the C ABI in ``include/astc_b200.h`` (loaded with ctypes
from the in-tree ``libastc_b200.so``).  The names mirror the reference's
host interface so call sites read the same:

    encode_option      astc_encode.h:14-28
    encode_astc()      astc_encode.h:87     (device texture -> device block buffer, async)
    read_gpu()         astc_save.h:34       (download + sync)
    save_astc()        astc_save.h:52       (16-byte header + blocks)
    load_tex()         main.cpp:19          (decode + vertical flip + RGBA8 + upload)

PyTorch is used only as plumbing (device memory, streams, torch.distributed).
There is no CPU encode path: if the CUDA library is missing or no GPU is
visible, the compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

__all__ = [
    "encode_option", "AstcError", "lib", "block_dim", "block_counts", "output_size", "band",
    "encode_astc", "encode_astc_host", "read_gpu", "save_astc", "save_astc_slice", "load_astc", "load_image", "load_tex",
    "decode_astc", "downsample2x2", "mip_chain", "mip_chain_layout", "mip_chain_by_level", "mufu", "bise_encode", "Batch", "Context", "launch_count", "unorm_lut", "version",
]

_PKG = Path(__file__).resolve().parent
import os as _os
# ASTC_B200_LIB selects an experiment build (tools/variants.py); the product is libastc_b200.so
_LIB_PATH = Path(_os.environ["ASTC_B200_LIB"]) if _os.environ.get("ASTC_B200_LIB") else _PKG / "libastc_b200.so"

BLOCK_BYTES = 16


class AstcError(RuntimeError):
    """A C-ABI call returned a negative astc_b200_status."""

    def __init__(self, status: int, where: str):
        l = lib()
        detail = l.astc_b200_last_cuda_error().decode() if status in (-2, -3, -4) else ""
        super().__init__(f"{where}: {l.astc_b200_strerror(status).decode()} ({status}) {detail}".strip())
        self.status = status


class _Option(C.Structure):
    _fields_ = [("is4x4", C.c_uint8), ("is6x6", C.c_uint8), ("is_normal_map", C.c_uint8),
                ("has_alpha", C.c_uint8), ("srgb", C.c_uint8), ("axis_method", C.c_uint8), ("reserved", C.c_uint8 * 2)]


class _HostImage(C.Structure):
    _fields_ = [("h_rgba", C.c_void_p), ("h_blocks", C.c_void_p), ("pitch_bytes", C.c_size_t),
                ("width", C.c_int32), ("height", C.c_int32)]


class _Image(C.Structure):
    _fields_ = [("d_rgba", C.c_void_p), ("d_blocks", C.c_void_p), ("pitch_bytes", C.c_size_t),
                ("width", C.c_int32), ("height", C.c_int32)]


@dataclass
class encode_option:
    """Same fields, order and defaults as the reference struct (astc_encode.h:14-28), plus one
    extension: axis_method = 1 selects max_accumulation_pixel_direction (ASTC_Encode.hlsl:170-227,
    the alternative the reference carries commented out at :514) instead of the PCA (0, default)."""
    is4x4: bool = True
    is6x6: bool = False
    is_normal_map: bool = False
    has_alpha: bool = False
    srgb: bool = False
    axis_method: int = 0

    def _abi(self) -> _Option:
        return _Option(int(self.is4x4), int(self.is6x6), int(self.is_normal_map),
                       int(self.has_alpha), int(self.srgb), int(self.axis_method))

    @classmethod
    def from_args(cls, args: Sequence[str]) -> "encode_option":
        """parse_cmd (main.cpp:140-178): exact flag matches, unknown flags ignored."""
        o = cls()
        for a in args:
            if a == "-4x4":
                o.is4x4 = True
            elif a == "-6x6":
                o.is6x6 = True
            elif a == "-norm":
                o.is_normal_map = True
            elif a == "-srgb":
                o.srgb = True
            elif a == "-alpha":
                o.has_alpha = True
            elif a == "-accum":                              # extension; the reference ignores unknown flags
                o.axis_method = 1
        return o


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); also the list tests check against include/astc_b200.h
_SIGNATURES = {
    "astc_b200_version": (C.c_char_p, []),
    "astc_b200_strerror": (C.c_char_p, [C.c_int]),
    "astc_b200_last_cuda_error": (C.c_char_p, []),
    "astc_b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "astc_b200_set_device": (C.c_int, [C.c_int]),
    "astc_b200_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "astc_b200_option_default": (None, [C.POINTER(_Option)]),
    "astc_b200_block_dim": (C.c_int, [C.POINTER(_Option)]),
    "astc_b200_block_counts": (C.c_int, [C.c_int, C.c_int, C.POINTER(_Option), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "astc_b200_output_size": (C.c_size_t, [C.c_int, C.c_int, C.POINTER(_Option)]),
    "astc_b200_band": (C.c_int, [C.c_int, C.c_int, C.POINTER(_Option), C.c_int, C.c_int, C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "astc_b200_encode_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(_Option), C.c_void_p, C.c_void_p]),
    "astc_b200_encode_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(_Option), C.c_void_p]),
    "astc_b200_context_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "astc_b200_context_destroy": (None, [C.c_void_p]),
    "astc_b200_context_trim": (C.c_int, [C.c_void_p]),
    "astc_b200_context_batch_encode_mip_chains_host": (C.c_int, [C.c_void_p, C.POINTER(_HostImage), C.c_int, C.POINTER(_Option)]),
    "astc_b200_mip_chain_output_size": (C.c_int, [C.c_int, C.c_int, C.POINTER(_Option), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "astc_b200_mip_chain_layout": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_int),
                                             C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "astc_b200_mip_chain_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "astc_b200_context_set_copy_threads": (C.c_int, [C.c_void_p, C.c_int]),
    "astc_b200_context_encode_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(_Option), C.c_void_p]),
    "astc_b200_context_batch_encode_host": (C.c_int, [C.c_void_p, C.POINTER(_HostImage), C.c_int, C.POINTER(_Option)]),
    "astc_b200_batch_create": (C.c_int, [C.POINTER(_Image), C.c_int, C.POINTER(_Option), C.POINTER(C.c_void_p)]),
    "astc_b200_batch_encode": (C.c_int, [C.c_void_p, C.c_void_p]),
    "astc_b200_batch_total_blocks": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "astc_b200_batch_destroy": (None, [C.c_void_p]),
    "astc_b200_launch_count": (C.c_uint64, []),
    "astc_b200_bise_encode_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "astc_b200_quant_layout": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "astc_b200_ise_bitcount": (C.c_uint32, [C.c_uint32, C.c_int]),
    "astc_b200_integer_from_trits": (C.c_int, [C.c_int] * 5),
    "astc_b200_integer_from_quints": (C.c_int, [C.c_int] * 3),
    "astc_b200_scramble": (C.c_int, [C.c_int, C.c_int]),
    "astc_b200_blockmode": (C.c_uint32, [C.c_int]),
    "astc_b200_unorm_lut": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    "astc_b200_decode_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "astc_b200_mufu_device": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "astc_b200_downsample2x2_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "astc_b200_malloc_device": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "astc_b200_free_device": (C.c_int, [C.c_void_p]),
    "astc_b200_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "astc_b200_host_free": (C.c_int, [C.c_void_p]),
    "astc_b200_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "astc_b200_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "astc_b200_memcpy2d_h2d": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "astc_b200_stream_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "astc_b200_stream_destroy": (C.c_int, [C.c_void_p]),
    "astc_b200_stream_synchronize": (C.c_int, [C.c_void_p]),
    "astc_b200_save_astc": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "astc_b200_save_astc_slice": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "astc_b200_load_astc": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "astc_b200_load_image": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.POINTER(C.c_void_p)]),
    "astc_b200_image_failure_reason": (C.c_char_p, []),
    "astc_b200_free_host_buffer": (None, [C.c_void_p]),
}


def lib() -> C.CDLL:
    """The native library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise ImportError(
                f"{_LIB_PATH} is missing: build it with `python -m astc_encoder_b200.build` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(status: int, where: str) -> None:
    if status != 0:
        raise AstcError(status, where)


def version() -> str:
    return lib().astc_b200_version().decode()


def launch_count() -> int:
    return int(lib().astc_b200_launch_count())


# ---------------------------------------------------------------- geometry --
def block_dim(option: encode_option) -> int:
    o = option._abi()
    return int(lib().astc_b200_block_dim(C.byref(o)))


def block_counts(width: int, height: int, option: encode_option) -> tuple[int, int]:
    o = option._abi()
    bx, by = C.c_int(), C.c_int()
    _check(lib().astc_b200_block_counts(width, height, C.byref(o), C.byref(bx), C.byref(by)), "block_counts")
    return bx.value, by.value


def output_size(width: int, height: int, option: encode_option) -> int:
    o = option._abi()
    return int(lib().astc_b200_output_size(width, height, C.byref(o)))


def band(width: int, height: int, option: encode_option, parts: int, part: int) -> tuple[int, int, int, int]:
    """(y0, rows, block_byte_offset, block_bytes) of band `part` of `parts`."""
    o = option._abi()
    y0, rows, off, nbytes = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
    _check(lib().astc_b200_band(width, height, C.byref(o), parts, part, C.byref(y0), C.byref(rows),
                                C.byref(off), C.byref(nbytes)), "band")
    return y0.value, rows.value, off.value, nbytes.value


# ------------------------------------------------------------------ encode --
def _stream_ptr(stream) -> int:
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return int(stream.cuda_stream)


def _effective(option: encode_option, srgb_texture: Optional[bool]) -> _Option:
    o = option._abi()
    if srgb_texture is not None:
        o.srgb = int(bool(srgb_texture))
    return o


def _check_src(src):
    """The source-tensor contract of every device entry point: CUDA uint8 (H, W, 4), texels packed,
    rows may be strided.  Returns (h, w, pitch_bytes)."""
    import torch
    if not (isinstance(src, torch.Tensor) and src.is_cuda and src.dtype == torch.uint8 and src.dim() == 3
            and src.shape[2] == 4 and src.stride(2) == 1 and src.stride(1) == 4):
        raise ValueError("src must be a CUDA uint8 tensor of shape (H, W, 4) with packed texels")
    h, w = int(src.shape[0]), int(src.shape[1])
    return h, w, (int(src.stride(0)) if h > 1 else w * 4)


def _check_out(out, nbytes: int, device) -> None:
    import torch
    if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous()
            and out.numel() >= nbytes):
        raise ValueError("out must be a contiguous CUDA uint8 tensor of at least output_size bytes")
    if out.device != device:
        raise ValueError(f"out is on {out.device}, the source on {device}")


def encode_astc(src, option: encode_option, out=None, stream=None, srgb_texture: Optional[bool] = None):
    """encode_astc (astc_encode.h:87): `src` is a CUDA uint8 tensor (H, W, 4), rows may be
    strided; returns a CUDA uint8 tensor (blocks, 16).  Asynchronous on `stream`
    (default: torch's current stream), like the reference's Dispatch.
    `srgb_texture` overrides option.srgb the way the texture format does (main.cpp:214)."""
    import torch
    h, w, pitch = _check_src(src)
    o = _effective(option, srgb_texture)
    nbytes = int(lib().astc_b200_output_size(w, h, C.byref(o)))
    if out is None:
        out = torch.empty((nbytes // BLOCK_BYTES, BLOCK_BYTES), dtype=torch.uint8, device=src.device)
    else:
        _check_out(out, nbytes, src.device)
    with torch.cuda.device(src.device):
        _check(lib().astc_b200_encode_device(src.data_ptr(), w, h, pitch, C.byref(o), out.data_ptr(),
                                             _stream_ptr(stream)), "encode_astc")
    return out


def encode_astc_host(rgba: np.ndarray, option: encode_option, out: Optional[np.ndarray] = None,
                     srgb_texture: Optional[bool] = None) -> np.ndarray:
    """Upload + encode + read-back in one synchronous call on host memory
    (load_tex's upload + encode_astc + read_gpu).  rgba: (H, W, 4) uint8."""
    rgba, o, out, pitch = _host_args(rgba, option, out, srgb_texture)
    h, w = rgba.shape[:2]
    _check(lib().astc_b200_encode_host(rgba.ctypes.data, w, h, pitch, C.byref(o), out.ctypes.data), "encode_astc_host")
    return out


def _host_args(rgba, option, out, srgb_texture):
    if rgba.dtype != np.uint8 or rgba.ndim != 3 or rgba.shape[2] != 4 or \
            (rgba.size and (rgba.strides[2] != 1 or rgba.strides[1] != 4)):
        raise ValueError("rgba must be a uint8 array of shape (H, W, 4) with packed texels")
    h, w = rgba.shape[:2]
    o = _effective(option, srgb_texture)
    nbytes = int(lib().astc_b200_output_size(w, h, C.byref(o)))
    if out is None:
        out = np.empty((nbytes // BLOCK_BYTES, BLOCK_BYTES), dtype=np.uint8)
    elif out.dtype != np.uint8 or not out.flags.c_contiguous or out.size < nbytes:
        raise ValueError("out must be a contiguous uint8 array of at least output_size bytes")
    pitch = rgba.strides[0] if (h > 1 and rgba.size) else w * 4
    return rgba, o, out, pitch


class Context:
    """Persistent host-side context (astc_b200_context_*): streams, device workspace and pinned staging
    kept across calls.  One per (host thread, device); create it with the target device current."""

    def __init__(self):
        self._h = C.c_void_p()
        _check(lib().astc_b200_context_create(C.byref(self._h)), "Context")

    def encode_host(self, rgba: np.ndarray, option: encode_option, out: Optional[np.ndarray] = None,
                    srgb_texture: Optional[bool] = None) -> np.ndarray:
        rgba, o, out, pitch = _host_args(rgba, option, out, srgb_texture)
        h, w = rgba.shape[:2]
        _check(lib().astc_b200_context_encode_host(self._h, rgba.ctypes.data, w, h, pitch, C.byref(o), out.ctypes.data),
               "Context.encode_host")
        return out

    def batch_encode_host(self, images: Sequence[np.ndarray], option: encode_option,
                          outs: Optional[Sequence[np.ndarray]] = None) -> list:
        """Many (H, W, 4) uint8 host images -> list of (blocks, 16) uint8 arrays, one synchronous call."""
        o = option._abi()
        imgs, res = (_HostImage * max(1, len(images)))(), []
        for i, im in enumerate(images):
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 4 or (im.size and (im.strides[2] != 1 or im.strides[1] != 4)):
                raise ValueError("images must be uint8 arrays of shape (H, W, 4) with packed texels")
            h, w = im.shape[:2]
            n = int(lib().astc_b200_output_size(w, h, C.byref(o)))
            dst = outs[i] if outs is not None else np.empty((n // BLOCK_BYTES, BLOCK_BYTES), dtype=np.uint8)
            if dst.dtype != np.uint8 or not dst.flags.c_contiguous or dst.size < n:
                raise ValueError("outs[i] must be a contiguous uint8 array of at least output_size bytes")
            res.append(dst)
            imgs[i] = _HostImage(im.ctypes.data, dst.ctypes.data, im.strides[0] if (h > 1 and im.size) else w * 4, w, h)
        _check(lib().astc_b200_context_batch_encode_host(self._h, imgs, len(images), C.byref(o)), "Context.batch_encode_host")
        return res

    def batch_encode_mip_chains_host(self, bases: Sequence[np.ndarray], option: encode_option) -> list:
        """Whole mip chains from their BASE levels only (astc_b200_context_batch_encode_mip_chains_host): the bases are
        uploaded, the levels below are generated and encoded on the device.  Returns, per base, the list of
        (blocks, 16) uint8 arrays of its levels (base first; views into one buffer per chain)."""
        o = option._abi()
        imgs, res = (_HostImage * max(1, len(bases)))(), []
        d = block_dim(option)
        for i, im in enumerate(bases):
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 4 or (im.size and (im.strides[2] != 1 or im.strides[1] != 4)):
                raise ValueError("bases must be uint8 arrays of shape (H, W, 4) with packed texels")
            h, w = im.shape[:2]
            nbytes, levels = C.c_size_t(), C.c_int()
            _check(lib().astc_b200_mip_chain_output_size(w, h, C.byref(o), C.byref(nbytes), C.byref(levels)), "mip_chain_output_size")
            buf = np.empty(nbytes.value, dtype=np.uint8)
            imgs[i] = _HostImage(im.ctypes.data, buf.ctypes.data, im.strides[0] if (h > 1 and im.size) else w * 4, w, h)
            views, off, lw, lh = [], 0, w, h
            for _ in range(levels.value):
                n = ((lw + d - 1) // d) * ((lh + d - 1) // d) * BLOCK_BYTES
                views.append(buf[off:off + n].reshape(-1, BLOCK_BYTES))
                off += n
                lw, lh = max(1, lw // 2), max(1, lh // 2)
            res.append(views)
        _check(lib().astc_b200_context_batch_encode_mip_chains_host(self._h, imgs, len(bases), C.byref(o)), "Context.batch_encode_mip_chains_host")
        return res

    def trim(self) -> None:
        _check(lib().astc_b200_context_trim(self._h), "Context.trim")

    def set_copy_threads(self, threads: int) -> None:
        """Worker threads of the staged (pageable-memory) pipeline besides the caller: -1 automatic, 0 none."""
        _check(lib().astc_b200_context_set_copy_threads(self._h, int(threads)), "Context.set_copy_threads")

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().astc_b200_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_gpu(buffer, stream=None) -> np.ndarray:
    """read_gpu (astc_save.h:34-50): download the block buffer and synchronise."""
    import torch
    if not (isinstance(buffer, torch.Tensor) and buffer.is_cuda and buffer.is_contiguous()):
        raise ValueError("buffer must be a contiguous CUDA tensor")
    host = torch.empty(buffer.shape, dtype=buffer.dtype, pin_memory=True)
    with torch.cuda.device(buffer.device):
        s = _stream_ptr(stream)
        _check(lib().astc_b200_memcpy_d2h(host.data_ptr(), buffer.data_ptr(), buffer.numel() * buffer.element_size(), s), "read_gpu")
        _check(lib().astc_b200_stream_synchronize(s), "read_gpu")
    return host.numpy().copy()


class Batch:
    """Many textures (e.g. all mips of many chains) encoded by ONE kernel launch
    over a prefix-summed block table (astc_b200_batch_*).  Keeps the tensors alive."""

    def __init__(self, sources: Sequence, option: encode_option, outputs: Optional[Sequence] = None):
        import torch
        self.option = option
        self.sources = list(sources)
        self._handle = None
        o = option._abi()
        # the same tensor contract as encode_astc for every source and output; one device for the whole batch
        # (the descriptor table and the launch live on it -- astc_b200_batch_encode checks the ordinal again)
        geom = [_check_src(s) for s in self.sources]
        self.device = self.sources[0].device if self.sources else torch.device("cuda", torch.cuda.current_device())
        for s in self.sources:
            if s.device != self.device:
                raise ValueError(f"mixed-device batch: {s.device} and {self.device}")
        sizes = [output_size(w, h, option) for h, w, _ in geom]
        if outputs is None:
            outputs = [torch.empty((n // BLOCK_BYTES, BLOCK_BYTES), dtype=torch.uint8, device=self.device) for n in sizes]
        self.outputs = list(outputs)
        if len(self.outputs) != len(self.sources):
            raise ValueError("one output per source")
        for d, n in zip(self.outputs, sizes):
            _check_out(d, n, self.device)
        imgs = (_Image * max(1, len(self.sources)))()
        for i, (s, d, (h, w, pitch)) in enumerate(zip(self.sources, self.outputs, geom)):
            imgs[i] = _Image(s.data_ptr(), d.data_ptr(), pitch, w, h)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(lib().astc_b200_batch_create(imgs, len(self.sources), C.byref(o), C.byref(handle)), "Batch")
        self._handle = handle
        nb, nt = C.c_uint64(), C.c_uint64()
        lib().astc_b200_batch_total_blocks(self._handle, C.byref(nb), C.byref(nt))
        self.total_blocks, self.total_texels = nb.value, nt.value

    def encode(self, stream=None):
        import torch
        with torch.cuda.device(self.device):
            _check(lib().astc_b200_batch_encode(self._handle, _stream_ptr(stream)), "Batch.encode")
        return self.outputs

    def close(self):
        if getattr(self, "_handle", None):
            lib().astc_b200_batch_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode_astc(blocks, width: int, height: int, dim: int, stream=None):
    """Device decode of the subset this encoder emits -> CUDA uint8 (H, W, 4)."""
    import torch
    if not (isinstance(blocks, torch.Tensor) and blocks.is_cuda and blocks.dtype == torch.uint8 and blocks.is_contiguous()):
        raise ValueError("blocks must be a contiguous CUDA uint8 tensor")
    if dim not in (4, 6):
        raise ValueError("dim must be 4 or 6")
    if blocks.numel() < ((width + dim - 1) // dim) * ((height + dim - 1) // dim) * BLOCK_BYTES:
        raise ValueError("blocks is smaller than the block grid of a width x height texture")
    out = torch.empty((height, width, 4), dtype=torch.uint8, device=blocks.device)
    with torch.cuda.device(blocks.device):
        _check(lib().astc_b200_decode_device(blocks.data_ptr(), width, height, dim, out.data_ptr(), width * 4,
                                             _stream_ptr(stream)), "decode_astc")
    return out


def downsample2x2(img, out=None, stream=None):
    """Next mip level of a CUDA uint8 (H, W, 4) image: 2x2 box filter, round half up -> (max(1,H//2), max(1,W//2), 4)."""
    import torch
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 4 or not img.is_cuda or img.stride(2) != 1 or img.stride(1) != 4:
        raise ValueError("downsample2x2 expects a CUDA uint8 (H, W, 4) tensor with contiguous texels")
    h, w = int(img.shape[0]), int(img.shape[1])
    oh, ow = max(1, h // 2), max(1, w // 2)
    if out is None:
        out = torch.empty((oh, ow, 4), dtype=torch.uint8, device=img.device)
    _check(lib().astc_b200_downsample2x2_device(img.data_ptr(), w, h, int(img.stride(0)), out.data_ptr(), int(out.stride(0)),
                                                _stream_ptr(stream)), "downsample2x2")
    return out


def mip_chain_layout(width: int, height: int):
    """(offsets, widths, heights, total_bytes) of the levels below a width x height base in one arena
    (astc_b200_mip_chain_layout)."""
    n, total = C.c_int(), C.c_size_t()
    offs, ws, hs = (C.c_size_t * 24)(), (C.c_int * 24)(), (C.c_int * 24)()
    _check(lib().astc_b200_mip_chain_layout(int(width), int(height), C.byref(n), offs, ws, hs, C.byref(total)), "mip_chain_layout")
    return list(offs[:n.value]), list(ws[:n.value]), list(hs[:n.value]), int(total.value)


def mip_chain(base, stream=None, arena=None):
    """[base, level 1, ..., 1x1] generated on the device by ONE call (astc_b200_mip_chain_device: one fused launch
    when both sides are multiples of 64, else one launch per level).  The levels are views into one arena tensor."""
    import torch
    h, w, pitch = _check_src(base)
    offs, ws, hs, total = mip_chain_layout(w, h)
    if not offs:
        return [base]
    if arena is None:
        arena = torch.empty(total, dtype=torch.uint8, device=base.device)
    elif not (arena.is_cuda and arena.dtype == torch.uint8 and arena.is_contiguous() and arena.numel() >= total and arena.device == base.device):
        raise ValueError("arena must be a contiguous CUDA uint8 tensor of at least mip_chain_layout's total_bytes on the base's device")
    with torch.cuda.device(base.device):
        _check(lib().astc_b200_mip_chain_device(base.data_ptr(), w, h, pitch, arena.data_ptr(), arena.numel(), _stream_ptr(stream)), "mip_chain")
    return [base] + [arena[o:o + lw * lh * 4].view(lh, lw, 4) for o, lw, lh in zip(offs, ws, hs)]


def mip_chain_by_level(base, stream=None):
    """The same chain, one astc_b200_downsample2x2_device launch per level (the checker of mip_chain)."""
    chain = [base]
    while chain[-1].shape[0] > 1 or chain[-1].shape[1] > 1:
        chain.append(downsample2x2(chain[-1], stream=stream))
    return chain


def mufu(op: str, x, stream=None):
    """rcp.approx.ftz / rsqrt.approx.ftz of a CUDA float32 tensor (op = "rcp" | "rsq")."""
    import torch
    x = x.contiguous()
    y = torch.empty_like(x)
    _check(lib().astc_b200_mufu_device({"rcp": 0, "rsq": 1}[op], x.data_ptr(), y.data_ptr(), x.numel(), _stream_ptr(stream)), "mufu")
    return y


def bise_encode(values, quant: int, stream=None):
    """values: CUDA uint8 (nseq, count).  Returns CUDA uint8 (nseq, 16) ISE streams."""
    import torch
    values = values.contiguous()
    nseq, count = int(values.shape[0]), int(values.shape[1])
    out = torch.zeros((nseq, 16), dtype=torch.uint8, device=values.device)
    _check(lib().astc_b200_bise_encode_device(values.data_ptr(), count, quant, nseq, out.data_ptr(),
                                              _stream_ptr(stream)), "bise_encode")
    return out


def unorm_lut(srgb: bool) -> np.ndarray:
    out = (C.c_float * 256)()
    _check(lib().astc_b200_unorm_lut(int(bool(srgb)), out), "unorm_lut")
    return np.frombuffer(bytes(out), dtype=np.float32).copy()


# --------------------------------------------------------------- host files --
def save_astc(astc_path: str, xdim: int, ydim: int, xsize: int, ysize: int, buffer) -> None:
    """save_astc (astc_save.h:52-76), byte-exact header + raw blocks."""
    buf = np.ascontiguousarray(np.asarray(buffer, dtype=np.uint8))
    _check(lib().astc_b200_save_astc(str(astc_path).encode(), xdim, ydim, xsize, ysize, buf.ctypes.data, buf.size),
           "save_astc")


def save_astc_slice(astc_path: str, xdim: int, ydim: int, xsize: int, ysize: int, byte_offset: int, buffer,
                    write_header: bool) -> None:
    """One rank's slice of a shared .astc file (astc_b200_save_astc_slice): safe in any order, never truncates."""
    buf = np.ascontiguousarray(np.asarray(buffer, dtype=np.uint8))
    _check(lib().astc_b200_save_astc_slice(str(astc_path).encode(), xdim, ydim, xsize, ysize, byte_offset,
                                           buf.ctypes.data, buf.size, int(bool(write_header))), "save_astc_slice")


def load_astc(astc_path: str) -> tuple[int, int, int, int, np.ndarray]:
    """Returns (xdim, ydim, xsize, ysize, blocks (n, 16) uint8)."""
    xd, yd, xs, ys = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    p, n = C.c_void_p(), C.c_size_t()
    _check(lib().astc_b200_load_astc(str(astc_path).encode(), C.byref(xd), C.byref(yd), C.byref(xs), C.byref(ys),
                                     C.byref(p), C.byref(n)), "load_astc")
    try:
        blocks = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value,)).copy() if n.value else \
            np.zeros(0, np.uint8)
    finally:
        lib().astc_b200_free_host_buffer(p)
    return xd.value, yd.value, xs.value, ys.value, blocks.reshape(-1, 16)


def load_image(path: str, flip_vertically: bool = True, with_components: bool = False):
    """stbi_load(path, ..., STBI_rgb_alpha) with the reference's vertical flip
    (main.cpp:24-25).  Returns (H, W, 4) uint8; with_components=True also returns the channel
    count the file held (stbi_load's `comp`)."""
    w, h, comp = C.c_int(), C.c_int(), C.c_int()
    p = C.c_void_p()
    rc = lib().astc_b200_load_image(str(path).encode(), int(flip_vertically), C.byref(w), C.byref(h), C.byref(comp),
                                    C.byref(p))
    if rc != 0:
        raise AstcError(rc, f"load_image({path}): {lib().astc_b200_image_failure_reason().decode()}")
    try:
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(h.value, w.value, 4)).copy()
    finally:
        lib().astc_b200_free_host_buffer(p)
    return (arr, comp.value) if with_components else arr


def load_tex(tex_path: str, device="cuda", flip_vertically: bool = True):
    """load_tex (main.cpp:19-56): decode, flip, force RGBA8, upload -> CUDA tensor."""
    import torch
    return torch.from_numpy(load_image(tex_path, flip_vertically)).to(device)
