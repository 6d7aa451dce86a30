"""Host side of the drop-in surface: .astc container (astc_save.h:3-14,52-76), option
parsing (main.cpp:140-178), image ingest (main.cpp:19-30) and the astc_cs_enc CLI."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, _has_cuda

CLI = ROOT / "astc_encoder_b200" / "bin" / "astc_cs_enc"


def test_save_astc_header_is_byte_exact(native, tmp_path):
    blocks = np.arange(6 * 16, dtype=np.uint8).reshape(6, 16)
    p = tmp_path / "t.astc"
    native.save_astc(str(p), 6, 6, 13, 9, blocks)              # 3 x 2 blocks of 6x6
    raw = p.read_bytes()
    assert raw[:4] == bytes([0x13, 0xAB, 0xA1, 0x5C])          # MAGIC_FILE_CONSTANT little-endian
    assert raw[4:7] == bytes([6, 6, 1])
    assert raw[7:10] == bytes([13, 0, 0]) and raw[10:13] == bytes([9, 0, 0]) and raw[13:16] == bytes([1, 0, 0])
    assert raw[16:] == blocks.tobytes()
    xd, yd, xs, ys, got = native.load_astc(str(p))
    assert (xd, yd, xs, ys) == (6, 6, 13, 9) and np.array_equal(got, blocks)


def test_save_astc_24bit_sizes(native, tmp_path):
    p = tmp_path / "big.astc"
    native.save_astc(str(p), 4, 4, 0x012345, 0x00ABCD, np.zeros((0, 16), np.uint8))
    raw = p.read_bytes()
    assert raw[7:10] == bytes([0x45, 0x23, 0x01]) and raw[10:13] == bytes([0xCD, 0xAB, 0x00])


def test_golden_roundtrips_through_writer(native, leaf_golden, tmp_path):
    xd, yd, xs, ys, blocks = leaf_golden
    p = tmp_path / "leaf.astc"
    native.save_astc(str(p), xd, yd, xs, ys, blocks)
    assert p.read_bytes() == (GOLDEN / "leaf.astc").read_bytes()


def test_load_astc_rejects_garbage(native, tmp_path):
    p = tmp_path / "bad.astc"
    p.write_bytes(b"not an astc file at all")
    with pytest.raises(native.AstcError):
        native.load_astc(str(p))
    with pytest.raises(native.AstcError):
        native.load_astc(str(tmp_path / "missing.astc"))
    q = tmp_path / "short.astc"
    q.write_bytes((GOLDEN / "leaf.astc").read_bytes()[:1000])
    with pytest.raises(native.AstcError):
        native.load_astc(str(q))


def test_option_parsing_matches_parse_cmd(native):
    E = native.encode_option
    assert E.from_args([]) == E()                               # defaults: 4x4 only (astc_encode.h:21-27)
    o = E.from_args(["-alpha", "-srgb", "-bogus", "--alpha", "-norm"])
    assert (o.is4x4, o.is6x6, o.has_alpha, o.srgb, o.is_normal_map) == (True, False, True, True, True)
    assert native.block_dim(E.from_args(["-6x6"])) == 6         # deviation: -6x6 is live here
    assert native.block_dim(E.from_args(["-4x4"])) == 4
    assert native.block_dim(E(is4x4=False)) == 6


def _png(w, h, color_type, depth, rows, palette=None, trns=None, interlace=0):
    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    raw = b"".join(b"\x00" + r for r in rows)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, interlace))
    if palette is not None:
        out += chunk(b"PLTE", palette)
    if trns is not None:
        out += chunk(b"tRNS", trns)
    return out + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")


def test_png_variants_decode_like_pil(native, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(5)
    cases = {}
    rgb = rng.integers(0, 256, (7, 5, 3), dtype=np.uint8)
    cases["rgb8"] = _png(5, 7, 2, 8, [bytes(r.tobytes()) for r in rgb])
    g = rng.integers(0, 256, (4, 9), dtype=np.uint8)
    cases["gray8"] = _png(9, 4, 0, 8, [bytes(r.tobytes()) for r in g])
    ga = rng.integers(0, 256, (3, 6, 2), dtype=np.uint8)
    cases["graya8"] = _png(6, 3, 4, 8, [bytes(r.tobytes()) for r in ga])
    pal = rng.integers(0, 256, (4, 3), dtype=np.uint8)
    idx = rng.integers(0, 4, (5, 8), dtype=np.uint8)
    packed = [bytes(np.packbits(np.unpackbits(r[:, None], axis=1)[:, 6:].reshape(-1)).tobytes()) for r in idx]
    cases["pal2"] = _png(8, 5, 3, 2, packed, palette=pal.tobytes(), trns=bytes([255, 0, 128]))
    rgb16 = rng.integers(0, 65536, (2, 3, 3), dtype=np.uint16)
    cases["rgb16"] = _png(3, 2, 2, 16, [bytes(r.astype(">u2").tobytes()) for r in rgb16])
    for name, data in cases.items():
        p = tmp_path / f"{name}.png"
        p.write_bytes(data)
        want = np.asarray(Image.open(p).convert("RGBA"))
        got = native.load_image(str(p), flip_vertically=False)
        if name == "rgb16":
            want = np.concatenate([(rgb16 >> 8).astype(np.uint8), np.full((2, 3, 1), 255, np.uint8)], axis=2)  # stb keeps the high byte
        assert np.array_equal(got, want), name
        assert np.array_equal(native.load_image(str(p), flip_vertically=True), want[::-1]), name


def test_png_saved_by_pil_and_other_formats(native, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (33, 17, 4), dtype=np.uint8)
    for ext, mode in (("png", "RGBA"), ("bmp", "RGB"), ("tga", "RGBA"), ("ppm", "RGB")):
        p = tmp_path / f"x.{ext}"
        Image.fromarray(img, "RGBA").convert(mode).save(p)
        want = np.asarray(Image.open(p).convert("RGBA"))
        assert np.array_equal(native.load_image(str(p), flip_vertically=False), want), ext


def test_load_image_failure_reason(native, tmp_path):
    with pytest.raises(native.AstcError) as e:
        native.load_image(str(tmp_path / "nope.png"))
    assert "fopen" in str(e.value)
    p = tmp_path / "junk.png"
    p.write_bytes(b"\x89PNG\r\n\x1a\n" + b"\x00" * 40)
    with pytest.raises(native.AstcError):
        native.load_image(str(p))


def test_cli_exists_and_arg_errors(native):
    assert CLI.exists()
    r = subprocess.run([str(CLI)], capture_output=True, text=True)
    assert r.returncode != 0 and "wrong args count" in r.stdout           # main.cpp:182-185


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_cli_without_gpu_fails_loudly(native, tmp_path):
    r = subprocess.run([str(CLI), str(GOLDEN / "leaf.png"), "-alpha"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "encode option setting:" in r.stdout and "has_alpha\ttrue" in r.stdout   # main.cpp:193-197
    assert "init cuda failed" in r.stdout
    assert not (GOLDEN / "leaf_out.astc").exists()


@pytest.mark.gpu
def test_cli_reproduces_readme_example(native, oracle, tmp_path, leaf_golden):
    """README.md:38-40 `astc_cs_enc textures/leaf.png -alpha -4x4` (the golden has no -srgb)."""
    src = tmp_path / "leaf.v2.png"
    src.write_bytes((GOLDEN / "leaf.png").read_bytes())
    r = subprocess.run([str(CLI), str(src), "-alpha", "-4x4", "-unknown"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = tmp_path / "leaf.v2.astc"                                       # only the LAST extension is stripped
    assert f"save astc to:{out}" in r.stdout
    xd, yd, xs, ys, blocks = native.load_astc(str(out))
    assert (xd, yd, xs, ys) == (4, 4, 1024, 1024)
    want = oracle.encode_image(native.load_image(str(src), True), block_dim=4, has_alpha=True)
    assert np.array_equal(blocks, want)
    assert (blocks == leaf_golden[4]).all(axis=1).mean() >= 0.999
    # -6x6 -srgb on the same file: header says 6x6, blocks match the oracle
    r = subprocess.run([str(CLI), str(src), "-6x6", "-srgb"], capture_output=True, text=True)
    assert r.returncode == 0 and "is 4x4 block\tfalse" in r.stdout
    xd, yd, xs, ys, blocks = native.load_astc(str(out))
    assert (xd, yd) == (6, 6) and len(blocks) == 171 * 171
    want = oracle.encode_image(native.load_image(str(src), True), block_dim=6, srgb=True)
    assert np.array_equal(blocks, want)


def test_copy_pool_unit(tmp_path):
    """The worker pool behind the staged (pageable-memory) pipeline of astc_b200_context_encode_host, compiled as plain
    C++ and run on the CPU: pitched / contiguous / tiny / multi-grain jobs against memcpy row by row."""
    exe = tmp_path / "copy_pool_test"
    subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-I", str(ROOT / "astc_encoder_b200" / "csrc"),
                    str(ROOT / "tests" / "cpp" / "copy_pool_test.cpp"), "-o", str(exe)], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "copy_pool_test: ok" in res.stdout, res.stdout + res.stderr


def test_cta_schedule_unit(tmp_path):
    """The launch schedule (csrc/astc_schedule.h): uniform and tapered plans cover every block id exactly once."""
    cuda_inc = "/usr/local/cuda/include"
    exe = tmp_path / "schedule_test"
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", str(ROOT / "astc_encoder_b200" / "csrc"), "-I", cuda_inc,
                    str(ROOT / "tests" / "cpp" / "schedule_test.cpp"), "-o", str(exe)], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "schedule_test: ok" in res.stdout, res.stdout + res.stderr
