#!/usr/bin/env python3
"""Input-format fixtures for tests/test_image_formats.py.

Writes a small corpus of image files in every format the reference's loader reads (stb_image v2.22,
main.cpp:24-25) into tests/golden/images/ and, next to it, expected.npz: the RGBA8 texels
stbi_load(path, ..., STBI_rgb_alpha) returns for each file with the reference's vertical flip -- produced by
the reference's own stb_image.h compiled where it lies (`make -C oracle stb_ref`, oracle/_ref/libstb_ref.so).
So this script runs only where /root/reference exists (the build container); its outputs are committed and
the test compares astc_b200_load_image with them everywhere.

    make -C oracle stb_ref && python tools/make_image_fixtures.py
"""
import ctypes as C
import io
import struct
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "golden" / "images"


def stb_lib():
    so = ROOT / "oracle" / "_ref" / "libstb_ref.so"
    if not so.exists():
        raise SystemExit("oracle/_ref/libstb_ref.so is missing: run `make -C oracle stb_ref` (needs /root/reference)")
    lib = C.CDLL(str(so))
    lib.stbi_load.restype = C.POINTER(C.c_uint8)
    lib.stbi_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    lib.stbi_failure_reason.restype = C.c_char_p
    lib.stbi_image_free.argtypes = [C.c_void_p]
    return lib


def stb_load(lib, path, flip=True):
    lib.stbi_set_flip_vertically_on_load(1 if flip else 0)
    x, y, c = C.c_int(), C.c_int(), C.c_int()
    p = lib.stbi_load(str(path).encode(), C.byref(x), C.byref(y), C.byref(c), 4)
    if not p:
        return None, lib.stbi_failure_reason().decode()
    arr = np.ctypeslib.as_array(p, shape=(y.value, x.value, 4)).copy()
    lib.stbi_image_free(p)
    return arr, c.value


def picture(w, h, seed=0, alpha=False):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(xx / 9.0 + yy / 17.0 + seed), 127 + 120 * np.cos(xx / 5.0 - yy / 11.0), (xx * 3 + yy * 5 + seed * 40) % 256], -1)
    img = img + rng.normal(0, 10, img.shape)
    out = np.clip(img, 0, 255).astype(np.uint8)
    if alpha:
        a = np.clip(255 * (0.5 + 0.5 * np.sin(xx / 7.0) * np.cos(yy / 5.0)) + rng.normal(0, 20, (h, w)), 0, 255).astype(np.uint8)
        a[: h // 4] = 255
        a[-(h // 5 + 1):] = 0
        out = np.dstack([out, a])
    return out


# ---------------------------------------------------------------------------- hand-written writers
def write_psd(path, rgba, channels=4, depth=8, rle=False):
    h, w = rgba.shape[:2]
    hdr = b"8BPS" + struct.pack(">H6xHIIHH", 1, channels, h, w, depth, 3) + struct.pack(">III", 0, 0, 0)
    planes = [rgba[..., c] for c in range(channels)]
    body = b""
    if not rle:
        body += struct.pack(">H", 0)
        for p in planes:
            body += (p.astype(">u2") * 257).tobytes() if depth == 16 else p.tobytes()
    else:
        rows, counts = [], []
        for p in planes:
            for y in range(h):
                enc = packbits(p[y].tobytes())
                rows.append(enc)
                counts.append(len(enc))
        body += struct.pack(">H", 1) + b"".join(struct.pack(">H", c) for c in counts) + b"".join(rows)
    Path(path).write_bytes(hdr + body)


def packbits(data: bytes) -> bytes:
    out, i, n = bytearray(), 0, len(data)
    while i < n:
        run = 1
        while i + run < n and run < 128 and data[i + run] == data[i]:
            run += 1
        if run >= 3:
            out += bytes([257 - run, data[i]])
            i += run
            continue
        j = i
        while j < n and j - i < 128:
            if j + 2 < n and data[j] == data[j + 1] == data[j + 2]:
                break
            j += 1
        out += bytes([j - i - 1]) + data[i:j]
        i = j
    return bytes(out)


def write_hdr(path, rgb_float, rle=True):
    h, w = rgb_float.shape[:2]
    m = rgb_float.max(axis=2)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, -128).astype(int)
    scale = np.where(m > 1e-32, 256.0 / np.exp2(e.astype(float)), 0.0)
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[..., :3] = np.clip(rgb_float * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    out = bytearray(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode())
    for y in range(h):
        if rle and 8 <= w < 32768:
            out += bytes([2, 2, w >> 8, w & 255])
            for k in range(4):
                line = rgbe[y, :, k].tobytes()
                i = 0
                while i < w:
                    run = 1
                    while i + run < w and run < 127 and line[i + run] == line[i]:
                        run += 1
                    if run >= 4:
                        out += bytes([128 + run, line[i]])
                        i += run
                    else:
                        j = i
                        while j < w and j - i < 128 and not (j + 3 < w and line[j] == line[j + 1] == line[j + 2] == line[j + 3]):
                            j += 1
                        out += bytes([j - i]) + line[i:j]
                        i = j
        else:
            out += rgbe[y].tobytes()
    Path(path).write_bytes(bytes(out))


def write_pic(path, rgba, with_alpha=True, mode=2):
    h, w = rgba.shape[:2]
    hdr = bytearray(b"\x53\x80\xF6\x34" + b"\0" * 84 + b"PICT")
    hdr += struct.pack(">HHIHH", w, h, 0x3F800000, 3, 0)
    packets = [(0xE0, rgba[..., :3])]
    if with_alpha:
        packets.append((0x10, rgba[..., 3:4]))
    for i, (chan, _) in enumerate(packets):
        hdr += bytes([1 if i + 1 < len(packets) else 0, 8, mode, chan])
    body = bytearray()
    for y in range(h):
        for chan, data in packets:
            row = data[y]
            if mode == 0:
                body += row.tobytes()
            else:                                               # mixed RLE: runs of identical pixels, else raw
                x = 0
                while x < w:
                    run = 1
                    while x + run < w and run < 127 and np.array_equal(row[x + run], row[x]):
                        run += 1
                    if run >= 2:
                        body += bytes([127 + run]) + row[x].tobytes()
                        x += run
                    else:
                        j = x + 1
                        while j < w and j - x < 128 and not (j + 1 < w and np.array_equal(row[j], row[j + 1])):
                            j += 1
                        body += bytes([j - x - 1]) + row[x:j].tobytes()
                        x = j
    Path(path).write_bytes(bytes(hdr) + bytes(body))


def write_bmp16(path, rgb, masks=(0xF800, 0x07E0, 0x001F), top_down=False):
    h, w = rgb.shape[:2]
    def pack(v, mask):
        shift = (mask & -mask).bit_length() - 1
        bits = bin(mask).count("1")
        return ((v.astype(np.uint32) >> (8 - bits)) << shift).astype(np.uint32)
    px = (pack(rgb[..., 0], masks[0]) | pack(rgb[..., 1], masks[1]) | pack(rgb[..., 2], masks[2])).astype("<u2")
    row = (w * 2 + 3) & ~3
    rows = [px[y].tobytes().ljust(row, b"\0") for y in (range(h) if top_down else range(h - 1, -1, -1))]
    off = 14 + 40 + 12
    info = struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, 16, 3, row * h, 2835, 2835, 0, 0) + struct.pack("<III", *masks)
    Path(path).write_bytes(b"BM" + struct.pack("<IHHI", off + row * h, 0, 0, off) + info + b"".join(rows))


def write_bmp32_zero_alpha(path, rgb):
    h, w = rgb.shape[:2]
    px = np.zeros((h, w, 4), np.uint8)
    px[..., 0], px[..., 1], px[..., 2] = rgb[..., 2], rgb[..., 1], rgb[..., 0]
    info = struct.pack("<IiiHHIIiiII", 40, w, h, 1, 32, 0, w * h * 4, 2835, 2835, 0, 0)
    Path(path).write_bytes(b"BM" + struct.pack("<IHHI", 54 + w * h * 4, 0, 0, 54) + info + px[::-1].tobytes())


def write_tga16(path, rgb, rle=False, top_down=False):
    h, w = rgb.shape[:2]
    px = ((rgb[..., 0].astype(np.uint16) >> 3) << 10 | (rgb[..., 1].astype(np.uint16) >> 3) << 5 | (rgb[..., 2].astype(np.uint16) >> 3)).astype("<u2")
    rows = px if top_down else px[::-1]
    hdr = struct.pack("<BBBHHBHHHHBB", 0, 0, 10 if rle else 2, 0, 0, 0, 0, 0, w, h, 16, 0x20 if top_down else 0)
    if not rle:
        body = rows.tobytes()
    else:
        flat, body, i = rows.reshape(-1), bytearray(), 0
        while i < len(flat):
            run = 1
            while i + run < len(flat) and run < 128 and flat[i + run] == flat[i]:
                run += 1
            if run > 1:
                body += bytes([0x80 | (run - 1)]) + flat[i:i + 1].tobytes()
                i += run
            else:
                j = i + 1
                while j < len(flat) and j - i < 128 and not (j + 1 < len(flat) and flat[j] == flat[j + 1]):
                    j += 1
                body += bytes([j - i - 1]) + flat[i:j].tobytes()
                i = j
        body = bytes(body)
    Path(path).write_bytes(hdr + body)


def write_gif_subframe(path):
    """Logical screen 40x30 with background index 2, one interlaced 20x12 frame at (7, 5), a transparent index."""
    im = Image.new("P", (20, 12))
    pal = []
    for i in range(256):
        pal += [(i * 7) % 256, (i * 13) % 256, (i * 29) % 256]
    im.putpalette(pal)
    rng = np.random.default_rng(4)
    im.putdata(list(rng.integers(0, 16, 20 * 12)))
    buf = io.BytesIO()
    im.save(buf, format="GIF", transparency=3, interlace=True)
    raw = bytearray(buf.getvalue())
    raw[6:8] = struct.pack("<H", 40)
    raw[8:10] = struct.pack("<H", 30)
    raw[11] = 2                                                 # background colour index
    at = raw.index(b"\x2C", 13 + 3 * (2 << (raw[10] & 7)))
    raw[at + 1:at + 5] = struct.pack("<HH", 7, 5)
    Path(path).write_bytes(bytes(raw))


def main():
    lib = stb_lib()
    OUT.mkdir(parents=True, exist_ok=True)
    for old in OUT.glob("*"):
        old.unlink()
    rgb = picture(61, 45, 1)
    rgba = picture(61, 45, 2, alpha=True)
    small = picture(16, 9, 3, alpha=True)
    P = lambda name: str(OUT / name)
    # JPEG
    Image.fromarray(rgb).save(P("base_444.jpg"), quality=85, subsampling=0)
    Image.fromarray(rgb).save(P("base_420.jpg"), quality=70, subsampling=2)
    Image.fromarray(rgb).save(P("base_422_restart.jpg"), quality=60, subsampling=1, restart_marker_blocks=4)
    Image.fromarray(rgb).save(P("prog_420.jpg"), quality=75, subsampling=2, progressive=True)
    Image.fromarray(rgb).save(P("prog_444_opt.jpg"), quality=92, subsampling=0, progressive=True, optimize=True)
    Image.fromarray(rgb).convert("L").save(P("grey.jpg"), quality=80)
    Image.fromarray(rgb).convert("L").save(P("grey_prog.jpg"), quality=50, progressive=True)
    Image.fromarray(rgb).convert("CMYK").save(P("cmyk.jpg"), quality=80)
    Image.fromarray(rgb).save(P("base_411.jpg"), quality=80, subsampling="4:1:1")
    Image.fromarray(picture(1, 1, 5)).save(P("one_pixel.jpg"), quality=90)
    Image.fromarray(picture(7, 19, 6)).save(P("narrow_420.jpg"), quality=90, subsampling=2)
    # PNG (already covered elsewhere; one of each kind keeps the corpus honest)
    Image.fromarray(rgba).save(P("rgba.png"))
    Image.fromarray(rgb).convert("P").save(P("palette.png"))
    Image.fromarray(rgba[..., 3]).save(P("grey.png"))
    # BMP
    Image.fromarray(rgb).save(P("rgb24.bmp"))
    Image.fromarray(rgba).save(P("rgba32.bmp"))
    Image.fromarray(rgb).convert("P").save(P("pal8.bmp"))
    Image.fromarray(rgb).convert("1").save(P("mono1.bmp"))
    Image.fromarray(rgb).convert("P", colors=16).save(P("pal8_16colours.bmp"))
    write_bmp16(P("rgb565.bmp"), rgb)
    write_bmp16(P("rgb555_topdown.bmp"), rgb, masks=(0x7C00, 0x03E0, 0x001F), top_down=True)
    write_bmp32_zero_alpha(P("rgb32_zero_alpha.bmp"), rgb)
    # TGA
    Image.fromarray(rgb).save(P("rgb24.tga"))
    Image.fromarray(rgba).save(P("rgba32_rle.tga"), compression="tga_rle")
    Image.fromarray(rgba).save(P("rgba32_topdown.tga"), orientation=1)
    Image.fromarray(rgb).convert("L").save(P("grey8_rle.tga"), compression="tga_rle")
    Image.fromarray(rgba).convert("LA").save(P("grey_alpha16.tga"))
    Image.fromarray(rgb).convert("P").save(P("pal8.tga"))
    write_tga16(P("rgb555.tga"), rgb)
    write_tga16(P("rgb555_rle_topdown.tga"), np.repeat(np.repeat(rgb[::4, ::4], 4, 0), 4, 1), rle=True, top_down=True)
    # GIF
    Image.fromarray(rgb).convert("P").save(P("plain.gif"))
    Image.fromarray(rgb).convert("P").save(P("interlaced.gif"), interlace=True)
    g = Image.fromarray(rgb).convert("P", colors=32)
    g.save(P("transparent.gif"), transparency=5)
    write_gif_subframe(P("subframe_background.gif"))
    # PNM
    Path(P("rgb.ppm")).write_bytes(b"P6\n# comment\n%d %d\n255\n" % (rgb.shape[1], rgb.shape[0]) + rgb.tobytes())
    Path(P("grey.pgm")).write_bytes(b"P5 %d %d 255\n" % (rgb.shape[1], rgb.shape[0]) + rgb[..., 1].tobytes())
    # PSD
    write_psd(P("rgba_raw.psd"), rgba)
    write_psd(P("rgba_rle.psd"), rgba, rle=True)
    write_psd(P("rgb_raw.psd"), rgba, channels=3)
    write_psd(P("rgba_16bit.psd"), rgba, depth=16)
    # HDR
    f = (picture(40, 21, 7).astype(np.float64) / 255.0) ** 2.2 * 4.0
    f[:3, :5] = 0.0
    write_hdr(P("rle.hdr"), f)
    write_hdr(P("flat.hdr"), f, rle=False)
    write_hdr(P("narrow_flat.hdr"), f[:, :6])
    # PIC
    write_pic(P("rgba_mixed.pic"), small)
    write_pic(P("rgb_raw.pic"), small, with_alpha=False, mode=0)

    expected, failed = {}, []
    for path in sorted(OUT.iterdir()):
        arr, info = stb_load(lib, path)
        if arr is None:
            failed.append((path.name, info))
            continue
        expected[path.name] = arr
        expected[path.name + ".comp"] = np.array(info)
    np.savez_compressed(OUT / "expected.npz", **expected)
    total = sum(p.stat().st_size for p in OUT.iterdir())
    print(f"{len(expected) // 2} files decoded by stb_image, {len(failed)} refused {failed}, {total / 1024:.0f} KiB under {OUT}")


if __name__ == "__main__":
    main()
