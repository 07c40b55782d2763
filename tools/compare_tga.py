#!/usr/bin/env python
"""Image comparison for the reference's output format: reads out_<scene>.tga files written by
tga_write_rgb24 (src/common/common.h:86-122 of the reference: 18-byte header, type 2, 24 bpp, bottom-left origin, BGR)
or the golden fixtures under tests/golden/render_<scene>.npz, prints per-channel RMSE / bias / max difference in 8-bit
units, and optionally exports PNGs (image and amplified difference).

  python tools/compare_tga.py out_large.tga tests/golden/render_large.npz [--png diff.png] [--export a.png]
"""
import argparse
import struct
import sys
import zlib

import numpy as np


def read_image(path):
    """-> uint8 [H, W, 3] RGB, row 0 = TOP of the picture"""
    if path.endswith(".npz"):
        return np.ascontiguousarray(np.load(path)["rgb"][::-1])          # fixtures keep the reference's bottom-up rows
    raw = open(path, "rb").read()
    if len(raw) < 18 or raw[2] != 2 or raw[16] != 24:
        raise ValueError("%s: not an uncompressed 24-bit TGA" % path)
    w, h = struct.unpack_from("<HH", raw, 12)
    body = np.frombuffer(raw, np.uint8, w * h * 3, 18 + raw[0]).reshape(h, w, 3)
    top_down = bool(raw[17] & 0x20)
    img = body[:, :, ::-1]                                               # BGR -> RGB
    return np.ascontiguousarray(img if top_down else img[::-1])


def write_png(path, rgb):
    h, w, _ = rgb.shape
    raw = b"".join(b"\x00" + rgb[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def compare(a, b):
    if a.shape != b.shape:
        raise ValueError("size mismatch %s vs %s" % (a.shape, b.shape))
    d = a.astype(np.float64) - b.astype(np.float64)
    return {"rmse": float(np.sqrt((d * d).mean())), "rmse_rgb": [float(np.sqrt((d[..., c] ** 2).mean())) for c in range(3)],
            "bias_rgb": [float(d[..., c].mean()) for c in range(3)], "max_abs": int(np.abs(d).max()),
            "identical": bool(np.array_equal(a, b))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("a")
    ap.add_argument("b", nargs="?")
    ap.add_argument("--png", help="write the amplified (x16) absolute difference")
    ap.add_argument("--export", help="write image A as PNG")
    args = ap.parse_args()
    a = read_image(args.a)
    if args.export:
        write_png(args.export, a)
    if not args.b:
        print("%s: %dx%d" % (args.a, a.shape[1], a.shape[0]))
        return 0
    b = read_image(args.b)
    r = compare(a, b)
    print("rmse %.4f /255 (r %.4f g %.4f b %.4f)  bias (%.3f %.3f %.3f)  max |diff| %d  identical %s" %
          (r["rmse"], *r["rmse_rgb"], *r["bias_rgb"], r["max_abs"], r["identical"]))
    if args.png:
        write_png(args.png, np.clip(np.abs(a.astype(np.int32) - b.astype(np.int32)) * 16, 0, 255).astype(np.uint8))
    return 0


if __name__ == "__main__":
    sys.exit(main())
