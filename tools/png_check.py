#!/usr/bin/env python
"""Every PNG under a directory (default: the reference's scenes/img) through the product's decoder (ph_png_decode) against PIL,
and the decoded pixels back through the product's encoder (ph_png_encode_rgba8) and PIL again.  A tool (the 4K textures take a
few seconds each), not a test.

    python tools/png_check.py [/root/reference/scenes/img]"""
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402

from portal_b200.host import png_decode, png_encode  # noqa: E402

Image.MAX_IMAGE_PIXELS = None


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/scenes/img"
    bad = 0
    for name in sorted(os.listdir(d)):
        if not name.lower().endswith(".png"):
            continue
        data = open(os.path.join(d, name), "rb").read()
        want = np.asarray(Image.open(io.BytesIO(data)).convert("RGBA"))
        t0 = time.time()
        got = png_decode(data)
        t1 = time.time()
        again = png_encode(got)
        t2 = time.time()
        back = np.asarray(Image.open(io.BytesIO(again)).convert("RGBA"))
        ok = np.array_equal(got, want) and np.array_equal(back, want)
        bad += not ok
        print(f"{name:28s} {want.shape[1]}x{want.shape[0]}  decode {t1 - t0:5.2f} s  encode {t2 - t1:5.2f} s  {len(data) / 1e6:6.2f} MB in, {len(again) / 1e6:6.2f} MB out  "
              f"{'identical to PIL both ways' if ok else 'DIFFERS'}", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
