#!/usr/bin/env python3
"""Regenerate tests/golden/jpeg_*.  Run from the repo root:  python tests/golden/gen_golden_jpeg.py

JPEG fixtures for the .jpg chip path (tf.image.decode_jpeg behind ImageCoder.decode_jpeg, reference
_img_to_tf_threaded.py:36-38,51-56,97-103).  TensorFlow cannot run here; the files AND the expected pixels come from
libjpeg-turbo — the libjpeg TensorFlow links — through cv2 (encode with chosen sampling / restart interval, decode) and
Pillow (second decoder: must agree).  Nothing here imports oracle/ or the product.
"""
import io
import os

import cv2
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
S = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
     "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
     "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
# name: (height, width, channels, sampling, quality, restart interval in MCUs)
CASES = {"420_q90": (40, 48, 3, "420", 90, 0), "422_q75_rst": (29, 37, 3, "422", 75, 2), "444_q100": (24, 24, 3, "444", 100, 0),
         "440_q85": (19, 2, 3, "440", 85, 0), "411_q60_rst": (17, 45, 3, "411", 60, 1), "grey_q90": (20, 33, 1, "444", 90, 0)}


def chip(h, w, c, rng):
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 90 * np.sin(xx / (3.0 + i)) * np.cos(yy / (4.0 + 2 * i)) for i in range(c)], -1)
    return np.clip(img + rng.normal(0, 14, img.shape), 0, 255).astype(np.uint8)


def main():
    rng = np.random.default_rng(20261018)
    for name, (h, w, c, sf, q, rst) in CASES.items():
        img = chip(h, w, c, rng)
        ok, buf = cv2.imencode(".jpg", img if c == 3 else img[..., 0],
                               [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, S[sf],
                                cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
        assert ok
        data = buf.tobytes()
        ref = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
        ref = ref[..., ::-1] if c == 3 else ref[..., None]                  # cv2 returns B,G,R
        pil = np.array(Image.open(io.BytesIO(data)))
        assert np.array_equal(pil.reshape(ref.shape), ref), name           # two libjpeg front ends, one answer
        open(os.path.join(HERE, "jpeg_%s.jpg" % name), "wb").write(data)
        np.save(os.path.join(HERE, "jpeg_%s.npy" % name), np.ascontiguousarray(ref))
        print(name, len(data), ref.shape)
    # encoder fixtures (convert_png_to_jpg): pixels in, the file libjpeg-turbo writes for them out (cv2.imencode: JFIF
    # density 1 x 1 without a unit; TensorFlow's 300 x 300 dpi differs in those five header bytes only)
    for name, (h, w, c, q) in {"rgb_q100": (24, 40, 3, 100), "rgb_q75": (37, 29, 3, 75), "grey_q100": (20, 33, 1, 100)}.items():
        img = chip(h, w, c, rng)
        ok, buf = cv2.imencode(".jpg", np.ascontiguousarray(img[..., ::-1]) if c == 3 else img[..., 0], [cv2.IMWRITE_JPEG_QUALITY, q])
        assert ok
        np.save(os.path.join(HERE, "jpegenc_%s_in.npy" % name), img)
        open(os.path.join(HERE, "jpegenc_%s_out.jpg" % name), "wb").write(buf.tobytes())
        print("encoder", name, len(buf))
    # a progressive file: out of scope, must be reported (status 3), never mis-decoded
    f = io.BytesIO()
    Image.fromarray(chip(16, 16, 3, rng)).save(f, "JPEG", quality=80, progressive=True)
    open(os.path.join(HERE, "jpeg_progressive.jpg"), "wb").write(f.getvalue())


if __name__ == "__main__":
    main()
