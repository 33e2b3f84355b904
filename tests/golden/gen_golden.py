#!/usr/bin/env python3
"""Regenerate tests/golden/*.  Run from the repo root:  python tests/golden/gen_golden.py

The reference ships no fixtures and cannot run in this container (tensorflow / rasterio / GDAL missing), so the
golden vectors are produced by the SAME third-party codecs the reference delegates to, where they are importable:
  * libtiff 4.7.1 (through cv2 and Pillow)  - LZW TIFF encode, i.e. what GDAL's GTiff driver links against
  * libpng / zlib (through Pillow)          - PNG encode
  * google.protobuf (upb)                   - tensorflow.Example deterministic serialisation (dynamic descriptor)
  * numpy.ma                                - the reference's own median call (_descartes_img_chips.py:565-567)
  * RFC 3720 B.4 / TensorFlow record format - CRC-32C and frame vectors
plus GDAL-style tiled files from synthetic.tiff_bytes that were checked against libtiff when generated.
Nothing here imports oracle/ or the product.
"""
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

import cv2  # noqa: E402
from PIL import Image  # noqa: E402

import synthetic as syn  # noqa: E402


def example_class():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fdp = descriptor_pb2.FileDescriptorProto(name="example_golden.proto", package="tensorflow", syntax="proto3")

    def msg(name):
        m = fdp.message_type.add()
        m.name = name
        return m
    msg("BytesList").field.add(name="value", number=1, type=12, label=3)
    f = msg("FloatList").field.add(name="value", number=1, type=2, label=3)
    f.options.packed = True
    f = msg("Int64List").field.add(name="value", number=1, type=3, label=3)
    f.options.packed = True
    fe = msg("Feature")
    fe.oneof_decl.add(name="kind")
    for i, (n, t) in enumerate([("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List")]):
        fe.field.add(name=n, number=i + 1, type=11, label=1, type_name=".tensorflow." + t, oneof_index=0)
    fs = msg("Features")
    en = fs.nested_type.add(name="FeatureEntry")
    en.options.map_entry = True
    en.field.add(name="key", number=1, type=9, label=1)
    en.field.add(name="value", number=2, type=11, label=1, type_name=".tensorflow.Feature")
    fs.field.add(name="feature", number=1, type=11, label=3, type_name=".tensorflow.Features.FeatureEntry")
    msg("Example").field.add(name="features", number=1, type=11, label=1, type_name=".tensorflow.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow.Example"))


def pb_example(Ex, img, lab, key, as_bytes):
    e = Ex()
    h, w, c = img.shape
    if as_bytes:
        e.features.feature["image/image_data"].bytes_list.value.append(img.tobytes())
        e.features.feature["target/target_data"].bytes_list.value.append(lab.tobytes())
    else:
        e.features.feature["image/image_data"].float_list.value.extend(img.flatten().astype(np.float32).tolist())
        e.features.feature["target/target_data"].float_list.value.extend(lab.flatten().astype(np.float32).tolist())
    for k, v in (("image/height", h), ("image/width", w), ("image/channels", c), ("target/height", lab.shape[0]),
                 ("target/width", lab.shape[1])):
        e.features.feature[k].int64_list.value.append(v)
    e.features.feature["identifier"].bytes_list.value.append(key.encode())
    return e.SerializeToString(deterministic=True)


def main():
    out = {}
    # --- CRC-32C / frame known answers
    out["crc32c"] = [
        {"hex": b"123456789".hex(), "crc": 0xE3069283, "masked": 0xC78AB0E5},
        {"hex": bytes(32).hex(), "crc": 0x8A9136AA, "masked": 0x0FD7FFFA},
        {"hex": (b"\xff" * 32).hex(), "crc": 0x62A8AB43, "masked": 0xF909B029},
        {"hex": bytes(range(32)).hex(), "crc": 0x46DD794E, "masked": 0x951F7892},
        {"hex": bytes(range(31, -1, -1)).hex(), "crc": 0x113FDB5C, "masked": 0x593B0D57},
    ]
    out["frames"] = [
        {"data_hex": "", "frame_hex": "000000000000000029039807d8ea82a2"},
        {"data_hex": b"abc".hex(), "frame_hex": "0300000000000000b099490e6162636e57f121"},
        {"data_hex": bytes(range(16)).hex(), "frame_hex": "100000000000000095fbfe18" + bytes(range(16)).hex() + "6a9e5ab4"},
    ]
    # --- Example bytes from google.protobuf
    Ex = example_class()
    img8, lab8, key8 = syn.cfg1_chip(0, size=24)
    img16, lab16, key16 = syn.cfg3_chip(0, size=20)
    np.save(os.path.join(HERE, "chip8_img.npy"), img8)
    np.save(os.path.join(HERE, "chip8_lab.npy"), lab8)
    np.save(os.path.join(HERE, "chip16_img.npy"), img16)
    np.save(os.path.join(HERE, "chip16_lab.npy"), lab16)
    open(os.path.join(HERE, "example_bytes.bin"), "wb").write(pb_example(Ex, img8, lab8, key8, True))
    open(os.path.join(HERE, "example_float.bin"), "wb").write(pb_example(Ex, img16, lab16, key16, False))
    out["example"] = {"key8": key8, "key16": key16}
    # --- encoded chips: libtiff (cv2: strips + predictor 2; Pillow: 1-band), libpng (Pillow), GDAL-style tiled (synthetic)
    big16, biglab, _ = syn.cfg3_chip(1, size=96)
    np.save(os.path.join(HERE, "tiff_img.npy"), big16)
    np.save(os.path.join(HERE, "tiff_lab.npy"), biglab)
    ok, enc = cv2.imencode(".tif", big16[..., [2, 1, 0, 3]], [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    assert ok
    open(os.path.join(HERE, "libtiff_cv2_lzw_u16x4.tif"), "wb").write(enc.tobytes())
    bio = io.BytesIO()
    Image.fromarray(biglab).save(bio, format="TIFF", compression="tiff_lzw")
    open(os.path.join(HERE, "libtiff_pil_lzw_u8.tif"), "wb").write(bio.getvalue())
    bio = io.BytesIO()
    Image.fromarray(biglab).save(bio, format="TIFF", compression="tiff_adobe_deflate")
    open(os.path.join(HERE, "libtiff_pil_deflate_u8.tif"), "wb").write(bio.getvalue())
    gd = syn.tiff_bytes(big16, tile=64)
    chk = cv2.imdecode(np.frombuffer(syn.tiff_bytes(big16, tile=64, photometric=2), np.uint8), cv2.IMREAD_UNCHANGED)
    assert np.array_equal(chk[..., [2, 1, 0, 3]], big16)                    # libtiff reads our tiled LZW identically
    open(os.path.join(HERE, "gdalstyle_tiled_lzw_u16x4.tif"), "wb").write(gd)
    gl = syn.tiff_bytes(biglab, tile=64, nodata=255)
    assert np.array_equal(np.array(Image.open(io.BytesIO(gl))), biglab)
    open(os.path.join(HERE, "gdalstyle_tiled_lzw_label.tif"), "wb").write(gl)
    png_img, png_lab, _ = syn.cfg1_chip(1, size=64)
    np.save(os.path.join(HERE, "png_img.npy"), png_img)
    np.save(os.path.join(HERE, "png_lab.npy"), png_lab)
    open(os.path.join(HERE, "libpng_rgb.png"), "wb").write(syn.png_bytes(png_img))
    open(os.path.join(HERE, "libpng_label.png"), "wb").write(syn.png_bytes(png_lab))
    # --- np.ma.median, the reference's own arithmetic
    stack, valid = syn.cfg4_tile(5, T=16, H=12, W=10, B=8)
    np.save(os.path.join(HERE, "median_stack.npy"), stack)
    np.save(os.path.join(HERE, "median_valid.npy"), valid)
    rep = np.repeat(valid[..., None], 8, axis=-1)
    med = np.ma.median(np.ma.masked_where(rep == 0, stack), axis=0)
    np.save(os.path.join(HERE, "median_out.npy"), med.filled(0.0))
    np.save(os.path.join(HERE, "median_mask.npy"), np.ma.getmaskarray(med))
    out["median_kat"] = [{"values": [5, 1, 9, 7], "valid": v, "median": m} for v, m in
                         (("1111", 6.0), ("1110", 5.0), ("0101", 4.0), ("0010", 9.0), ("0000", None))]
    # --- partition / identifier known answers
    import random
    idx = list(range(20))
    random.seed(12345)
    random.shuffle(idx)
    out["shuffle20"] = idx
    out["identifier"] = {"path": "x/images/60#2#10.0#43#-380#3491.tif", "key": "60:2:10.0:43:-380:3491"}
    json.dump(out, open(os.path.join(HERE, "vectors.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
