"""-m gpu: K1 chip decode (TIFF LZW / DEFLATE / stored, PNG inflate + un-filter) through the C ABI vs the oracle.

Bit-exact on every pixel.  Inputs come from three independent encoders: libtiff via cv2, libpng via Pillow,
and the synthetic writers (GDAL-style tiled LZW, planar, big-endian, chosen PNG filters / deflate block types).
"""
import zlib

import cv2
import numpy as np
import pytest

import synthetic as syn
from oracle import imagecodecs as oic

pytestmark = pytest.mark.gpu


def _decode(dev, blobs):
    from dl_image_segmentation_b200 import _codec
    arrays, status = _codec.decode_blobs(blobs, device=dev)
    return [None if a is None else a.cpu().numpy() for a in arrays], list(status)


def _check(dev, blobs, wants=None):
    got, status = _decode(dev, blobs)
    for i, b in enumerate(blobs):
        want = oic.decode_image(b) if wants is None else wants[i]
        assert status[i] == 0, (i, status[i])
        assert got[i].dtype == want.dtype and got[i].shape == want.shape, (i, got[i].dtype, got[i].shape, want.shape)
        np.testing.assert_array_equal(got[i], want, err_msg="image %d" % i)


def test_gdal_style_tiled_lzw_chip_pairs(dev):
    blobs, wants = [], []
    for i in range(3):
        img, lab, _ = syn.cfg3_chip(i)
        blobs += [syn.tiff_bytes(img, tile=256), syn.tiff_bytes(lab, tile=256, nodata=255)]
        wants += [img, lab[:, :, None]]
    _check(dev, blobs, wants)


def test_libtiff_encoded_strips_predictor2(dev):
    img, lab, _ = syn.cfg3_chip(7, size=200)
    ok, enc = cv2.imencode(".tif", img[..., [2, 1, 0, 3]], [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    ok2, encl = cv2.imencode(".tif", lab, [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    ok3, enc8 = cv2.imencode(".tif", (img[..., :3] >> 6).astype(np.uint8), [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    assert ok and ok2 and ok3
    _check(dev, [enc.tobytes(), encl.tobytes(), enc8.tobytes()],
           [img, lab[:, :, None], (img[..., :3] >> 6).astype(np.uint8)[..., ::-1]])


@pytest.mark.parametrize("kw", [
    dict(tile=None, predictor=2), dict(tile=128, predictor=2, planar=2), dict(tile=256, compression="deflate"),
    dict(tile=None, compression="none", big_endian=True), dict(tile=256, big_endian=True, predictor=2),
    dict(tile=64, planar=2), dict(tile=None, rows_per_strip=7, compression="deflate", predictor=2),
    dict(tile=32, compression="none"), dict(tile=None, rows_per_strip=1000)])
def test_tiff_variants(dev, kw):
    img, lab, _ = syn.cfg3_chip(11, size=300)
    img = img[:300, :280]
    _check(dev, [syn.tiff_bytes(img, **kw), syn.tiff_bytes(lab[:77, :130], **kw)], [img, lab[:77, :130, None]])


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint32, np.float32, np.float64])
def test_tiff_dtypes(dev, dtype):
    rng = np.random.default_rng(8)
    base = syn.smooth_field(rng, 150, 170)
    img = (np.stack([base, base[::-1], base.T[:150, :170] if False else base * 0.5], -1) * 1000).astype(dtype)
    kws = [dict(tile=64), dict(tile=None, compression="deflate")]
    if np.issubdtype(dtype, np.integer):
        kws.append(dict(tile=64, predictor=2, big_endian=True))
    _check(dev, [syn.tiff_bytes(img, **kw) for kw in kws], [img] * len(kws))


def test_lzw_pathological_streams(dev):
    rng = np.random.default_rng(9)
    cases = [np.zeros((256, 256), np.uint8),                                     # one long run: KwKwK chains
             np.full((100, 300), 7, np.uint8),
             rng.integers(0, 256, (256, 256), dtype=np.uint8),                   # incompressible: all literals, many Clears
             np.tile(np.arange(256, dtype=np.uint8), (256, 1)),                  # periodic
             np.repeat(rng.integers(0, 4, (64, 64), dtype=np.uint8), 4, axis=1), # short runs
             rng.integers(0, 2, (301, 257), dtype=np.uint8),
             np.zeros((1, 1), np.uint8), np.arange(5, dtype=np.uint8).reshape(1, 5)]
    blobs = [syn.tiff_bytes(c, tile=None, rows_per_strip=c.shape[0]) for c in cases] + \
            [syn.tiff_bytes(c, tile=256) for c in cases[:4]]
    _check(dev, blobs, [c[:, :, None] for c in cases] + [c[:, :, None] for c in cases[:4]])


def test_png_pillow_chip_pairs(dev):
    blobs, wants = [], []
    for i in range(4):
        img, lab, _ = syn.cfg1_chip(i)
        blobs += [syn.png_bytes(img), syn.png_bytes(lab)]
        wants += [img, lab[:, :, None]]
    _check(dev, blobs, wants)


@pytest.mark.parametrize("filters", [(0,), (1,), (2,), (3,), (4,), (0, 1, 2, 3, 4), (4, 3, 4, 1)])
@pytest.mark.parametrize("channels", [1, 2, 3, 4])
def test_png_each_filter_and_colour_type(dev, filters, channels):
    rng = np.random.default_rng(channels)
    img = np.stack([(syn.smooth_field(rng, 70, 93) * 255).astype(np.uint8) for _ in range(channels)], -1)
    img[::7] = rng.integers(0, 256, img[::7].shape, dtype=np.uint8)
    _check(dev, [syn.png_bytes_manual(img, filter_types=filters, idat_chunk=997)], [img])


def test_png_deflate_block_types(dev):
    img, lab, _ = syn.cfg1_chip(3, size=96)
    blobs = [syn.png_bytes_manual(img, zlevel=0),                                 # stored blocks
             syn.png_bytes_manual(img, strategy=zlib.Z_FIXED),                    # fixed Huffman
             syn.png_bytes_manual(img, zlevel=9),                                 # dynamic Huffman
             syn.png_bytes_manual(lab, zlevel=1), syn.png_bytes_manual(np.zeros((300, 500, 3), np.uint8)),   # long matches
             syn.png_bytes_manual(lab, strategy=zlib.Z_HUFFMAN_ONLY), syn.png_bytes_manual(lab, strategy=zlib.Z_RLE)]
    _check(dev, blobs)


@pytest.mark.parametrize("as_tf", [True, False])
def test_png_palette_subbyte_and_16bit_flavours(dev, as_tf):
    """Palette, 1/2/4-bit and 16-bit PNGs under both presentations (tf.image.decode_png vs rasterio/GDAL) == oracle
    (which tests/test_oracle_golden.py pins against Pillow and OpenCV's libpng), mixed with ordinary chips in one batch."""
    import torch
    from dl_image_segmentation_b200 import _codec
    from test_oracle_golden import _png_flavour_cases
    cases = _png_flavour_cases()
    img, lab, _ = syn.cfg1_chip(5, size=64)
    blobs = [syn.png_bytes(img)] + [b for _, b in cases] + [syn.png_bytes(lab)]
    wants = [oic.decode_png(b, as_tf) for b in blobs]
    arrays, status, infos = _codec.decode_blobs(blobs, device=dev, want_infos=True, png_as_tf=as_tf)
    assert not np.asarray(status).any(), list(status)
    for (name, _), a, w, info in zip([("rgb", 0)] + cases + [("label", 0)], arrays, wants, infos):
        g = a.cpu().numpy() if a.dtype != torch.uint16 else a.view(torch.int16).cpu().numpy().view(np.uint16)
        assert g.dtype == w.dtype and g.shape == w.shape, (name, g.dtype, g.shape, w.dtype, w.shape)
        np.testing.assert_array_equal(g, w, err_msg=name)
        assert (info.height, info.width, info.samples) == w.shape
        p = _codec.probe(blobs[0], png_as_tf=as_tf)
        assert (p.height, p.width, p.samples, p.png_bit_depth, p.png_color_type) == (64, 64, 3, 8, 2)
    # Adam7 files of awkward geometries (empty passes) in one batch; sub-byte interlaced stays out of scope (flagged)
    from test_oracle_golden import _png_interlaced_cases
    inter = _png_interlaced_cases()
    blobs = [b for _, b, _, _, _ in inter]
    arrays, status = _codec.decode_blobs(blobs, device=dev, png_as_tf=as_tf)
    assert not np.asarray(status).any(), list(status)
    for (name, b, _, _, _), a in zip(inter, arrays):
        w = oic.decode_png(b, as_tf)
        g = a.cpu().numpy() if a.dtype != torch.uint16 else a.view(torch.int16).cpu().numpy().view(np.uint16)
        assert g.dtype == w.dtype and g.shape == w.shape, name
        np.testing.assert_array_equal(g, w, err_msg=name)
    assert _codec.probe(syn.png_bytes_flavour(np.zeros((5, 5), np.uint8), 4, 0, interlace=True), png_as_tf=as_tf).status == 3


def test_error_behaviour_of_zlib_and_libpng_is_mirrored(dev):
    """What the reference's decoders reject, the device path rejects: zlib's inflate_table refuses an incomplete
    literal/length set in the block header (even if the data never uses a missing code), and libpng treats a CRC
    mismatch in a critical chunk (IHDR, IDAT) as fatal.  The hand-made stream with a complete set is the control."""
    ok = syn.handmade_dynamic_deflate(None, {0: 2, 65: 2, 66: 2, 256: 2}, [0, 65, 256])
    incomplete = syn.handmade_dynamic_deflate(None, {0: 2, 65: 2, 256: 2}, [0, 65, 256])
    one_bit_only = syn.handmade_dynamic_deflate(None, {0: 1, 256: 2, 65: 2}, [0, 65, 256])         # complete, control
    good = syn.png_bytes_raw_zlib(1, 1, 1, ok)
    bad_idat = bytearray(good)
    bad_idat[-16] ^= 1                                              # last CRC byte of the IDAT chunk
    bad_ihdr = bytearray(good)
    bad_ihdr[8 + 8 + 13] ^= 0x80                                    # first CRC byte of the IHDR chunk
    # Adler-32 trailer wrong / missing (chunk CRC consistent): "incorrect data check" in libpng and libtiff alike
    img = syn.cfg1_chip(0, size=32)[0]
    z = bytearray(zlib.compress(syn.png_filter_rows(img, (0, 1, 2, 3, 4)), 6))
    z[-1] ^= 0x55
    tif = bytearray(syn.tiff_bytes(img, tile=None, compression="deflate"))
    t = oic.parse_tiff(bytes(tif))
    tif[t["offsets"][-1] + t["counts"][-1] - 1] ^= 0x55
    blobs = [good, syn.png_bytes_raw_zlib(1, 1, 1, incomplete), syn.png_bytes_raw_zlib(1, 1, 1, one_bit_only),
             bytes(bad_idat), bytes(bad_ihdr), syn.png_bytes_raw_zlib(32, 32, 3, bytes(z)),
             syn.png_bytes_raw_zlib(32, 32, 3, bytes(z[:-4])), bytes(tif)]
    want = []
    for b in blobs:
        try:
            want.append(oic.decode_image(b))
        except oic.DecodeError:
            want.append(None)
    assert [w is None for w in want] == [False, True, False, True, True, True, True, True]
    got, status = _decode(dev, blobs)
    for g, w, st in zip(got, want, status):
        if w is None:
            assert st != 0 and g is None
        else:
            assert st == 0
            np.testing.assert_array_equal(g, w)


def test_corrupt_chips_are_skipped_not_fatal(dev):
    img, lab, _ = syn.cfg3_chip(1, size=128)
    good = syn.tiff_bytes(img, tile=64)
    png = syn.png_bytes(syn.cfg1_chip(0, size=64)[0])
    trunc = good[:len(good) // 2]
    garb = bytearray(good)
    garb[len(good) // 2:len(good) // 2 + 400] = bytes(400)
    badpng = bytearray(png)
    badpng[60:90] = bytes(30)
    blobs = [good, trunc, bytes(garb), b"not an image at all", png, bytes(badpng), png[:100], good]
    got, status = _decode(dev, blobs)
    assert status[0] == 0 and status[4] == 0 and status[7] == 0
    np.testing.assert_array_equal(got[0], img)
    np.testing.assert_array_equal(got[7], img)
    for i in (1, 3, 6):
        assert status[i] != 0 and got[i] is None
    for i, b in ((2, bytes(garb)), (5, bytes(badpng))):          # payload damage: either flagged, or decodes like the oracle
        try:
            want = oic.decode_image(b)
        except oic.DecodeError:
            want = None
        if status[i] == 0 and want is not None:
            np.testing.assert_array_equal(got[i], want)
        if i == 5:                                               # PNG: chunk CRC + Adler-32 make damage detectable
            assert want is None and status[i] != 0


def test_probe_header_only(dev):
    from dl_image_segmentation_b200 import _codec
    img, lab, _ = syn.cfg3_chip(0, size=100)
    i1 = _codec.probe(syn.tiff_bytes(img, tile=64, nodata=None))
    assert (i1.height, i1.width, i1.samples, i1.status) == (100, 100, 4, 0)
    i2 = _codec.probe(syn.tiff_bytes(lab, nodata=255))
    assert (i2.height, i2.width, i2.samples, i2.has_nodata, i2.nodata) == (100, 100, 1, 1, 255.0)
    i3 = _codec.probe(syn.png_bytes(syn.cfg1_chip(0, size=40)[0]))
    assert (i3.height, i3.width, i3.samples, i3.format) == (40, 40, 3, 2)
    assert _codec.probe(b"garbage").status != 0


# ------------------------------------------------------------------------------------------------ round-2 decoders: fuzz
def _grey_png_from_zlib(rows, width, zstream):
    """A grey 8-bit PNG whose IDAT is the given zlib stream over `rows` x (1 + width) filtered bytes."""
    return syn.png_bytes_raw_zlib(width, rows, 1, zstream)


def test_inflate_chunk_parallel_decode_against_zlib_fuzz(dev):
    """The chunk-parallel symbol decode (self-synchronising restart rounds, staged literals, dependency-ordered matches,
    serial fallback for match-dense blocks) on streams of every character: incompressible, runs, short and long matches,
    all zlib strategies / levels (fixed Huffman, Huffman-only, RLE, stored), several blocks, sizes around the chunk
    boundaries — decoded bytes must equal what zlib itself returns (filter type 0 rows make the PNG's pixels = the data)."""
    import zlib
    rng = np.random.default_rng(2024)
    blobs, want = [], []

    def add(data, level, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=15, width=None):
        width = width or 255
        rows = (len(data) + width - 1) // width
        data = data + bytes(rows * width - len(data))
        arr = np.frombuffer(data, np.uint8).reshape(rows, width)
        filtered = np.concatenate([np.zeros((rows, 1), np.uint8), arr], axis=1).tobytes()        # filter type 0 on every row
        co = zlib.compressobj(level, zlib.DEFLATED, wbits, 8, strategy)
        z = co.compress(filtered) + co.flush()
        assert zlib.decompress(z) == filtered
        blobs.append(_grey_png_from_zlib(rows, width, z))
        want.append(arr[:, :, None])
    for n in (1, 7, 79, 80, 81, 639, 640, 641, 5000, 70000, 200000):
        add(rng.integers(0, 256, n, dtype=np.uint8).tobytes(), 6)                                  # incompressible: all literals
    for level in (1, 6, 9):
        noise = rng.integers(0, 4, 150000, dtype=np.uint8).tobytes()                               # short codes, few matches
        add(noise, level)
        add(noise, level, zlib.Z_HUFFMAN_ONLY)
        text = bytes(rng.choice(np.frombuffer(b"abcdefgh \n", np.uint8), 120000))                  # many short matches
        add(text, level)
        add(text, level, zlib.Z_FIXED)
        runs = np.repeat(rng.integers(0, 256, 3000, dtype=np.uint8), rng.integers(1, 300, 3000)).tobytes()   # long runs: dist 1
        add(runs, level)
        add(runs, level, zlib.Z_RLE)
    add(rng.integers(0, 256, 90000, dtype=np.uint8).tobytes(), 0)                                  # stored blocks only
    mixed = b"".join([rng.integers(0, 256, 30000, dtype=np.uint8).tobytes(), bytes(40000), rng.integers(0, 3, 50000, dtype=np.uint8).tobytes(),
                      b"xyz" * 20000])                                                             # the character changes block by block
    add(mixed, 6)
    add(mixed, 6, wbits=9)                                                                         # 512-byte window: short distances only
    co = zlib.compressobj(6)
    parts = b""
    src = rng.integers(0, 16, 100000, dtype=np.uint8).tobytes()
    rows = len(src) // 250
    filt = np.concatenate([np.zeros((rows, 1), np.uint8), np.frombuffer(src[:rows * 250], np.uint8).reshape(rows, 250)], axis=1).tobytes()
    for k in range(0, len(filt), 7001):                                                            # Z_FULL_FLUSH: many small blocks + empty stored blocks
        parts += co.compress(filt[k:k + 7001]) + co.flush(zlib.Z_FULL_FLUSH)
    parts += co.flush()
    blobs.append(_grey_png_from_zlib(rows, 250, parts))
    want.append(np.frombuffer(src[:rows * 250], np.uint8).reshape(rows, 250, 1))
    got, status = _decode(dev, blobs)
    for i, (g, w) in enumerate(zip(got, want)):
        assert status[i] == 0, (i, status[i])
        assert np.array_equal(g, w), i
    # damaged streams: a flipped bit anywhere must never be silently accepted with the same bytes AND a valid check
    bad = []
    for i in [k for k, b in enumerate(blobs) if len(b) > 4000][::5][:4]:
        b = bytearray(blobs[i])
        pos = b.find(b"IDAT") + 4 + 10 + (i * 37) % 200
        b[pos] ^= 0x10
        bad.append(bytes(b))
    _, st = _decode(dev, bad)
    assert all(s != 0 for s in st), list(st)


def test_lzw_segment_parallel_decode_fuzz(dev):
    """The segment-parallel LZW decoder on tiles of every character (incompressible: strings of 1-2 bytes; constant and
    run-length data: strings hundreds of bytes long, several output windows per segment; text-like; tiny tiles; tiles that
    end in mid-string; strips instead of tiles) against the oracle's sequential table decoder."""
    rng = np.random.default_rng(77)
    blobs = []
    S = 96
    imgs = [rng.integers(0, 65536, (S, S, 4)).astype(np.uint16),                                  # incompressible
            np.zeros((S, S, 4), np.uint16), np.full((S, S, 4), 0xABCD, np.uint16),                 # one very long string chain
            np.repeat(rng.integers(0, 256, (S, S // 8, 1)), 8, axis=1).astype(np.uint8).repeat(3, axis=2),      # runs
            (np.arange(S * S * 2).reshape(S, S, 2) % 251).astype(np.uint8),                       # periodic
            rng.integers(0, 3, (S, S, 1)).astype(np.uint8),                                        # tiny alphabet: deep forests
            rng.integers(0, 256, (5, 7, 1)).astype(np.uint8), rng.integers(0, 256, (1, 1, 1)).astype(np.uint8)]
    for im in imgs:
        blobs += [syn.tiff_bytes(im, tile=32), syn.tiff_bytes(im, tile=256), syn.tiff_bytes(im, tile=None, rows_per_strip=3),
                  syn.tiff_bytes(im, tile=64, predictor=2)]
    big = np.zeros((512, 512), np.uint8)
    big[100:300, 50:400] = 9
    blobs.append(syn.tiff_bytes(big, tile=256))                                                   # label-like: segments of > 12 KiB output
    got, status = _decode(dev, blobs)
    for i, b in enumerate(blobs):
        assert status[i] == 0, (i, status[i])
        assert np.array_equal(got[i], oic.decode_image(b)), i


def test_lzw_streams_that_alternate_full_and_short_segments(dev):
    """Strips of noisy 16-bit chips: every stream is one full-table segment (3838 codes) followed by a short one, thousands
    of streams per batch — the decoder tries its 5-codes-per-thread pass first, gives up without a barrier and runs the
    15-codes-per-thread pass on the same segment (the two passes once shared the word that holds the segment's end: a
    late warp of the short pass read the full pass's initial value and went its own way).  Restart streams
    (a Clear every 1024 input bytes, what this package's GeoTIFF writer emits) are all short segments."""
    blobs, wants = [], []
    for i in range(6):
        img, lab, _ = syn.cfg3_chip(20 + i)
        blobs += [syn.tiff_bytes(img, tile=None, predictor=2, photometric=2), syn.tiff_bytes(lab, tile=None, predictor=2),
                  syn.tiff_bytes(img, tile=256, lzw_restart=1024), syn.tiff_bytes(img, tile=256, lzw_restart=48)]
        wants += [img, lab[:, :, None], img, img]
    for _ in range(3):
        _check(dev, blobs, wants)


@pytest.mark.parametrize("channels", [1, 2, 3, 4])
def test_png_unfilter_with_eight_warps_per_image(dev, channels):
    """Images of 96 rows and more (8-bit, progressive, rows of at most 4096 bytes) are un-filtered by eight warps that hand
    the last row of every band of 32 rows on through shared memory: every filter type and mixes of them, heights that end
    in a partial band, one band per warp and several, the longest row taken and the first one left to the one-warp kernel,
    many images at once so that the CTAs of several waves overlap."""
    rng = np.random.default_rng(10 + channels)
    blobs, wants = [], []
    shapes = [(96, 33), (97, 64), (256, 256), (300, 130), (129, 4096 // channels), (128, 4096 // channels + 1), (700, 40)]
    for j, (h, w) in enumerate(shapes):
        img = rng.integers(0, 256, (h, w, channels), dtype=np.uint8)
        img[h // 3:h // 2] = (syn.smooth_field(rng, h // 2 - h // 3, w)[..., None] * 255).astype(np.uint8)
        for filters in ((4,), (3,), (2,), (1, 0), (0, 1, 2, 3, 4), (4, 3, 2)):
            if j >= 4 and filters != (0, 1, 2, 3, 4):
                continue
            blobs.append(syn.png_bytes_manual(img, filter_types=filters, zlevel=1))
            wants.append(img)
    blobs, wants = blobs * 6, wants * 6
    _check(dev, blobs, wants)
