"""Label rasterisation (`create_label_array_for_tile`, `_descartes_img_chips.py:633-689`).

CPU: the oracle's restatement of GDAL's ALL_TOUCHED burn (oracle/rasterize.py) is pinned from the outside on the property
that defines ALL_TOUCHED — a pixel is burnt iff its closed square meets the polygon — computed independently with exact
segment / box clipping, and on hand-checked cases (last feature wins, holes, attribute values, clipping at the raster's edge).
GPU: the kernels reproduce the oracle bit for bit, through the drop-in `create_label_array_for_tile` and the C ABI.
"""
import json

import numpy as np
import pytest

from oracle import rasterize as orr


def _seg_hits_box(p, q, x0, y0, x1, y1):
    """Closed segment vs closed box (Liang-Barsky)."""
    dx, dy = q[0] - p[0], q[1] - p[1]
    t0, t1 = 0.0, 1.0
    for pp, qq in ((-dx, p[0] - x0), (dx, x1 - p[0]), (-dy, p[1] - y0), (dy, y1 - p[1])):
        if pp == 0:
            if qq < 0:
                return False
        else:
            r = qq / pp
            if pp < 0:
                if r > t1:
                    return False
                t0 = max(t0, r)
            else:
                if r < t0:
                    return False
                t1 = min(t1, r)
    return t0 <= t1


def _inside(pt, rings):
    c = False
    for r in rings:
        for i in range(len(r) - 1):
            a, b = r[i], r[i + 1]
            if (a[1] > pt[1]) != (b[1] > pt[1]) and pt[0] < a[0] + (pt[1] - a[1]) * (b[0] - a[0]) / (b[1] - a[1]):
                c = not c
    return c


def _exact_all_touched(rings, H, W):
    out = np.zeros((H, W), bool)
    for y in range(H):
        for x in range(W):
            out[y, x] = _inside((x + 0.5, y + 0.5), rings) or any(
                _seg_hits_box(r[i], r[i + 1], x, y, x + 1, y + 1) for r in rings for i in range(len(r) - 1))
    return out


def _star(rng, centre, rmin, rmax, k):
    ang = np.sort(rng.random(k)) * 2 * np.pi
    rad = rng.uniform(rmin, rmax, k)
    ring = np.stack([centre[0] + rad * np.cos(ang), centre[1] + rad * np.sin(ang)], 1)
    return np.vstack([ring, ring[:1]])


def _random_layer(rng, n, size, with_holes=True):
    feats = []
    for f in range(n):
        c = rng.uniform(-0.1 * size, 1.1 * size, 2)              # some polygons hang over the raster's edge
        shell = _star(rng, c, 0.03 * size, 0.35 * size, int(rng.integers(3, 12)))
        rings = [shell]
        if with_holes and f % 3 == 0:
            rings.append((c + (shell - c) * 0.45)[::-1].copy())
        if f % 5 == 4:                                           # a multi-polygon: a second shell elsewhere
            rings.append(_star(rng, rng.uniform(0, size, 2), 0.02 * size, 0.1 * size, 5))
        feats.append((rings, int(rng.integers(0, 256))))
    return feats


def test_oracle_burns_exactly_the_pixels_whose_square_meets_the_polygon():
    rng = np.random.default_rng(11)
    for it in range(40):
        S = 20
        rings = [_star(rng, rng.uniform(3, 17, 2), 1.5, 11, int(rng.integers(3, 9)))]
        if it % 3 == 0:
            c = rings[0][:-1].mean(0)
            rings.append((c + (rings[0] - c) * 0.4)[::-1].copy())
        d = np.abs(np.diff(np.vstack(rings), axis=0))
        if ((d[:, 0] < 0.02) | (d[:, 1] < 0.02)).any():          # GDAL's 0.01 shortcuts for near-axis-parallel edges
            continue
        got = orr.rasterize([(rings, 1)], S, 0) == 1
        assert np.array_equal(got, _exact_all_touched(rings, S, S)), it
        centre_only = orr.rasterize([(rings, 1)], S, 0, all_touched=False) == 1
        want_centre = np.array([[_inside((x + 0.5, y + 0.5), rings) for x in range(S)] for y in range(S)])
        assert np.array_equal(centre_only, want_centre), it


def test_oracle_hand_cases():
    sq = np.array([[2.0, 2.0], [6.0, 2.0], [6.0, 5.0], [2.0, 5.0], [2.0, 2.0]])      # exactly on pixel boundaries
    r = orr.rasterize([([sq], 7)], 8, 255)
    assert (r[2:5, 2:6] == 7).all() and (r == 7).sum() >= 12 and r[0, 0] == 255
    # last feature wins a shared pixel, whatever the values
    a = np.array([[1.2, 1.2], [4.8, 1.2], [4.8, 4.8], [1.2, 4.8], [1.2, 1.2]])
    b = a + 2.0
    r = orr.rasterize([([a], 200), ([b], 3)], 8, 0)
    assert r[4, 4] == 3 and r[1, 1] == 200
    r = orr.rasterize([([b], 3), ([a], 200)], 8, 0)
    assert r[4, 4] == 200
    # a hole is not burnt, except where ALL_TOUCHED catches its rim
    shell = np.array([[0.5, 0.5], [15.5, 0.5], [15.5, 15.5], [0.5, 15.5], [0.5, 0.5]])
    hole = np.array([[4.5, 4.5], [4.5, 11.5], [11.5, 11.5], [11.5, 4.5], [4.5, 4.5]])
    r = orr.rasterize([([shell, hole], 1)], 16, 0)
    assert r[8, 8] == 0 and r[4, 8] == 1 and r[2, 2] == 1
    # geotransform: a DLTile-style north-up tile
    gt = (500000.0, 10.0, 0.0, 4100000.0, 0.0, -10.0)
    px = orr.to_pixel_space(np.array([[500025.0, 4099975.0]]), gt)
    assert np.allclose(px, [[2.5, 2.5]])
    lab = orr.create_label_array_for_tile(12, 2, gt, [([np.array([[500020.0, 4099980.0], [500060.0, 4099980.0], [500060.0, 4099940.0],
                                                                 [500020.0, 4099940.0], [500020.0, 4099980.0]])], {"cls": 9})], "cls", 255)
    assert lab.shape == (16, 16) and lab.dtype == np.uint8 and lab[3, 3] == 9 and lab[10, 10] == 255


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda", 0)


@pytest.mark.gpu
def test_gpu_rasterize_is_bit_identical_with_the_oracle(dev):
    from dl_image_segmentation_b200 import ops
    rng = np.random.default_rng(21)
    for size, n in ((64, 12), (97, 40), ((50, 130), 25), (512, 60)):
        H, W = (size, size) if np.isscalar(size) else size
        feats = _random_layer(rng, n, min(H, W))
        for all_touched in (True, False):
            got = ops.rasterize_polygons(feats, size, 255, all_touched, device=dev).cpu().numpy()
            want = orr.rasterize(feats, size, 255, all_touched)
            assert got.dtype == np.uint8 and got.shape == (H, W)
            assert np.array_equal(got, want), (size, all_touched, int((got != want).sum()))
    # degenerate inputs: nothing to burn, a sliver, axis-parallel edges on pixel boundaries, everything outside
    assert (ops.rasterize_polygons([], 16, 7, device=dev).cpu().numpy() == 7).all()
    cases = [[([np.array([[3.0, 3.0], [9.0, 3.0], [9.0, 8.0], [3.0, 8.0], [3.0, 3.0]])], 1)],
             [([np.array([[1.5, 1.5], [14.5, 1.5000001], [1.5, 1.5]])], 2)],
             [([np.array([[-50.0, -40.0], [-10.0, -40.0], [-10.0, -5.0], [-50.0, -40.0]])], 3)],
             [([np.array([[-5.0, 4.3], [40.0, 7.9], [40.0, -3.0], [-5.0, 4.3]])], 4)],
             [([np.array([[2.0, 2.5], [12.0, 2.5], [12.0, 9.5], [2.0, 9.5], [2.0, 2.5]])], 5)]]     # edges ON the centre lines
    for feats in cases:
        got = ops.rasterize_polygons(feats, 16, 0, device=dev).cpu().numpy()
        assert np.array_equal(got, orr.rasterize(feats, 16, 0)), feats[0][1]


@pytest.mark.gpu
def test_gpu_create_label_array_for_tile_dropin(dev, tmp_path):
    """The reference's call (`create_chips_for_tile`, :775-777) on a GeoJSON layer in the tile's CRS, cfg3 tile geometry
    (448 px + 2 x 32 padding = 512): attribute burn and the default burn value 1."""
    import dl_image_segmentation_b200 as pkg

    class Tile:
        tilesize, pad = 448, 32
        geotrans = (499680.0, 10.0, 0.0, 5300360.0, 0.0, -10.0)
    rng = np.random.default_rng(5)
    feats = _random_layer(rng, 30, 512)
    gt = Tile.geotrans
    to_map = lambda r: np.stack([gt[0] + r[:, 0] * gt[1], gt[3] + r[:, 1] * gt[5]], 1)
    layer = [([to_map(r) for r in rings], {"landuse": v, "name": "f%d" % i}) for i, (rings, v) in enumerate(feats)]
    gj = {"type": "FeatureCollection", "features": [
        {"type": "Feature", "properties": props,
         "geometry": {"type": "Polygon", "coordinates": [r.tolist() for r in rings]} if len(rings) == 1 else
                     {"type": "MultiPolygon", "coordinates": [[r.tolist()] for r in rings]}} for rings, props in layer]}
    path = tmp_path / "labels.geojson"
    path.write_text(json.dumps(gj))
    for attrib in ("landuse", None):
        got = pkg.create_label_array_for_tile(Tile, str(path), attrib_to_burn=attrib, background_value=255)
        want = orr.create_label_array_for_tile(448, 32, gt, pkg.read_geojson_layer(str(path)), attrib, 255)
        assert got.is_cuda and tuple(got.shape) == (512, 512)
        assert np.array_equal(got.cpu().numpy(), want)
    assert (want != 255).any() and set(np.unique(want)) <= {1, 255}
