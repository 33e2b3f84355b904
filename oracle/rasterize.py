"""Label rasterisation (oracle; test infrastructure only).

Restates what ``create_label_array_for_tile`` (``_descartes_img_chips.py:633-689``) asks of GDAL: a ``(S, S)`` uint8 raster
filled with ``background_value`` (``:662-666``), then ``gdal.RasterizeLayer(..., options=['ALL_TOUCHED=TRUE'])`` with either the
features' ``ATTRIBUTE`` (``:684-685``) or ``burn_values=[1]`` (``:686-687``), features burnt in layer order so that a pixel shared
by several polygons keeps the value of the LAST one (the reference's own comment, ``:676-683``).

GDAL is not installable here and the reference pins no version: **parity unpinned** for this row.  The
arithmetic below restates GDAL's public algorithm (alg/llrasterize.cpp) as documented in its source comments:

* geometry -> pixel/line space through the inverse geotransform (north-up tiles: ``px = (x - gt0) / gt1``, ``py = (y - gt3) / gt5``,
  evaluated as GDAL does: ``inv0 + x * inv1`` with ``inv1 = 1 / gt1``, ``inv0 = -gt0 / gt1``);
* ``GDALdllImageFilledPolygon``: for every scanline the pixel-centre line ``y + 0.5`` is intersected with every edge of every
  ring (half-open in y: ``dy1 <= dy < dy2``), the crossings are rounded with ``floor(x + 0.5)``, sorted, and the spans between
  pairs ``[x_2i, x_2i+1 - 1]`` are burnt (even-odd rule, so holes and multi-polygons work); an edge lying exactly on the centre
  line burns ``floor(x_lo + 0.5) .. floor(x_hi + 0.5) - 1`` when it runs right to left;
* ``ALL_TOUCHED``: ``GDALdllImageLineAllTouched`` walks every edge pixel by pixel and burns each pixel it passes through:
  near-vertical edges (same pixel column, or |dx| < 0.01) burn the column ``floor(x_end)`` from ``floor(y_lo)`` to
  ``floor(y_hi)``; near-horizontal ones likewise along a row; all others are clipped to the raster and stepped from one
  pixel boundary to the next in x, or to the next scanline when the step would cross it (with GDAL's 1e-9 nudge).

``tests/test_rasterize.py`` pins this restatement from the outside on the one property ALL_TOUCHED is defined by: away from
GDAL's 0.01 shortcuts, the burnt set is exactly the set of pixels whose closed square meets the polygon.
"""
import math

import numpy as np


def to_pixel_space(xy, geotrans):
    """(N,2) map coordinates -> (N,2) pixel / line coordinates for a north-up geotransform (gt2 == gt4 == 0)."""
    gt = [float(v) for v in geotrans]
    assert gt[2] == 0.0 and gt[4] == 0.0, "rotated geotransforms are out of scope"
    inv1, inv5 = 1.0 / gt[1], 1.0 / gt[5]
    inv0, inv3 = -gt[0] * inv1, -gt[3] * inv5
    xy = np.asarray(xy, dtype=np.float64).reshape(-1, 2)
    return np.stack([inv0 + xy[:, 0] * inv1, inv3 + xy[:, 1] * inv5], axis=1)


def _fill_polygon(out, rings, value):
    H, W = out.shape
    ys = np.concatenate([r[:, 1] for r in rings])
    miny, maxy = int(ys.min()), int(ys.max())
    miny, maxy = max(miny, 0), min(maxy, H - 1)
    for y in range(miny, maxy + 1):
        dy = y + 0.5
        ints = []
        for r in rings:
            n = len(r)
            # edges (ind1 -> ind2) with ind1 the previous vertex: GDAL walks i = 0..n-1 with ind1 = i-1 (wrapping to the last)
            for i in range(n):
                i1 = i - 1 if i > 0 else n - 1
                dy1, dy2 = float(r[i1, 1]), float(r[i, 1])
                if (dy1 < dy and dy2 < dy) or (dy1 > dy and dy2 > dy):
                    continue
                if dy1 < dy2:
                    dx1, dx2 = float(r[i1, 0]), float(r[i, 0])
                elif dy1 > dy2:
                    dy2, dy1 = float(r[i1, 1]), float(r[i, 1])
                    dx2, dx1 = float(r[i1, 0]), float(r[i, 0])
                else:                                           # the edge lies on the centre line
                    if r[i1, 0] > r[i, 0]:
                        hx1 = int(math.floor(float(r[i, 0]) + 0.5))
                        hx2 = int(math.floor(float(r[i1, 0]) + 0.5))
                        if not (hx1 > W - 1 or hx2 <= 0):
                            out[y, max(hx1, 0):min(hx2 - 1, W - 1) + 1] = value
                    continue
                if dy < dy2 and dy >= dy1:
                    inter = (dy - dy1) * (dx2 - dx1) / (dy2 - dy1) + dx1
                    ints.append(int(math.floor(inter + 0.5)))
        ints.sort()
        for i in range(0, len(ints) - 1, 2):
            if ints[i] <= W - 1 and ints[i + 1] > 0:
                out[y, max(ints[i], 0):min(ints[i + 1] - 1, W - 1) + 1] = value


def _line_all_touched(out, ring, value):
    H, W = out.shape
    n = len(ring)
    for i in range(1, n):
        x, y = float(ring[i - 1, 0]), float(ring[i - 1, 1])
        xe, ye = float(ring[i, 0]), float(ring[i, 1])
        if (y > H and ye > H) or (y < 0.0 and ye < 0.0) or (x > W and xe > W) or (x < 0.0 and xe < 0.0):
            continue
        if x > xe:
            x, xe, y, ye = xe, x, ye, y
        if math.floor(x) == math.floor(xe) or abs(x - xe) < 0.01:          # vertical
            if ye < y:
                y, ye = ye, y
            ix = int(math.floor(xe))
            iy, iye = int(math.floor(y)), int(math.floor(ye))
            if ix < 0 or ix >= W:
                continue
            iy, iye = max(iy, 0), min(iye, H - 1)
            if iy <= iye:
                out[iy:iye + 1, ix] = value
            continue
        if math.floor(y) == math.floor(ye) or abs(y - ye) < 0.01:          # horizontal
            ix, iy, ixe = int(math.floor(x)), int(math.floor(y)), int(math.floor(xe))
            if iy < 0 or iy >= H:
                continue
            ix, ixe = max(ix, 0), min(ixe, W - 1)
            if ix <= ixe:
                out[iy, ix:ixe + 1] = value
            continue
        slope = (ye - y) / (xe - x)
        if xe > W:
            ye -= (xe - W) * slope
            xe = float(W)
        if x < 0.0:
            y += (0.0 - x) * slope
            x = 0.0
        if ye > y:
            if y < 0.0:
                x += (0.0 - y) / slope
                y = 0.0
            if ye >= H:
                xe += (ye - H) / slope
        else:
            if y >= H:
                x += (H - y) / slope
                y = float(H)
            if ye < 0.0:
                xe -= (ye - 0.0) / slope
        while x >= 0.0 and x < xe:
            ix, iy = int(math.floor(x)), int(math.floor(y))
            if 0 <= iy < H and ix < W:
                out[iy, ix] = value
            step_x = math.floor(x + 1.0) - x
            step_y = step_x * slope
            if int(math.floor(y + step_y)) == iy:
                x += step_x
                y += step_y
            elif slope < 0:
                step_y = iy - y
                if step_y > -0.000000001:
                    step_y = -0.000000001
                step_x = step_y / slope
                x += step_x
                y += step_y
            else:
                step_y = (iy + 1) - y
                if step_y < 0.000000001:
                    step_y = 0.000000001
                step_x = step_y / slope
                x += step_x
                y += step_y


def rasterize(features, size, background_value=255, all_touched=True):
    """features: list of (rings, value); rings = list of (N,2) float64 arrays in PIXEL space, closed (first == last vertex).
    -> (H, W) uint8.  Features are burnt in order: the last one wins a shared pixel."""
    H, W = (size, size) if np.isscalar(size) else size
    out = np.full((H, W), background_value, np.uint8)
    for rings, value in features:
        rings = [np.asarray(r, dtype=np.float64).reshape(-1, 2) for r in rings]
        rings = [r for r in rings if len(r) >= 2]
        if not rings:
            continue
        # the scanline fill walks every vertex with its predecessor (wrapping): drop the duplicated closing vertex there
        open_rings = [r[:-1] if len(r) > 1 and r[0, 0] == r[-1, 0] and r[0, 1] == r[-1, 1] else r for r in rings]
        _fill_polygon(out, [r for r in open_rings if len(r) >= 1], np.uint8(value))
        if all_touched:
            for r in rings:
                _line_all_touched(out, r, np.uint8(value))
    return out


def create_label_array_for_tile(tilesize, pad, geotrans, layer, attrib_to_burn=None, background_value=255):
    """``create_label_array_for_tile`` (:633-689) on an in-memory layer: list of (rings in MAP coordinates, attributes dict)."""
    size = tilesize + 2 * pad                                   # :660
    feats = []
    for rings, attrs in layer:
        value = int(attrs[attrib_to_burn]) if attrib_to_burn else 1       # :684-687
        feats.append(([to_pixel_space(r, geotrans) for r in rings], value))
    return rasterize(feats, size, background_value, all_touched=True)
