"""`descarteslabs` stand-in: an in-memory scene catalogue with the client calls the reference makes (test infrastructure).

Reference call sites (`_descartes_img_chips.py`): `dl.scenes.search(aoi=, products=, start_datetime=, end_datetime=,
query=)` `:512,552,560,612`; `dl.properties.cloud_fraction < x` `:610`; `SceneCollection.stack(bands, ctx, processing_level=,
bands_axis=, data_type=)` `:557,561`; `.sorted(key, reverse=True)` `:623`; `.mosaic(bands=, ctx=, bands_axis=, processing_level=)`
`:513,626`; `scene.properties['date']` `:466`.

Documented service behaviour restated here (the client itself is not vendored):
  * search keeps scenes with `start_datetime <= acquired < end_datetime` and those satisfying the query expression,
    in acquisition order;
  * `stack` returns a masked array (scene, y, x, band); `mosaic` paints the scenes in collection order, "where multiple
    scenes overlap, only data from the scene that comes last in the SceneCollection is used", masked pixels never
    overwrite; an empty collection raises ValueError.
Register scenes with `catalog.clear()` / `catalog.add(product, Scene(...))` before calling the reference.
"""
import datetime as _dt
import types as _types

import numpy as _np

__version__ = "0.0-refstub"


class Scene:
    def __init__(self, date, bands, mask=None, cloud_fraction=None):
        """bands: {name: (H,W) array}; mask: (H,W) bool, True = no data in this scene."""
        self.properties = {"date": date if isinstance(date, _dt.datetime) else _dt.datetime.combine(date, _dt.time())}
        if cloud_fraction is not None:
            self.properties["cloud_fraction"] = float(cloud_fraction)
        self._bands, self._mask = bands, mask

    def ndarray(self, bands, ctx=None, bands_axis=-1, data_type=None, **kw):
        names = bands.split(" ") if isinstance(bands, str) else list(bands)
        arr = _np.stack([_np.asarray(self._bands[n]) for n in names], axis=-1)
        if data_type == "Byte":
            arr = arr.astype(_np.uint8)
        m = _np.zeros(arr.shape, bool) if self._mask is None else _np.repeat(_np.asarray(self._mask, bool)[..., None], len(names), -1)
        assert bands_axis == -1
        return _np.ma.MaskedArray(arr, mask=m)


class SceneCollection(list):
    def sorted(self, *predicates, reverse=False):
        key = predicates[0] if len(predicates) == 1 else (lambda s: tuple(p(s) for p in predicates))
        return SceneCollection(sorted(self, key=key, reverse=reverse))          # Python's sort is stable, as the client's

    def stack(self, bands, ctx, flatten=None, mask_nodata=True, mask_alpha=None, bands_axis=1, raster_info=False,
              resampler="near", processing_level=None, scaling=None, data_type=None, max_workers=None):
        if len(self) == 0:
            raise ValueError("This SceneCollection is empty")
        arrs = [s.ndarray(bands, ctx, bands_axis=bands_axis, data_type=data_type) for s in self]
        return _np.ma.stack(arrs, axis=0)

    def mosaic(self, bands, ctx, mask_nodata=True, mask_alpha=None, bands_axis=0, resampler="near", processing_level=None,
               scaling=None, data_type=None, raster_info=False):
        if len(self) == 0:
            raise ValueError("This SceneCollection is empty")
        out = None
        for s in self:                                          # painter's order: last in the collection wins
            a = s.ndarray(bands, ctx, bands_axis=bands_axis, data_type=data_type)
            if out is None:
                out = _np.ma.MaskedArray(_np.zeros(a.shape, a.dtype), mask=_np.ones(a.shape, bool))
            ok = ~_np.ma.getmaskarray(a)
            out.data[ok] = a.data[ok]
            out.mask[ok] = False
        return out


class _Expr:
    def __init__(self, fn):
        self.fn = fn


class _Property:
    def __init__(self, name):
        self.name = name

    def __lt__(self, v):
        return _Expr(lambda s: s.properties.get(self.name) is not None and s.properties[self.name] < v)

    def __le__(self, v):
        return _Expr(lambda s: s.properties.get(self.name) is not None and s.properties[self.name] <= v)

    def __gt__(self, v):
        return _Expr(lambda s: s.properties.get(self.name) is not None and s.properties[self.name] > v)

    def __ge__(self, v):
        return _Expr(lambda s: s.properties.get(self.name) is not None and s.properties[self.name] >= v)


class _Properties:
    def __getattr__(self, name):
        return _Property(name)


properties = _Properties()


class _Catalog:
    def __init__(self):
        self.products = {}

    def clear(self):
        self.products = {}

    def add(self, product, scene):
        self.products.setdefault(product, []).append(scene)


catalog = _Catalog()


def _parse_dt(s):
    if s is None:
        return None
    if isinstance(s, _dt.datetime):
        return s
    if isinstance(s, _dt.date):
        return _dt.datetime.combine(s, _dt.time())
    d = _dt.datetime.fromisoformat(s)
    return d


def _search(aoi, products=None, start_datetime=None, end_datetime=None, cloud_fraction=None, query=None, limit=100,
            sort_field=None, sort_order="asc", **kw):
    lo, hi = _parse_dt(start_datetime), _parse_dt(end_datetime)
    scenes = []
    for s in sorted(catalog.products.get(products, []), key=lambda s: s.properties["date"]):     # acquisition order (stable)
        d = s.properties["date"]
        if lo is not None and d < lo:
            continue
        if hi is not None and d >= hi:
            continue
        if query is not None and not query.fn(s):
            continue
        scenes.append(s)
    return SceneCollection(scenes), aoi


scenes = _types.SimpleNamespace(search=_search, SceneCollection=SceneCollection, Scene=Scene, DLTile=object)
