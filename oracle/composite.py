"""Per-pixel compositors over a time stack (oracle; test infrastructure only).

``median_composite``  = ``_descartes_img_chips.py:562-567`` verbatim: repeat the (T,H,W,1) validity
band over the B bands, ``np.ma.masked_where(mask == 0, stack)``, ``np.ma.median(axis=0)``.  NumPy is the
library the reference itself calls, so this *is* the reference arithmetic (float64 result, masked
where no scene is valid).

``nearest_date_mosaic`` = ``create_img_array_for_tile`` ``:571-629`` with the Descartes Labs service
replaced by its documented behaviour: search filter ``start_datetime <= date < end_datetime``
(``:603-606``), ``cloud_fraction < max_cloud_fraction`` (``:607-610``), no scenes -> ``None``
(``:614-615``), scenes stable-sorted by ``abs(date - reference_date)`` descending (``:461-469,622-623``),
then painted in that order so the last valid scene wins each pixel (``:619-621,626``).
``stack_products`` = ``np.dstack`` of per-product mosaics (``:516``).
"""
import numpy as np


def median_composite(stack, valid, nodata_mask=None):
    """stack (T,H,W,B); valid (T,H,W) or (T,H,W,1), 0 = cloudy/invalid. -> np.ma.MaskedArray (H,W,B) float64."""
    stack = np.asarray(stack)
    valid = np.asarray(valid)
    if valid.ndim == 3:
        valid = valid[..., None]
    n_bands = stack.shape[-1]
    rep = np.repeat(valid, repeats=n_bands, axis=-1)            # :563
    base = stack if nodata_mask is None else np.ma.masked_array(stack, mask=np.asarray(nodata_mask) != 0)
    data = np.ma.masked_where(rep == 0, base)                   # :565 (ORs with an existing mask)
    return np.ma.median(data, axis=0)                           # :567


def scene_order(scene_day, scene_cf, ref_day, min_day=None, max_day=None, max_cf=None):
    """Indices of the scenes that survive the search filter, in painting order (first painted first)."""
    scene_day = [int(d) for d in scene_day]
    keep = []
    for i, d in enumerate(scene_day):
        if min_day is not None and d < min_day:
            continue
        if max_day is not None and d >= max_day:
            continue
        if max_cf is not None and not (float(scene_cf[i]) < max_cf):
            continue
        keep.append(i)
    # Python's sorted(..., reverse=True) is stable: ties keep search order
    return sorted(keep, key=lambda i: abs(scene_day[i] - ref_day), reverse=True)


def nearest_date_mosaic(stack, valid, scene_day, scene_cf, ref_day, min_day=None, max_day=None, max_cf=None):
    """-> (out (H,W,B) same dtype, mask (H,W) bool [True = no valid scene], src (H,W) int16) or None."""
    stack = np.asarray(stack)
    valid = np.asarray(valid)
    if valid.ndim == 4:
        valid = valid[..., 0]
    order = scene_order(scene_day, scene_cf, ref_day, min_day, max_day, max_cf)
    if not order:
        return None
    T, H, W, B = stack.shape
    out = np.zeros((H, W, B), dtype=stack.dtype)
    src = np.full((H, W), -1, dtype=np.int16)
    for t in order:                                             # painter's loop: later overwrites
        m = valid[t] != 0
        out[m] = stack[t][m]
        src[m] = t
    return out, src < 0, src


def stack_products(arrays):
    return np.dstack(arrays)                                    # :516


def search_filter(scene_dates, min_date=None, max_date=None):
    """Indices kept by ``dl.scenes.search(start_datetime=min_date, end_datetime=max_date)`` (``:549-552``): acquired in
    ``[min_date, max_date)``.  Dates are ``datetime.date`` / ``datetime`` objects or plain day numbers."""
    import datetime as _dt

    def day(d):
        if isinstance(d, _dt.datetime):
            return d.date().toordinal()
        return d.toordinal() if isinstance(d, _dt.date) else int(d)
    lo = None if min_date is None else day(min_date)
    hi = None if max_date is None else day(max_date)
    return [i for i, d in enumerate(scene_dates) if (lo is None or day(d) >= lo) and (hi is None or day(d) < hi)]


def create_cloudmasked_s2_array(scene_dates, stack, valid_cloudfree, nodata_mask=None, min_date=None, max_date=None):
    """``create_cloudmasked_s2_array`` ``:521-568`` on an in-memory catalogue: date search, ``None`` when nothing is left
    (``:554-555``), then the masked median of the surviving scenes."""
    keep = search_filter(scene_dates, min_date, max_date)
    if not keep:
        return None
    nd = None if nodata_mask is None else np.asarray(nodata_mask)[keep]
    return median_composite(np.asarray(stack)[keep], np.asarray(valid_cloudfree)[keep], nd)
