"""osgeo.gdal stand-in: constants + a recording MEM / GTiff dataset.

`GetDriverByName(name).Create(path, xsize, ysize, bands, type, options=)` returns a dataset that records
`SetProjection`, `SetGeoTransform`, per-band `WriteArray` / `SetNoDataValue`; the finished datasets are kept in
`created` (by path) so a test can see exactly what the reference asked GDAL to write.  `RasterizeLayer` delegates to
`rasterize_hook` (set by the test to the rasteriser under test's independent checker)."""
import numpy as np

GDT_Byte, GDT_UInt16, GDT_Int16, GDT_UInt32, GDT_Int32, GDT_Float32, GDT_Float64 = 1, 2, 3, 4, 5, 6, 7
_NP = {1: np.uint8, 2: np.uint16, 3: np.int16, 4: np.uint32, 5: np.int32, 6: np.float32, 7: np.float64}
created = {}
rasterize_hook = None


class _Band:
    def __init__(self, ds, i):
        self.ds, self.i, self.nodata = ds, i, None

    def WriteArray(self, arr):
        a = np.asarray(arr)
        assert a.shape == (self.ds.ysize, self.ds.xsize), (a.shape, self.ds.ysize, self.ds.xsize)
        self.ds.data[self.i] = a.astype(self.ds.dtype)          # GDAL converts to the band type on write

    def SetNoDataValue(self, v):
        self.nodata = v
        self.ds.nodata[self.i] = v

    def ReadAsArray(self):
        return self.ds.data[self.i].copy()


class _Dataset:
    def __init__(self, driver, path, xsize, ysize, bands, gdt, options):
        self.driver, self.path, self.xsize, self.ysize, self.bands, self.gdt = driver, path, xsize, ysize, bands, gdt
        self.options = list(options or [])
        self.dtype = _NP[gdt]
        self.data = np.zeros((bands, ysize, xsize), self.dtype)
        self.nodata = [None] * bands
        self.projection, self.geotransform = "", (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)

    def SetProjection(self, wkt):
        self.projection = wkt

    def SetGeoTransform(self, gt):
        self.geotransform = tuple(gt)

    def GetRasterBand(self, i):
        return _Band(self, i - 1)

    def FlushCache(self):
        pass

    def ReadAsArray(self):
        return self.data[0].copy() if self.bands == 1 else self.data.copy()


class _Driver:
    def __init__(self, name):
        self.name = name

    def Create(self, path, xsize, ysize, bands=1, eType=GDT_Byte, options=None):
        ds = _Dataset(self.name, path, xsize, ysize, bands, eType, options)
        created[path] = ds
        return ds


def GetDriverByName(name):
    return _Driver(name)


def RasterizeLayer(ds, bands, layer, burn_values=None, options=None):
    if rasterize_hook is None:
        raise NotImplementedError("refstub gdal.RasterizeLayer: set osgeo.gdal.rasterize_hook")
    rasterize_hook(ds, bands, layer, burn_values, list(options or []))
    return 0
