"""rasterio.plot.reshape_as_image / reshape_as_raster (pure axis moves)."""
import numpy as np


def reshape_as_image(arr):
    """(bands, rows, cols) -> (rows, cols, bands)"""
    return np.ma.transpose(arr, [1, 2, 0])


def reshape_as_raster(arr):
    """(rows, cols, bands) -> (bands, rows, cols)"""
    return np.transpose(arr, [2, 0, 1])
