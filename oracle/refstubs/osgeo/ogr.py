"""osgeo.ogr placeholder: `ogr.Open(path).GetLayerByIndex(i)` hands back the polygons registered for `path`."""
datasets = {}          # path -> list of layers; layer = list of (polygon rings [[(x,y),...], ...], {attr: value})


class _DS:
    def __init__(self, layers):
        self.layers = layers

    def GetLayerByIndex(self, i):
        return self.layers[i]


def Open(path):
    if path not in datasets:
        return None
    return _DS(datasets[path])
