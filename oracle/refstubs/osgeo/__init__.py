"""`osgeo` stand-in (gdal / ogr): the type constants the reference reads at import (`_descartes_img_chips.py:853-875`)
and a recording dataset for `create_chips_for_tile` (`:781-797`, `:804-849`).  Test infrastructure; see ../README.md."""
from . import gdal, ogr  # noqa: F401
