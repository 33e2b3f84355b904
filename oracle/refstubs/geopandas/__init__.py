"""`geopandas` placeholder: the reference imports it at module level (`_descartes_img_chips.py:1`); only the tile
discovery (`:387-457`, out of scope: needs the Descartes Labs service) calls into it."""
__version__ = "0.0-refstub"
