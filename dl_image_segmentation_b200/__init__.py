"""dl_image_segmentation_b200 — B200-native hot path of harry-gibson/dl_image_segmentation.

Same public names as the reference package (``dl_segmentation_utils/__init__.py:1-15``); the
arithmetic runs in hand-written sm_100a CUDA behind the C ABI of ``libb2chips.so``
(``include/b2chips.h``).  Importing this package does not need a GPU; calling anything does, and raises
if the library or the device is missing — there is no CPU fallback.
"""
from ._descartes_img_chips import (DLTileJobConfig, MaskedResult, SceneStack, SyntheticSceneSource, create_cloudmasked_s2_array,  # noqa: F401
                                   create_chips_for_tile, create_img_array_for_tile, create_label_array_for_tile, median_composite,
                                   nearest_date_mosaic, read_geojson_layer, stack_products_for_tile)
from ._tfrecord_image_translation import (convert_to_example, featuretemplate_bytestring_imagechip,  # noqa: F401
                                          featuretemplate_ndarray_imagechip, parse_8bit_array_proto,
                                          parse_encoded_gdal_proto_eager, parse_encoded_gdal_proto_wrapped,
                                          parse_encoded_rgb_img_proto, parse_encoded_shard, parse_higher_dtype_array_proto)
from ._tfrecord_image_translation import iter_parse_encoded_shards  # noqa: F401,E402

from ._geotiff import encode_geotiffs, write_chip_pair  # noqa: F401

__version__ = "0.1.0"
from ._img_to_tf_mp import process_dataset_mp as images_to_tfrecords_mp  # noqa: F401,E402
from ._img_to_tf_threaded import process_dataset_multithreaded as images_to_tfrecords_mt  # noqa: F401,E402
