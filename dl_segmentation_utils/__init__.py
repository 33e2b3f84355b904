"""Drop-in alias: ``import dl_segmentation_utils`` resolves to the B200 implementation.

Mirrors the re-exports of the reference's ``dl_segmentation_utils/__init__.py:1-15`` that are on the hot path
(the Descartes Labs catalog / OGR configuration classes are out of scope)."""
from dl_image_segmentation_b200 import *  # noqa: F401,F403
from dl_image_segmentation_b200 import (_descartes_img_chips, _img_to_tf_mp, _img_to_tf_threaded,  # noqa: F401
                                        _tfrecord_image_translation)
from dl_image_segmentation_b200 import images_to_tfrecords_mp, images_to_tfrecords_mt  # noqa: F401
