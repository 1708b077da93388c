"""B200-native acoustic front-end behind the feature-extraction boundary of
30stomercury/Automatic-Speech-Recognition (preprocess.py::process_audios and
utils/augmentation.py).  CUDA only: importing the compute API without the built
library raises; there is no CPU fallback."""
from . import tables, synth, sharding, audio_io            # noqa: F401  (host-only helpers)
from ._lib import FrontendLibraryError, library_path       # noqa: F401
from .frontend import Frontend, FrontendConfig, pack_pcm, num_frames   # noqa: F401
from .preprocess import process_audios, process_pcm, process_libri_feats, to_object_array  # noqa: F401
from .augmentation import SpeedAugmentation, VolumeAugmentation        # noqa: F401
from . import speechpy_shim                                            # noqa: F401
from . import tfrecord                                                 # noqa: F401  (create_tfrecord.py drop-in)
from .tfrecord import create_tfrecords                                 # noqa: F401
from . import bucketing                                                # noqa: F401  (tfrecord_data_loader.py batching)

__all__ = ["Frontend", "FrontendConfig", "pack_pcm", "num_frames", "process_audios", "process_pcm",
           "process_libri_feats", "to_object_array", "SpeedAugmentation", "VolumeAugmentation",
           "tables", "synth", "sharding", "audio_io", "speechpy_shim", "tfrecord", "create_tfrecords", "bucketing", "FrontendLibraryError", "library_path"]
