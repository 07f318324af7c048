"""sourmash_b200 -- B200-native FracMinHash sketching and sorted-hash intersection.

Drop-in for the two hot paths of sourmash (include/sourmash_b200.h):

    from sourmash_b200 import MinHash, SourmashSignature          # object model over the C ABI
    from sourmash_b200.compare import compare_all_pairs           # N x N on the GPU
    from sourmash_b200.index import LinearIndex, CounterGather    # search / prefetch / gather
    from sourmash_b200 import batch                               # raw batched entry points
"""
from . import batch  # noqa: F401
from ._lowlevel import ffi, lib  # noqa: F401
from .minhash import (FrozenMinHash, MinHash, get_minhash_default_seed, get_minhash_max_hash,  # noqa: F401
                      hash_murmur, translate_codon)
from .signature import (ComputeParameters, SourmashSignature, load_signatures,  # noqa: F401
                        load_signatures_from_json, save_signatures_to_json)

__version__ = "0.1.0"
