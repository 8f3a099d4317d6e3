"""`import deepreduce` — the module name of the reference's pytorch/deepreduce.py, backed by deepreduce_b200."""
from deepreduce_b200 import (DeepReduce, IndexCompressor, SparseCompressor, ValueCompressor, compressor,  # noqa: F401
                             deepreduce_from_params, deepreduce_wrapper, register)
from deepreduce_b200.codecs.bloom import Bloom, Bloomfilter  # noqa: F401
from deepreduce_b200.codecs.bloom_cpu import BloomCPU  # noqa: F401
from deepreduce_b200.codecs.lossless import Gzip, Huffman  # noqa: F401
from deepreduce_b200.codecs.polyfit import (GetInputMatrix_Polynomial, LeastSquares, PolyFit, PolyFitCPU,  # noqa: F401
                                            RestoreValues, find_breaks, fit_curve, get_segments, restore_curve)
from deepreduce_b200.codecs.qsgd import QSGD  # noqa: F401
from deepreduce_b200.codecs.rle import RunLength  # noqa: F401
from deepreduce_b200.spec import get_BFconfig  # noqa: F401
