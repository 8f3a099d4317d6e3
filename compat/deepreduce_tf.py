"""The classes of the reference's tensorflow/deepreduce.py (Compressor, Values_Approximation_Helper,
BloomFilterCompressor, DoubleExpCompressor, PolySegCompressor) — torch-hosted, see deepreduce_b200/tf_compat.py.
(Named deepreduce_tf so it can sit next to the PyTorch-side `deepreduce` module on one PYTHONPATH.)"""
from deepreduce_b200.tf_compat import (BloomFilterCompressor, Compressor, DoubleExpCompressor,  # noqa: F401
                                       PolySegCompressor, Values_Approximation_Helper)
