from deepreduce_b200.grace import ThresholdCompressor  # noqa: F401
