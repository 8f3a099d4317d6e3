from deepreduce_b200.grace import TopKCompressor  # noqa: F401
