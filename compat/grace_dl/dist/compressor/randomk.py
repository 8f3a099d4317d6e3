from deepreduce_b200.grace import RandomKCompressor  # noqa: F401
