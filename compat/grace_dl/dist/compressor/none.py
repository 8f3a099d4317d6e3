from deepreduce_b200.grace import NoneCompressor  # noqa: F401
