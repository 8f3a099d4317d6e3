from deepreduce_b200.grace import ResidualMemory  # noqa: F401
