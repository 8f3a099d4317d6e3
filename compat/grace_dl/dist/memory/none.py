from deepreduce_b200.grace import NoneMemory  # noqa: F401
