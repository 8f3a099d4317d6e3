from deepreduce_b200.grace import Allgather  # noqa: F401
