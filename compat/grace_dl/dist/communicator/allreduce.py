from deepreduce_b200.grace import Allreduce  # noqa: F401
