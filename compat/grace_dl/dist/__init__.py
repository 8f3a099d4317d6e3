from deepreduce_b200.grace import Communicator, Compressor, Memory  # noqa: F401
