from deepreduce_b200.grace import grace_from_params, tensor_bits  # noqa: F401
