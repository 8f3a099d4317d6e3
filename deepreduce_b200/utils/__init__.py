from .metrics import (METRICS, Metrics, StepLogger, bitstream_str, log_bitstream_compressor,
                      log_bitstream_decompressor, log_compressor, log_decompressor, log_integer_codec, log_values,
                      relative_volume)

__all__ = ["METRICS", "Metrics", "StepLogger", "bitstream_str", "log_bitstream_compressor",
           "log_bitstream_decompressor", "log_compressor", "log_decompressor", "log_integer_codec", "log_values",
           "relative_volume"]
