from .metrics import METRICS, Metrics, log_compressor, log_values, relative_volume

__all__ = ["METRICS", "Metrics", "log_compressor", "log_values", "relative_volume"]
