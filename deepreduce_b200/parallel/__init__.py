from .plan import BucketPlan, TensorPlan
from .engine import BucketEngine, decode_slot_oracle, engine_oracle, stats_from_slot

__all__ = ["BucketPlan", "TensorPlan", "BucketEngine", "decode_slot_oracle", "engine_oracle", "stats_from_slot"]
