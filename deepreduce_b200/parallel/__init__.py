from .plan import BucketPlan, TensorPlan
from .engine import BucketEngine, engine_oracle, stats_from_slot

__all__ = ["BucketPlan", "TensorPlan", "BucketEngine", "engine_oracle", "stats_from_slot"]
