from .plan import BucketPlan, TensorPlan
from .engine import BucketEngine, engine_oracle

__all__ = ["BucketPlan", "TensorPlan", "BucketEngine", "engine_oracle"]
