from .plan import BucketPlan, TensorPlan
from .engine import BucketEngine, decode_slot_oracle, engine_oracle, stats_from_slot
from .ddp import DeepReduceDDP

__all__ = ["BucketPlan", "TensorPlan", "BucketEngine", "DeepReduceDDP", "decode_slot_oracle", "engine_oracle", "stats_from_slot"]
