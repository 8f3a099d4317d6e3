"""Validated, frozen view of the GRACE / DeepReduce ``params`` dict.

The reference threads one mutable dict through every layer and also uses it as a
side channel (``params['dense_tensor']`` pytorch/deepreduce.py:117,
``params['hash_table']`` :33,44; the TF half adds derived keys ``m, k, N, K,
X_train, num_of_segments`` tensorflow/deepreduce.py:312,385,392,412,454,481).
Here the dict stays the front door (API compatibility: reference README.md:30-48,
run_deepreduce.sh ``--grace_config``), but it is parsed once into
:class:`DeepReduceConfig` — typed, range-checked, immutable — and nothing is ever
written back into the user's dict.  Unknown keys are reported (typos such as
``'compres_ratio'`` silently fall back to defaults in the reference).
"""
from __future__ import annotations

import warnings
from dataclasses import asdict, dataclass
from typing import Optional

COMPRESSORS = ("none", "topk", "threshold", "randomk")
OUT_OF_SCOPE_COMPRESSORS = ("SKCompressCPU", "SKCompressGPU", "sketch")        # comparison baselines of a GRACE fork
MEMORIES = ("none", "residual")
COMMUNICATORS = ("allgather", "allreduce")
MODES = (None, "value", "index", "both")
POLICIES = ("leftmost", "leftmostK", "random", "randomK", "p0", "policy_zero", "conflict_sets", "p2")

# keys the PyTorch side reads (reference pytorch/deepreduce.py:36,57,106,384-385,512-513,857-858) + the ones this
# framework adds; the TF-side keys are handled by tf_compat and only listed so they do not trigger the typo warning
KNOWN_KEYS = frozenset({
    "compressor", "memory", "communicator", "compress_ratio", "threshold", "deepreduce", "value", "index", "fpr",
    "policy", "sort", "poly_degree", "quantum_num", "bucket_size", "micro-benchmark", "world_size", "average",
    "beta", "gamma", "seed", "code", "hint", "min_numel", "dense_tensor", "hash_table", "split_numel", "pack_mapping",
    "qsgd_seed", "gzip_level", "dexp_min_numel", "overlap_grid", "capacity_ratio", "calibrate_partition",
    # TF-side (tensorflow/deepreduce.py:34-36,57-59,282,307-343,361-369,458-490)
    "use_memory", "horovod_size", "bloom_fpr", "bloom_on", "threshold_val", "bloom_false_positives_aware",
    "bloom_policy", "bloom_logs_path", "gradient_id", "bloom_verbosity_frequency", "bloom_verbosity", "mem_mode",
    "suffix", "model_name", "approximation_mode", "polynomial_degree", "tensor_name", "step", "rank",
})


class ConfigError(ValueError):
    pass


@dataclass(frozen=True)
class DeepReduceConfig:
    compressor: str = "none"
    memory: str = "none"
    communicator: str = "allreduce"
    compress_ratio: float = 0.01
    threshold: float = 0.0
    deepreduce: Optional[str] = None
    value: str = "polyfit"
    index: str = "bloom"
    fpr: Optional[float] = None
    policy: str = "leftmost"
    poly_degree: int = 5
    quantum_num: int = 127
    bucket_size: int = 512
    micro_benchmark: bool = False
    average: bool = True
    beta: float = 1.0
    gamma: float = 1.0
    world_size: Optional[int] = None
    hint: bool = True
    min_numel: int = 1000

    # ------------------------------------------------------------------
    @classmethod
    def from_params(cls, params: dict, *, strict: bool = False) -> "DeepReduceConfig":
        """Parse + validate.  ``strict`` turns the unknown-key warning into an error."""
        if not isinstance(params, dict):
            raise ConfigError(f"params must be a dict (got {type(params).__name__})")
        unknown = sorted(k for k in params if k not in KNOWN_KEYS)
        if unknown:
            msg = f"unknown params key(s) {unknown}; known keys: {sorted(KNOWN_KEYS)}"
            if strict:
                raise ConfigError(msg)
            warnings.warn(msg, stacklevel=3)
        g = params.get
        comp = g("compressor", "none") or "none"
        if comp in OUT_OF_SCOPE_COMPRESSORS:
            raise NotImplementedError(
                f"compressor '{comp}' is a comparison baseline from a GRACE fork and is out of scope (SURVEY §2.5)")
        cfg = cls(
            compressor=comp, memory=g("memory", "none") or "none", communicator=g("communicator", "allreduce"),
            compress_ratio=float(g("compress_ratio", 0.01)), threshold=float(g("threshold", 0.0)),
            deepreduce=g("deepreduce", None) or None, value=g("value", "polyfit"), index=g("index", "bloom"),
            fpr=None if g("fpr", None) is None else float(g("fpr")), policy=g("policy", "leftmost"),
            poly_degree=int(g("poly_degree", 5)), quantum_num=int(g("quantum_num", 127)),
            bucket_size=int(g("bucket_size", 512)), micro_benchmark=bool(g("micro-benchmark", False)),
            average=bool(g("average", True)), beta=float(g("beta", 1.0)), gamma=float(g("gamma", 1.0)),
            world_size=None if g("world_size", None) is None else int(g("world_size")),
            hint=bool(g("hint", True)), min_numel=int(g("min_numel", 1000)))
        cfg.validate()
        return cfg

    def validate(self) -> None:
        from .codecs import compressor as registry

        def need(cond, msg):
            if not cond:
                raise ConfigError(msg)

        need(self.compressor in COMPRESSORS, f"'compressor' must be one of {COMPRESSORS} (got {self.compressor!r})")
        need(self.memory in MEMORIES, f"'memory' must be one of {MEMORIES} (got {self.memory!r})")
        need(self.communicator in COMMUNICATORS,
             f"'communicator' must be one of {COMMUNICATORS} (got {self.communicator!r})")
        need(self.deepreduce in MODES, f"'deepreduce' must be one of {MODES} (got {self.deepreduce!r})")
        need(0.0 < self.compress_ratio <= 1.0, f"'compress_ratio' must be in (0, 1] (got {self.compress_ratio})")
        need(self.threshold >= 0.0, f"'threshold' must be >= 0 (got {self.threshold})")
        need(self.fpr is None or 0.0 < self.fpr < 1.0, f"'fpr' must be in (0, 1) (got {self.fpr})")
        need(self.policy in POLICIES, f"'policy' must be one of {POLICIES} (got {self.policy!r})")
        need(1 <= self.poly_degree <= 7, f"'poly_degree' must be in [1, 7] (got {self.poly_degree})")
        need(1 <= self.quantum_num <= 32767, f"'quantum_num' must be in [1, 32767] (got {self.quantum_num})")
        need(self.bucket_size >= 1, f"'bucket_size' must be >= 1 (got {self.bucket_size})")
        need(self.world_size is None or self.world_size >= 1, f"'world_size' must be >= 1 (got {self.world_size})")
        if self.deepreduce in ("value", "both"):
            need(self.value in registry, f"unknown value codec {self.value!r}; registered: {sorted(registry)}")
        if self.deepreduce in ("index", "both"):
            need(self.index in registry, f"unknown index codec {self.index!r}; registered: {sorted(registry)}")
        if self.deepreduce is not None:
            need(self.compressor != "none", "'deepreduce' needs a sparsifier: set 'compressor' to topk/threshold/randomk")
            need(self.communicator == "allgather",
                 "sparse payloads differ per rank: 'deepreduce' requires 'communicator': 'allgather'")
        if self.compressor in ("topk", "threshold"):         # randomk with a shared seed is all-reducible, like GRACE
            need(self.communicator == "allgather",
                 f"compressor {self.compressor!r} produces per-rank index sets: use 'communicator': 'allgather'")

    def to_params(self) -> dict:
        """Back to the dict form the codecs/wrappers take (a fresh dict; the 'micro-benchmark' key keeps its dash)."""
        d = asdict(self)
        d["micro-benchmark"] = d.pop("micro_benchmark")
        return {k: v for k, v in d.items() if v is not None}


def validate_params(params: dict, *, strict: bool = False) -> DeepReduceConfig:
    return DeepReduceConfig.from_params(params, strict=strict)
