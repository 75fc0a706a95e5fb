"""API value types of the env path (mirror of the reference's core/types.py:19-102 field for field).

``SafetyConstraint`` keeps the reference signature ``(name, check_fn, penalty, critical, description)``;
constraints the CUDA kernels can evaluate themselves additionally carry a private ``_native`` descriptor
(built-in check / declarative bound). A constraint without one is an arbitrary Python callable: the env
evaluates it on the host copy of the pre-step state and hands the kernel a per-env violation bit.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np

Array = np.ndarray
StateArray = Array
ActionArray = Array
ObservationArray = Array
MetricsDict = Dict[str, Union[float, int, str]]
HyperparametersDict = Dict[str, Union[float, int, str, bool]]


class DatasetQuality(Enum):
    EXPERT = "expert"
    MEDIUM = "medium"
    MIXED = "mixed"
    RANDOM = "random"


@dataclass
class IndustrialState:
    observation: StateArray
    safety_metrics: Dict[str, float]
    timestamp: float
    system_status: str = "normal"
    confidence_score: float = 1.0
    uncertainty_bounds: Tuple[float, float] = (0.0, 0.0)
    anomaly_score: float = 0.0

    @property
    def is_safe(self) -> bool:
        return self.system_status in ("normal", "warning") and self.anomaly_score < 0.5 and self.confidence_score > 0.7

    def update_confidence(self, prediction_variance: float) -> None:
        self.confidence_score = max(0.0, min(1.0, 1.0 - prediction_variance))
        self.uncertainty_bounds = (-0.5 * prediction_variance, 0.5 * prediction_variance)


@dataclass
class SafetyConstraint:
    name: str
    check_fn: Callable[[StateArray, ActionArray], bool]
    penalty: float
    critical: bool = False
    description: str = ""
    # ("builtin", id) | ("bound", si, ai, coef, lo, hi) | None (host-evaluated callable)
    _native: Optional[tuple] = field(default=None, repr=False, compare=False)
    # batched envs: check_fn takes (states [n, S], actions [n, A]) and returns bool [n] (one call per step instead of n)
    vectorized: bool = field(default=False, compare=False)


@dataclass
class SafetyMetrics:
    constraints_satisfied: int
    total_constraints: int
    violation_count: int
    critical_violations: int
    safety_score: float
    adaptive_threshold: float = 0.95
    confidence_interval: Tuple[float, float] = (0.0, 1.0)
    violation_severity: Optional[Dict[str, float]] = None

    def __post_init__(self):
        if self.violation_severity is None:
            self.violation_severity = {}

    @property
    def satisfaction_rate(self) -> float:
        return 1.0 if self.total_constraints == 0 else self.constraints_satisfied / self.total_constraints

    @property
    def adaptive_safety_score(self) -> float:
        penalty = abs(self.confidence_interval[1] - self.confidence_interval[0]) * 0.1
        return max(0.0, self.safety_score - penalty)

    def update_adaptive_threshold(self, performance_history: List[float]) -> None:
        if len(performance_history) >= 10:
            recent = performance_history[-10:]
            self.adaptive_threshold = max(0.8, min(0.99, float(np.mean(recent) - 2 * np.std(recent))))


_POPCOUNT8 = np.array([bin(i).count("1") for i in range(256)], np.int64)


class BatchedSafetyMetrics:
    """``SafetyMetrics`` for N envs at once: same field names, array values, computed lazily from the
    per-env violation bit mask the kernel wrote (bit k = constraint k violated on the pre-step state)."""

    def __init__(self, violation_mask: Array, total_constraints: int, critical_bits: int):
        self.violation_mask = violation_mask
        self.total_constraints = int(total_constraints)
        self._critical_bits = np.uint8(critical_bits)

    @property
    def violation_count(self) -> Array:
        return _POPCOUNT8[self.violation_mask]

    @property
    def critical_violations(self) -> Array:
        return _POPCOUNT8[self.violation_mask & self._critical_bits]

    @property
    def constraints_satisfied(self) -> Array:
        return self.total_constraints - self.violation_count

    @property
    def safety_score(self) -> Array:
        if self.total_constraints == 0:
            return np.ones(self.violation_mask.shape, np.float64)
        return self.constraints_satisfied / self.total_constraints

    satisfaction_rate = safety_score

    def __len__(self):
        return len(self.violation_mask)

    def __getitem__(self, i: int) -> SafetyMetrics:
        nv = int(self.violation_count[i])
        return SafetyMetrics(self.total_constraints - nv, self.total_constraints, nv, int(self.critical_violations[i]),
                             float(self.safety_score[i]))
