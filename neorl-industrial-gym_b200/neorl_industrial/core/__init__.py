from .types import DatasetQuality, IndustrialState, SafetyConstraint, SafetyMetrics, BatchedSafetyMetrics

__all__ = ["DatasetQuality", "IndustrialState", "SafetyConstraint", "SafetyMetrics", "BatchedSafetyMetrics"]
