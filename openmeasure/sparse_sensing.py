"""`openmeasure.sparse_sensing` -> openmeasure_b200.sparse_sensing (same classes, same signatures)."""
from openmeasure_b200.sparse_sensing import ROM, SPR, SensorMatrix  # noqa: F401
