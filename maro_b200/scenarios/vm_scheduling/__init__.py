"""vm_scheduling scenario: host loader (data.py) and the reference's agent-facing types (common.py)."""
from .common import Action, AllocateAction, DecisionEvent, Latency, PostponeAction, VmCategory

__all__ = ["Action", "AllocateAction", "DecisionEvent", "Latency", "PostponeAction", "VmCategory"]
