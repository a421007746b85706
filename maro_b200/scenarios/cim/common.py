"""CIM agent-facing types with the reference's names and fields
(maro/simulator/scenarios/cim/common.py:18-150): ActionType, Action, ActionScope, DecisionEvent."""
from enum import Enum


class ActionType(Enum):
    LOAD = "load"
    DISCHARGE = "discharge"


class Action:
    summary_key = ["port_idx", "vessel_idx", "action_type", "quantity"]

    def __init__(self, vessel_idx: int, port_idx: int, quantity: int, action_type: ActionType):
        assert action_type is not None
        assert quantity >= 0
        self.vessel_idx = vessel_idx
        self.port_idx = port_idx
        self.quantity = quantity
        self.action_type = action_type

    def __repr__(self):
        return "%s {action_type: %r, port_idx: %r, vessel_idx: %r, quantity: %r}" % (
            self.__class__.__name__, str(self.action_type), self.port_idx, self.vessel_idx, self.quantity)


class ActionScope:
    def __init__(self, load: int, discharge: int):
        self.load = load
        self.discharge = discharge

    def __repr__(self):
        return "%s {load: %r, discharge: %r}" % (self.__class__.__name__, self.load, self.discharge)


class DecisionEvent:
    """Decision payload.  The reference evaluates ``action_scope`` / ``early_discharge`` lazily through callbacks;
    here they are the values the kernel wrote at the decision point (the state cannot change before the action)."""

    summary_key = ["tick", "port_idx", "vessel_idx", "snapshot_list", "action_scope", "early_discharge"]

    def __init__(self, tick, port_idx, vessel_idx, snapshot_list, action_scope, early_discharge):
        self.tick = tick
        self.port_idx = port_idx
        self.vessel_idx = vessel_idx
        self.snapshot_list = snapshot_list
        self.action_scope = action_scope
        self.early_discharge = early_discharge

    def __getstate__(self):
        return {"tick": self.tick, "port_idx": self.port_idx, "vessel_idx": self.vessel_idx,
                "action_scope": self.action_scope, "early_discharge": self.early_discharge}

    def __setstate__(self, state):
        self.tick = state["tick"]
        self.port_idx = state["port_idx"]
        self.vessel_idx = state["vessel_idx"]
        self.action_scope = state["action_scope"]
        self.early_discharge = state["early_discharge"]
        self.snapshot_list = None

    def __repr__(self):
        return "%s {port_idx: %r, vessel_idx: %r, action_scope: %r, early_discharge: %r}" % (
            self.__class__.__name__, self.port_idx, self.vessel_idx, self.action_scope, self.early_discharge)


def encode_action(action, out_row) -> None:
    """Write one Action (ours or the reference's — duck-typed) into an int32[4] row of the C ABI."""
    at = action.action_type
    name = getattr(at, "name", None) or str(at)
    out_row[0] = action.vessel_idx
    out_row[1] = action.port_idx
    out_row[2] = action.quantity
    out_row[3] = 1 if name.upper().endswith("DISCHARGE") else 0
