"""CIM agent-facing types with the reference's names and fields
(maro/simulator/scenarios/cim/common.py:18-150): ActionType, Action, ActionScope, DecisionEvent.

These are plain value holders here: the kernel has already evaluated everything a decision carries, so there are no
lazy callbacks into a business engine; reference ``Action`` objects are accepted as well (duck typing in
``encode_action``)."""
from enum import Enum


class ActionType(Enum):
    LOAD = "load"
    DISCHARGE = "discharge"


class _Record:
    """value holder: positional fields, ``ClassName {field: value, ...}`` repr in `_shown` order"""

    _fields: tuple = ()
    _shown: tuple = ()

    def _assign(self, values):
        self.__dict__.update(zip(self._fields, values))  # (one dict update: VectorEnv builds these per env per step)

    def __repr__(self):
        body = ", ".join(f"{k}: {self._show(k)!r}" for k in (self._shown or self._fields))
        return f"{type(self).__name__} {{{body}}}"

    def _show(self, key):
        return getattr(self, key)


class Action(_Record):
    summary_key = ["port_idx", "vessel_idx", "action_type", "quantity"]
    _fields = ("vessel_idx", "port_idx", "quantity", "action_type")
    _shown = ("action_type", "port_idx", "vessel_idx", "quantity")

    def __init__(self, vessel_idx: int, port_idx: int, quantity: int, action_type: ActionType):
        if action_type is None or quantity < 0:
            raise AssertionError("an Action needs an ActionType and a non-negative quantity")
        self._assign((vessel_idx, port_idx, quantity, action_type))

    def _show(self, key):
        return str(self.action_type) if key == "action_type" else getattr(self, key)


class ActionScope(_Record):
    _fields = ("load", "discharge")

    def __init__(self, load: int, discharge: int):
        self.load = load
        self.discharge = discharge


class DecisionEvent(_Record):
    """Decision payload.  The reference evaluates ``action_scope`` / ``early_discharge`` lazily through callbacks;
    here they are the values the kernel wrote at the decision point (the state cannot change before the action).
    Pickles without its ``snapshot_list`` (a handle on device memory), like the reference drops its callbacks."""

    summary_key = ["tick", "port_idx", "vessel_idx", "snapshot_list", "action_scope", "early_discharge"]
    _fields = ("tick", "port_idx", "vessel_idx", "snapshot_list", "action_scope", "early_discharge")
    _shown = ("port_idx", "vessel_idx", "action_scope", "early_discharge")

    def __init__(self, tick, port_idx, vessel_idx, snapshot_list, action_scope, early_discharge):
        self.tick = tick
        self.port_idx = port_idx
        self.vessel_idx = vessel_idx
        self.snapshot_list = snapshot_list
        self.action_scope = action_scope
        self.early_discharge = early_discharge

    def __getstate__(self):
        return {k: getattr(self, k) for k in self._fields if k != "snapshot_list"}

    def __setstate__(self, state):
        self.snapshot_list = None
        for k, v in state.items():
            setattr(self, k, v)


def action_row(action) -> tuple:
    """(vessel, port, quantity, type) of one Action (ours or the reference's — duck-typed) as plain ints"""
    t = action.action_type
    if t is ActionType.LOAD:
        return (action.vessel_idx, action.port_idx, action.quantity, 0)
    if t is ActionType.DISCHARGE:
        return (action.vessel_idx, action.port_idx, action.quantity, 1)
    kind = getattr(t, "name", None) or str(t)  # the reference's own ActionType enum (duck typing)
    return (action.vessel_idx, action.port_idx, action.quantity, 1 if kind.upper().endswith("DISCHARGE") else 0)


def encode_action(action, out_row) -> None:
    """Write one Action (ours or the reference's — duck-typed) into an int32[4] row of the C ABI."""
    kind = getattr(action.action_type, "name", None) or str(action.action_type)
    out_row[:] = (action.vessel_idx, action.port_idx, action.quantity, int(kind.upper().endswith("DISCHARGE")))
