"""citi_bike agent-facing types with the reference's names and fields
(maro/simulator/scenarios/citi_bike/common.py:9-160)."""
from enum import Enum


class DecisionType(Enum):
    Supply = "supply"
    Demand = "demand"


class ExtraCostMode(Enum):
    Source = "source"
    Target = "target"
    TargetNeighbors = "target_neighbors"


class Action:
    def __init__(self, from_station_idx: int, to_station_idx: int, number: int):
        self.from_station_idx = from_station_idx
        self.to_station_idx = to_station_idx
        self.number = number

    def __repr__(self):
        return "%s {from_station_idx: %r, to_station_idx: %r, number:%r}" % (
            self.__class__.__name__, self.from_station_idx, str(self.to_station_idx), self.number)


class DecisionEvent:
    summary_key = ["station_idx", "tick", "frame_index", "type", "action_scope"]

    def __init__(self, station_idx, tick, frame_index, action_scope, decision_type):
        self.station_idx = station_idx
        self.tick = tick
        self.frame_index = frame_index
        self.type = decision_type
        self.action_scope = action_scope  # dict station index -> max supply / demand number (a mapping)

    def __getstate__(self):
        return {"station_idx": self.station_idx, "tick": self.tick, "frame_index": self.frame_index, "type": self.type,
                "action_scope": self.action_scope}

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __repr__(self):
        return "%s {station_idx: %r, type: %r, action_scope:%r}" % (
            self.__class__.__name__, self.station_idx, str(self.type), self.action_scope)


def encode_bike_action(action, out_row) -> None:
    out_row[0] = action.from_station_idx
    out_row[1] = action.to_station_idx
    out_row[2] = action.number
    out_row[3] = 0


def decode_bike_decision(row, snapshot_list=None) -> DecisionEvent:
    n = int(row[4])
    scope = {int(row[8 + 2 * k]): int(row[9 + 2 * k]) for k in range(n)}
    # the reference appends the station itself last (decision_strategy.py:287-291)
    me = int(row[1])
    if me in scope:
        scope[me] = scope.pop(me)
    return DecisionEvent(me, int(row[0]), int(row[2]), scope, DecisionType.Supply if row[3] == 0 else DecisionType.Demand)
