"""vm_scheduling agent-facing types with the reference's names and fields
(maro/simulator/scenarios/vm_scheduling/common.py:9-170, enums.py)."""
from enum import IntEnum
from typing import List

from ... import _abi


class VmCategory(IntEnum):
    DELAY_INSENSITIVE = 0
    INTERACTIVE = 1
    UNKNOWN = 2


class Action:
    def __init__(self, vm_id: int):
        self.vm_id = vm_id

    def __repr__(self):
        return "%s {vm_id: %r}" % (type(self).__name__, self.vm_id)


class AllocateAction(Action):
    def __init__(self, vm_id: int, pm_id: int):
        super().__init__(vm_id)
        self.pm_id = pm_id

    def __repr__(self):
        return "%s {vm_id: %r, pm_id: %r}" % (type(self).__name__, self.vm_id, self.pm_id)


class PostponeAction(Action):
    def __init__(self, vm_id: int, postpone_step: int):
        super().__init__(vm_id)
        self.postpone_step = postpone_step

    def __repr__(self):
        return "%s {vm_id: %r, postpone_step: %r}" % (type(self).__name__, self.vm_id, self.postpone_step)


class DecisionEvent:
    summary_key = ["frame_index", "valid_pms", "vm_id", "vm_cpu_cores_requirement", "vm_memory_requirement",
                   "remaining_buffer_time"]

    def __init__(self, frame_index: int, valid_pms: List[int], vm_id: int, vm_cpu_cores_requirement: int,
                 vm_memory_requirement: int, vm_sub_id: int, vm_category: int, remaining_buffer_time: int):
        self.frame_index = frame_index
        self.valid_pms = valid_pms
        self.vm_id = vm_id
        self.vm_cpu_cores_requirement = vm_cpu_cores_requirement
        self.vm_memory_requirement = vm_memory_requirement
        self.vm_sub_id = vm_sub_id
        self.vm_category = vm_category
        self.remaining_buffer_time = remaining_buffer_time

    def __repr__(self):
        return "%s {%s}" % (type(self).__name__, ", ".join(
            f"{k}: {getattr(self, k)!r}" for k in ("frame_index", "valid_pms", "vm_id", "vm_cpu_cores_requirement",
                                                   "vm_memory_requirement", "vm_sub_id", "vm_category",
                                                   "remaining_buffer_time")))


class Latency:
    """Accumulated postponement latency (common.py:140-170)."""

    def __init__(self, due_to_agent: int = 0, due_to_resource: int = 0):
        self.due_to_agent = due_to_agent
        self.due_to_resource = due_to_resource

    def __repr__(self):
        return "%s {due_to_agent: %r, due_to_resource: %r}" % (type(self).__name__, self.due_to_agent, self.due_to_resource)


def encode_vm_action(action, out_row) -> None:
    """-> the 4-int32 action row of include/maro_b200.h (kind 0 AllocateAction, 1 PostponeAction)."""
    if isinstance(action, AllocateAction):
        out_row[:] = (action.vm_id, _abi.VM_ACTION_ALLOCATE, action.pm_id, 0)
    elif isinstance(action, PostponeAction):
        out_row[:] = (action.vm_id, _abi.VM_ACTION_POSTPONE, action.postpone_step, 0)
    else:
        raise TypeError(f"vm_scheduling actions are AllocateAction / PostponeAction, got {type(action).__name__}")


def decode_vm_decision(row) -> DecisionEvent:
    n = int(row[_abi.VM_DEC_N_VALID])
    ev = DecisionEvent(int(row[_abi.VM_DEC_FRAME_INDEX]), [int(x) for x in row[_abi.VM_DEC_HEAD:_abi.VM_DEC_HEAD + n]],
                       int(row[_abi.VM_DEC_VM_ID]), int(row[_abi.VM_DEC_CPU]), int(row[_abi.VM_DEC_MEMORY]),
                       int(row[_abi.VM_DEC_SUB_ID]), VmCategory(int(row[_abi.VM_DEC_CATEGORY])),
                       int(row[_abi.VM_DEC_BUFFER_TIME]))
    # extension (not part of the reference's DecisionEvent): remaining CPU cores of each valid PM in the decision's frame — what
    # the reference's rule-based agents fetch with a snapshot query (rule_based_algorithm/best_fit.py:38-44)
    ext = int(row[11]) if len(row) > 11 else 0
    ev.valid_pms_remaining_cpu_cores = [int(x) for x in row[ext:ext + n]] if ext else None
    return ev


def decode_vm_metrics(row) -> dict:
    """-> the keys of VmSchedulingBusinessEngine.get_metrics (business_engine.py:551-571)."""
    d = _abi.vm_metrics_dict(row)
    d["total_latency"] = Latency(d.pop("latency_due_to_agent"), d.pop("latency_due_to_resource"))
    return d
