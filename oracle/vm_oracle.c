/*
 * vm_oracle.c — TEST INFRASTRUCTURE.  CPU restatement (plain C, one replica) of the reference's vm_scheduling
 * `Env.step` hot path.  Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity status: PINNED — tests/test_vm_oracle_golden.py checks it against traces of the unmodified reference
 * (tests/golden/gen_vm_golden.py: the reference's own toy fixture tests/data/vm_scheduling/azure.2019.toy and
 * synthetic traces in the same .bin schema) and the known answers of tests/vm_scheduling/test_vm_scheduling_scenario.py
 * (config counts :112-136, price model :155-171).  Integers bit-exact; floats (incomes, energy, profit) to 1e-9
 * relative: the reference sums utilisation over a Python `set` of VM ids, whose iteration order is not restated.
 *
 * Follows the reference's structure:
 *   Env._simulate / step / _assign_action        maro/simulator/core.py:92-133, 301-381
 *   EventBuffer.execute, EventLinkedList          maro/event_buffer/event_buffer.py:177-247, event_linked_list.py:53-137
 *   VmSchedulingBusinessEngine.step / post_step   maro/simulator/scenarios/vm_scheduling/business_engine.py:449-525
 *   _update_* / _overload / energy model          :575-688
 *   _postpone_vm_request / _get_valid_pms         :690-768
 *   _process_finished_vm / _on_vm_required / _on_action_received   :770-905
 *   PhysicalMachine.update_cpu_utilization        physical_machine.py:54-63  (numpy round(x, 2), float64 live attribute)
 *   VirtualMachine utilisation series             virtual_machine.py:60-90
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -I../include vm_oracle.c -o _build/libvm_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "maro_b200.h"

enum { EV_REQUEST, EV_PENDING_DECISION, EV_TAKE_ACTION };
enum { ST_PENDING, ST_EXECUTING, ST_FINISHED };
typedef struct Event {
    int tick, type, state, cascade;
    int vm;                 /* vm index */
    int32_t actions[16][4];
    int n_actions;
    struct Event* next;
    struct Event* imm_head;
    struct Event* imm_tail;
    int imm_count;
} Event;
typedef struct { Event* head; Event* tail; int count; } EvList;

typedef struct { int32_t id, cpu_cap, mem_cap, pm_type, cpu_alloc, mem_alloc, oversub, region, zone, dc, cluster, rack;
                 double cpu_util, energy; int n_live; } Pm;
typedef struct { int state; /* 0 unseen, 1 pending, 2 live, 3 gone */ int pm, creation, deletion, budget; double cpu_util; } Vm;

typedef struct VmOracle {
    MaroVmTopology t;
    int start_tick, max_tick, snap_res, ring_rows;
    Pm* pms;
    Vm* vms;
    int* live;   /* insertion-ordered live VM indices (python dict order) */
    int n_live;
    int32_t *rack_empty, *cluster_empty, *dc_empty, *zone_empty, *region_empty;
    EvList* pending;
    int n_lists;
    int64_t total_vm_requests, successful_allocation, successful_completion, failed_allocation, failed_completion;
    int64_t latency_agent, latency_resource, total_oversubscriptions, total_overload_pms, total_overload_vms;
    double total_incomes, total_profit, energy_cost, total_energy;
    int tick, gen_state, ep_step, pending_action_vm;
    Event* pending_decision;
    int64_t n_events, n_ticks, n_snapshots, n_steps;
    int frame_words;
    int32_t* ring;
    int32_t* ring_frame;
    int error;
} VmOracle;

static void* dup_arr(const void* p, size_t bytes) { void* q = malloc(bytes ? bytes : 1); if (bytes) memcpy(q, p, bytes); return q; }

/*
 * Scalar types.  The reference's default ("static") backend builds its structured dtype from the decoded attribute
 * type NAME (np_backend.pyx:177, :143-148, :293): "float" -> float64, "short" -> int16.  So the live frame keeps
 * cpu_utilization / energy_consumption in DOUBLE precision (live row and snapshot rows alike) and hands them back as
 * np.float64.  Every expression below is therefore plain double arithmetic (the exported frame words carry the two
 * attributes as float32, the device ring's storage format; maro_b200.batch.VmBatch.query lifts them back), and
 * round(np.float64, 2) is numpy's around: rint(x * 100) / 100.
 */
static double np_round2(double x) { return rint(x * 100.0) / 100.0; }

/* _cpu_utilization_to_energy_consumption (:671-688) */
static double energy_of(const VmOracle* o, int pm_type, double cpu_utilization) {
    const double* pc = o->t.pmtype_power + 3 * pm_type;
    cpu_utilization /= 100;
    if (1 < cpu_utilization) cpu_utilization = 1;
    double per_hour = pc[2] + (pc[1] - pc[2]) * (2 * cpu_utilization - pow(cpu_utilization, pc[0]));
    return (per_hour / o->t.ticks_per_hour) / 1000;
}

/* PhysicalMachine.update_cpu_utilization (physical_machine.py:54-63) */
static void pm_set_util(Pm* pm, double cpu_utilization) {
    pm->cpu_util = cpu_utilization > 0 ? np_round2(cpu_utilization) : 0.0;
}
static void pm_add_vm_util(Pm* pm, int cores, double vm_util) {
    pm_set_util(pm, ((double)pm->cpu_cap * pm->cpu_util + (double)cores * vm_util) / (double)pm->cpu_cap);
}

/* utilisation series of a VM: series[k] = forward-filled reading at request tick + k */
static double series_at(const VmOracle* o, int vm, int k) { return o->t.util_val[o->t.util_offset[vm] + k]; }
static int series_has(const VmOracle* o, int vm, int k) { return o->t.util_has[o->t.util_offset[vm] + k]; }
static double clamp100(double x) { return x < 0 ? 0 : (x > 100 ? 100 : x); }

static Event* gen_event(int tick, int type, int cascade) {
    Event* e = (Event*)calloc(1, sizeof(Event));
    e->tick = tick; e->type = type; e->cascade = cascade;
    return e;
}
static void insert_event(VmOracle* o, Event* e) {
    if (e->tick < 0 || e->tick >= o->n_lists) { free(e); return; }
    EvList* l = &o->pending[e->tick];
    e->next = NULL;
    if (l->tail) l->tail->next = e; else l->head = e;
    l->tail = e; l->count++;
}
static void add_immediate_event(Event* parent, Event* e, int is_head) {
    if (parent->imm_count == 0) { parent->imm_head = parent->imm_tail = e; e->next = NULL; }
    else if (is_head) { e->next = parent->imm_head; parent->imm_head = e; }
    else { parent->imm_tail->next = e; parent->imm_tail = e; e->next = NULL; }
    parent->imm_count++;
}
static void clear_finished(EvList* l) {
    while (l->head && l->head->state == ST_FINISHED) {
        Event* e = l->head;
        l->head = e->next;
        if (!l->head) l->tail = NULL;
        l->count--;
        if (e->cascade && e->imm_count) {
            e->imm_tail->next = l->head;
            if (!l->head) l->tail = e->imm_tail;
            l->head = e->imm_head;
            l->count += e->imm_count;
        }
        free(e);
    }
}

static void live_remove(VmOracle* o, int vm) {
    int k = 0;
    while (k < o->n_live && o->live[k] != vm) k++;
    for (; k + 1 < o->n_live; k++) o->live[k] = o->live[k + 1];
    o->n_live--;
}

/* _postpone_vm_request (:690-713) */
static void postpone(VmOracle* o, int resource, int vm, int remaining_buffer_time) {
    if (remaining_buffer_time >= o->t.delay_duration) {
        if (resource) o->latency_resource += o->t.delay_duration; else o->latency_agent += o->t.delay_duration;
        o->vms[vm].budget -= o->t.delay_duration;  /* postpone_payload.remaining_buffer_time -= delay */
        Event* e = gen_event(o->tick + o->t.delay_duration, EV_REQUEST, 1);
        e->vm = vm;
        insert_event(o, e);
    } else {
        o->vms[vm].state = 3;
        o->failed_allocation += 1;
    }
}

/* _get_valid_pms (:715-768) */
static int valid_pms(const VmOracle* o, int vm, int32_t* out) {
    const int32_t* a = o->t.vm_attr + 8 * vm;
    int cores = a[6], mem = a[7], cat = a[5], n = 0;
    for (int p = 0; p < o->t.n_pm; p++) {
        const Pm* pm = &o->pms[p];
        if (cat == 1 || cat == 2) {
            if ((pm->oversub == 0 || pm->oversub == -1) && pm->cpu_alloc + cores <= pm->cpu_cap && pm->mem_alloc + mem <= pm->mem_cap)
                out[n++] = p;
        } else {
            if ((pm->oversub == 0 || pm->oversub == 1) && pm->cpu_alloc + cores <= o->t.max_cpu_over * pm->cpu_cap &&
                pm->mem_alloc + mem <= o->t.max_mem_over * pm->mem_cap &&
                pm->cpu_util / 100 * pm->cpu_cap + cores <= o->t.max_util_rate * pm->cpu_cap)
                out[n++] = p;
        }
    }
    return n;
}

static void on_vm_required(VmOracle* o, Event* ev) { /* :783-826 */
    int vm = ev->vm;
    o->vms[vm].state = 1;
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (size_t)(o->t.n_pm + 1));
    int n = valid_pms(o, vm, tmp);
    free(tmp);
    if (n > 0) {
        o->pending_action_vm = vm;
        Event* d = gen_event(ev->tick, EV_PENDING_DECISION, 1);
        d->vm = vm;
        add_immediate_event(ev, d, 0);
    } else {
        postpone(o, 1, vm, o->vms[vm].budget);
    }
}

static int find_vm(const VmOracle* o, int vm_id) {
    int lo = 0, hi = o->t.n_vm - 1;
    while (lo <= hi) {
        int mid = (lo + hi) / 2;
        if (o->t.vm_sorted_ids[mid] == vm_id) return o->t.vm_sorted_idx[mid];
        if (o->t.vm_sorted_ids[mid] < vm_id) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

static void on_action_received(VmOracle* o, Event* ev) { /* :828-905 */
    if (ev->n_actions == 0) {
        o->vms[o->pending_action_vm].state = 3;  /* popped from the pending payloads */
        return;
    }
    for (int i = 0; i < ev->n_actions; i++) {
        int vm = find_vm(o, ev->actions[i][0]);
        if (vm < 0 || o->vms[vm].state != 1) { o->error = 1; return; }  /* "The VM id ... sent by agent is invalid." */
        const int32_t* a = o->t.vm_attr + 8 * vm;
        if (ev->actions[i][1] == MARO_VM_ACTION_ALLOCATE) {
            int p = ev->actions[i][2];
            if (p < 0 || p >= o->t.n_pm) { o->error = 1; return; }
            Vm* v = &o->vms[vm];
            v->pm = p; v->creation = ev->tick; v->deletion = ev->tick + a[4];
            v->cpu_util = clamp100(series_at(o, vm, 0));  /* get_utilization(cur_tick): series[cur - creation] */
            v->state = 2;
            o->live[o->n_live++] = vm;
            Pm* pm = &o->pms[p];
            if (pm->oversub == 0) pm->oversub = a[5] == 0 ? 1 : -1;
            pm->n_live++;
            pm->cpu_alloc += a[6];
            pm->mem_alloc += a[7];
            pm_add_vm_util(pm, a[6], v->cpu_util);
            pm->energy = energy_of(o, pm->pm_type, pm->cpu_util);
            o->successful_allocation += 1;
        } else {
            int step = ev->actions[i][2];
            postpone(o, 0, vm, o->vms[vm].budget - step * o->t.delay_duration);
        }
    }
}

static Event* execute(VmOracle* o, int tick) {
    EvList* l = &o->pending[tick];
    while (l->count) {
        clear_finished(l);
        Event* e = l->head;
        if (!e) break;
        if (e->type == EV_PENDING_DECISION && e->state != ST_EXECUTING) return e;
        e->state = ST_EXECUTING;
        if (e->type == EV_REQUEST) on_vm_required(o, e);
        else if (e->type == EV_TAKE_ACTION) on_action_received(o, e);
        e->state = ST_FINISHED;
        o->n_events++;
    }
    return NULL;
}

/* canonical frame words (attr-major inside each node type, alphabetical attribute order) */
static void export_frame(const VmOracle* o, int32_t* w) {
    int N = o->t.n_pm;
    for (int p = 0; p < N; p++) {
        const Pm* pm = &o->pms[p];
        int32_t v[14];
        v[0] = pm->cluster; v[1] = pm->cpu_alloc; v[2] = pm->cpu_cap; { float f = (float)pm->cpu_util; memcpy(&v[3], &f, 4); } v[4] = pm->dc;
        { float f = (float)pm->energy; memcpy(&v[5], &f, 4); } v[6] = pm->id; v[7] = pm->mem_alloc; v[8] = pm->mem_cap; v[9] = pm->oversub;
        v[10] = pm->pm_type; v[11] = pm->rack; v[12] = pm->region; v[13] = pm->zone;
        for (int a = 0; a < 14; a++) w[a * N + p] = v[a];
    }
    int32_t* q = w + 14 * N;
    int R = o->t.n_rack;
    for (int i = 0; i < R; i++) {
        const int32_t* ids = o->t.rack_ids + 4 * i;
        int tot = o->t.rack_range[2 * i + 1] - o->t.rack_range[2 * i];
        int32_t v[7] = {ids[3], ids[2], o->rack_empty[i], i, ids[0], tot, ids[1]};
        for (int a = 0; a < 7; a++) q[a * R + i] = v[a];
    }
    q += 7 * R;
    int C = o->t.n_cluster;
    for (int i = 0; i < C; i++) {
        const int32_t* ids = o->t.cluster_ids + 3 * i;
        int tot = 0;
        for (int k = o->t.cluster_range[2 * i]; k < o->t.cluster_range[2 * i + 1]; k++) tot += o->t.rack_range[2 * k + 1] - o->t.rack_range[2 * k];
        int32_t v[6] = {ids[2], o->cluster_empty[i], i, ids[0], tot, ids[1]};
        for (int a = 0; a < 6; a++) q[a * C + i] = v[a];
    }
    q += 6 * C;
    int D = o->t.n_dc;
    for (int i = 0; i < D; i++) {
        int tot = 0;
        for (int c = o->t.dc_range[2 * i]; c < o->t.dc_range[2 * i + 1]; c++)
            for (int k = o->t.cluster_range[2 * c]; k < o->t.cluster_range[2 * c + 1]; k++) tot += o->t.rack_range[2 * k + 1] - o->t.rack_range[2 * k];
        int32_t v[5] = {o->dc_empty[i], i, o->t.dc_ids[2 * i], tot, o->t.dc_ids[2 * i + 1]};
        for (int a = 0; a < 5; a++) q[a * D + i] = v[a];
    }
    q += 5 * D;
    int Z = o->t.n_zone;
    for (int i = 0; i < Z; i++) {
        int tot = 0;
        for (int d = o->t.zone_range[2 * i]; d < o->t.zone_range[2 * i + 1]; d++)
            for (int c = o->t.dc_range[2 * d]; c < o->t.dc_range[2 * d + 1]; c++)
                for (int k = o->t.cluster_range[2 * c]; k < o->t.cluster_range[2 * c + 1]; k++) tot += o->t.rack_range[2 * k + 1] - o->t.rack_range[2 * k];
        int32_t v[4] = {o->zone_empty[i], i, o->t.zone_ids[i], tot};
        for (int a = 0; a < 4; a++) q[a * Z + i] = v[a];
    }
    q += 4 * Z;
    int G = o->t.n_region;
    for (int i = 0; i < G; i++) {
        int tot = 0;
        for (int z = o->t.region_range[2 * i]; z < o->t.region_range[2 * i + 1]; z++)
            for (int d = o->t.zone_range[2 * z]; d < o->t.zone_range[2 * z + 1]; d++)
                for (int c = o->t.dc_range[2 * d]; c < o->t.dc_range[2 * d + 1]; c++)
                    for (int k = o->t.cluster_range[2 * c]; k < o->t.cluster_range[2 * c + 1]; k++) tot += o->t.rack_range[2 * k + 1] - o->t.rack_range[2 * k];
        int32_t v[3] = {o->region_empty[i], i, tot};
        for (int a = 0; a < 3; a++) q[a * G + i] = v[a];
    }
}
static void take_snapshot(VmOracle* o, int frame_index) {
    int row = frame_index % o->ring_rows;
    export_frame(o, o->ring + (size_t)row * o->frame_words);
    o->ring_frame[row] = frame_index;
    o->n_snapshots++;
}
static int frame_index_of(const VmOracle* o, int tick) { return (tick - o->start_tick) / o->snap_res; }

static void be_step(VmOracle* o, int tick) { /* :449-493 */
    /* _process_finished_vm (:770-781) */
    int kept = 0;
    for (int k = 0; k < o->n_live; k++) {
        int vm = o->live[k];
        Vm* v = &o->vms[vm];
        if (v->deletion == tick) {
            Pm* pm = &o->pms[v->pm];
            pm->cpu_alloc -= o->t.vm_attr[8 * vm + 6];
            pm->mem_alloc -= o->t.vm_attr[8 * vm + 7];
            pm->n_live--;
            if (pm->n_live == 0) pm->oversub = 0;
            v->state = 3;
            o->successful_completion += 1;
        } else {
            o->live[kept++] = vm;
        }
    }
    o->n_live = kept;
    /* _update_vm_workload (:575-592): live VMs whose trace has a reading at this tick take series[tick - creation] */
    for (int k = 0; k < o->n_live; k++) {
        int vm = o->live[k];
        Vm* v = &o->vms[vm];
        int rel = tick - o->t.vm_attr[8 * vm + 3];
        if (series_has(o, vm, rel)) v->cpu_util = clamp100(series_at(o, vm, tick - v->creation));
    }
    /* _update_pm_workload (:640-652) */
    for (int p = 0; p < o->t.n_pm; p++) {
        Pm* pm = &o->pms[p];
        double used = 0.0;
        for (int k = 0; k < o->n_live; k++) {
            int vm = o->live[k];
            if (o->vms[vm].pm == p) used += o->vms[vm].cpu_util * o->t.vm_attr[8 * vm + 6];
        }
        pm_set_util(pm, used / pm->cpu_cap);
        pm->energy = energy_of(o, pm->pm_type, pm->cpu_util);
    }
    /* _update_upper_level_metrics (:594-638) */
    for (int i = 0; i < o->t.n_rack; i++) {
        int c = 0;
        for (int p = o->t.rack_range[2 * i]; p < o->t.rack_range[2 * i + 1]; p++) c += o->pms[p].cpu_alloc == 0;
        o->rack_empty[i] = c;
    }
    for (int i = 0; i < o->t.n_cluster; i++) { int c = 0; for (int k = o->t.cluster_range[2 * i]; k < o->t.cluster_range[2 * i + 1]; k++) c += o->rack_empty[k]; o->cluster_empty[i] = c; }
    for (int i = 0; i < o->t.n_dc; i++) { int c = 0; for (int k = o->t.dc_range[2 * i]; k < o->t.dc_range[2 * i + 1]; k++) c += o->cluster_empty[k]; o->dc_empty[i] = c; }
    for (int i = 0; i < o->t.n_zone; i++) { int c = 0; for (int k = o->t.zone_range[2 * i]; k < o->t.zone_range[2 * i + 1]; k++) c += o->dc_empty[k]; o->zone_empty[i] = c; }
    for (int i = 0; i < o->t.n_region; i++) { int c = 0; for (int k = o->t.region_range[2 * i]; k < o->t.region_range[2 * i + 1]; k++) c += o->zone_empty[k]; o->region_empty[i] = c; }
    /* new requests of this tick */
    for (int vm = o->t.req_offset[tick]; vm < o->t.req_offset[tick + 1]; vm++) {
        o->vms[vm].budget = o->t.buffer_budget;
        Event* e = gen_event(tick, EV_REQUEST, 1);
        e->vm = vm;
        insert_event(o, e);
        o->total_vm_requests += 1;
    }
}

static int be_post_step(VmOracle* o, int tick) { /* :495-525 */
    double total_energy = 0.0, total_energy_cost = 0.0;
    for (int p = 0; p < o->t.n_pm; p++) {
        Pm* pm = &o->pms[p];
        if (pm->oversub && pm->cpu_alloc > pm->cpu_cap) o->total_oversubscriptions += 1;
        total_energy += pm->energy;
        double pm_cost = pm->energy * o->t.unit_energy_price * o->t.pue;
        total_energy_cost += pm_cost;
        if (pm->cpu_util > 100) { /* _overload (:654-669) */
            int n = 0;
            for (int k = 0; k < o->n_live; k++) n += o->vms[o->live[k]].pm == p;
            if (o->t.kill_all) {
                int kept = 0;
                for (int k = 0; k < o->n_live; k++) {
                    int vm = o->live[k];
                    if (o->vms[vm].pm == p) {
                        o->total_incomes -= o->t.vm_price[vm] * (tick - o->vms[vm].creation);
                        o->vms[vm].state = 3;
                    } else o->live[kept++] = vm;
                }
                o->n_live = kept;
                pm->n_live = 0;
                o->failed_completion += n;
            }
            o->total_overload_vms += n;
        }
    }
    o->total_energy += total_energy;
    o->energy_cost += total_energy_cost;
    for (int k = 0; k < o->n_live; k++) o->total_incomes += o->t.vm_price[o->live[k]];
    o->total_profit = o->total_incomes - o->energy_cost;
    if ((tick + 1) % o->snap_res == 0) take_snapshot(o, frame_index_of(o, tick));
    return tick + 1 >= o->max_tick;
}

static void free_events(VmOracle* o) {
    for (int i = 0; i < o->n_lists; i++) {
        Event* e = o->pending[i].head;
        while (e) {
            Event* n = e->next;
            Event* s = e->imm_head;
            for (int k = 0; k < e->imm_count && s; k++) { Event* sn = s->next; free(s); s = sn; }
            free(e);
            e = n;
        }
        o->pending[i].head = o->pending[i].tail = NULL; o->pending[i].count = 0;
    }
}

void vm_oracle_reset(VmOracle* o) {
    free_events(o);
    for (int p = 0; p < o->t.n_pm; p++) {
        const int32_t* a = o->t.pm_attr + 8 * p;
        Pm* pm = &o->pms[p];
        memset(pm, 0, sizeof(Pm));
        pm->id = p; pm->cpu_cap = a[0]; pm->mem_cap = a[1]; pm->pm_type = a[2]; pm->region = a[3]; pm->zone = a[4];
        pm->dc = a[5]; pm->cluster = a[6]; pm->rack = a[7];
        pm->energy = o->t.pm_idle_energy[p];
    }
    memset(o->vms, 0, sizeof(Vm) * (size_t)(o->t.n_vm ? o->t.n_vm : 1));
    o->n_live = 0;
    for (int i = 0; i < o->t.n_rack; i++) o->rack_empty[i] = o->t.rack_range[2 * i + 1] - o->t.rack_range[2 * i];
    for (int i = 0; i < o->t.n_cluster; i++) { int c = 0; for (int k = o->t.cluster_range[2 * i]; k < o->t.cluster_range[2 * i + 1]; k++) c += o->rack_empty[k]; o->cluster_empty[i] = c; }
    for (int i = 0; i < o->t.n_dc; i++) { int c = 0; for (int k = o->t.dc_range[2 * i]; k < o->t.dc_range[2 * i + 1]; k++) c += o->cluster_empty[k]; o->dc_empty[i] = c; }
    for (int i = 0; i < o->t.n_zone; i++) { int c = 0; for (int k = o->t.zone_range[2 * i]; k < o->t.zone_range[2 * i + 1]; k++) c += o->dc_empty[k]; o->zone_empty[i] = c; }
    for (int i = 0; i < o->t.n_region; i++) { int c = 0; for (int k = o->t.region_range[2 * i]; k < o->t.region_range[2 * i + 1]; k++) c += o->zone_empty[k]; o->region_empty[i] = c; }
    for (int i = 0; i < o->ring_rows; i++) o->ring_frame[i] = -1;
    o->total_vm_requests = o->successful_allocation = o->successful_completion = o->failed_allocation = o->failed_completion = 0;
    o->latency_agent = o->latency_resource = o->total_oversubscriptions = o->total_overload_pms = o->total_overload_vms = 0;
    o->total_incomes = o->total_profit = o->energy_cost = o->total_energy = 0.0;
    o->tick = o->start_tick; o->gen_state = 0; o->ep_step = 0; o->pending_decision = NULL; o->pending_action_vm = 0; o->error = 0;
}

VmOracle* vm_oracle_create(const MaroVmTopology* t, int start_tick, int snapshot_resolution, int max_snapshots) {
    VmOracle* o = (VmOracle*)calloc(1, sizeof(VmOracle));
    o->t = *t;
    int nu = t->util_offset[t->n_vm];
#define DUPI(f, n) o->t.f = (const int32_t*)dup_arr(t->f, sizeof(int32_t) * (size_t)(n))
#define DUPD(f, n) o->t.f = (const double*)dup_arr(t->f, sizeof(double) * (size_t)(n))
    DUPI(pm_attr, 8 * t->n_pm); DUPD(pm_idle_energy, t->n_pm); DUPD(pmtype_power, 3 * t->n_pm_types);
    DUPI(rack_range, 2 * t->n_rack); DUPI(rack_ids, 4 * t->n_rack); DUPI(cluster_range, 2 * t->n_cluster);
    DUPI(cluster_ids, 3 * t->n_cluster); DUPI(dc_range, 2 * t->n_dc); DUPI(dc_ids, 2 * t->n_dc); DUPI(zone_range, 2 * t->n_zone);
    DUPI(zone_ids, t->n_zone); DUPI(region_range, 2 * t->n_region); DUPI(vm_attr, 8 * t->n_vm); DUPD(vm_price, t->n_vm);
    DUPI(req_offset, t->max_tick + 1); DUPI(vm_sorted_ids, t->n_vm); DUPI(vm_sorted_idx, t->n_vm); DUPI(util_offset, t->n_vm + 1);
    DUPD(util_val, nu); DUPI(util_has, nu);
#undef DUPI
#undef DUPD
    o->start_tick = start_tick; o->max_tick = t->max_tick; o->snap_res = snapshot_resolution;
    int total_frames = (o->max_tick - start_tick + snapshot_resolution - 1) / snapshot_resolution;
    o->ring_rows = max_snapshots > 0 && max_snapshots < total_frames ? max_snapshots : total_frames;
    if (o->ring_rows < 1) o->ring_rows = 1;
    o->pms = (Pm*)calloc(t->n_pm, sizeof(Pm));
    o->vms = (Vm*)calloc(t->n_vm ? t->n_vm : 1, sizeof(Vm));
    o->live = (int*)calloc(t->n_vm ? t->n_vm : 1, sizeof(int));
    o->rack_empty = (int32_t*)calloc(t->n_rack, 4); o->cluster_empty = (int32_t*)calloc(t->n_cluster, 4);
    o->dc_empty = (int32_t*)calloc(t->n_dc, 4); o->zone_empty = (int32_t*)calloc(t->n_zone, 4); o->region_empty = (int32_t*)calloc(t->n_region, 4);
    o->n_lists = o->max_tick + 1;
    o->pending = (EvList*)calloc(o->n_lists, sizeof(EvList));
    o->frame_words = 14 * t->n_pm + 7 * t->n_rack + 6 * t->n_cluster + 5 * t->n_dc + 4 * t->n_zone + 3 * t->n_region;
    o->ring = (int32_t*)calloc((size_t)o->ring_rows * o->frame_words, 4);
    o->ring_frame = (int32_t*)calloc(o->ring_rows, 4);
    vm_oracle_reset(o);
    return o;
}
void vm_oracle_destroy(VmOracle* o) {
    if (!o) return;
    free_events(o);
    free(o->pending); free(o->ring); free(o->ring_frame); free(o->pms); free(o->vms); free(o->live);
    free(o->rack_empty); free(o->cluster_empty); free(o->dc_empty); free(o->zone_empty); free(o->region_empty);
    free((void*)o->t.pm_attr); free((void*)o->t.pm_idle_energy); free((void*)o->t.pmtype_power); free((void*)o->t.rack_range);
    free((void*)o->t.rack_ids); free((void*)o->t.cluster_range); free((void*)o->t.cluster_ids); free((void*)o->t.dc_range);
    free((void*)o->t.dc_ids); free((void*)o->t.zone_range); free((void*)o->t.zone_ids); free((void*)o->t.region_range);
    free((void*)o->t.vm_attr); free((void*)o->t.vm_price); free((void*)o->t.req_offset); free((void*)o->t.vm_sorted_ids);
    free((void*)o->t.vm_sorted_idx); free((void*)o->t.util_offset); free((void*)o->t.util_val); free((void*)o->t.util_has);
    free(o);
}

static void put_f64(int64_t* m, int i, double x) { memcpy(&m[i], &x, 8); }
static void fill_metrics(const VmOracle* o, int64_t* m) {
    memset(m, 0, 8 * MARO_VM_METRIC_WORDS);
    m[0] = o->total_vm_requests; put_f64(m, 1, o->total_incomes); put_f64(m, 2, o->energy_cost); put_f64(m, 3, o->total_profit);
    put_f64(m, 4, o->total_energy); m[5] = o->successful_allocation; m[6] = o->successful_completion; m[7] = o->failed_allocation;
    m[8] = o->failed_completion; m[9] = o->latency_agent; m[10] = o->latency_resource; m[11] = o->total_oversubscriptions;
    m[12] = o->total_overload_pms; m[13] = o->total_overload_vms;
}

int vm_oracle_decision_words(const VmOracle* o) { return MARO_VM_DEC_HEAD + o->t.n_pm; }

int vm_oracle_step(VmOracle* o, const int32_t* actions, int n_actions, int32_t* decision, int64_t* metrics) {
    memset(decision, 0, sizeof(int32_t) * (size_t)vm_oracle_decision_words(o));
    if (o->gen_state >= 2) {
        o->gen_state = 3;
        decision[MARO_VM_DEC_STATUS] = MARO_STATUS_FINISHED;
        memset(metrics, 0, 8 * MARO_VM_METRIC_WORDS);
        return MARO_STATUS_FINISHED;
    }
    o->n_steps++;
    int resume = o->gen_state == 1;
    o->gen_state = 1;
    if (resume) {
        Event* d = o->pending_decision;
        d->state = ST_EXECUTING;
        Event* a = gen_event(o->tick, EV_TAKE_ACTION, 1);
        a->n_actions = n_actions > 16 ? 16 : (n_actions < 0 ? 0 : n_actions);
        for (int i = 0; i < a->n_actions; i++) memcpy(a->actions[i], actions + 4 * i, 16);
        add_immediate_event(d, a, 1);
        o->pending_decision = NULL;
    }
    for (;;) {
        if (!resume) { be_step(o, o->tick); o->n_ticks++; }
        resume = 0;
        Event* d = execute(o, o->tick);
        if (o->error) { decision[MARO_VM_DEC_STATUS] = MARO_STATUS_BAD_ACTION; o->gen_state = 3; return MARO_STATUS_BAD_ACTION; }
        if (d) {
            take_snapshot(o, frame_index_of(o, o->tick));
            o->pending_decision = d;
            const int32_t* a = o->t.vm_attr + 8 * d->vm;
            decision[MARO_VM_DEC_TICK] = o->tick;
            decision[MARO_VM_DEC_VM_ID] = a[0];
            decision[MARO_VM_DEC_FRAME_INDEX] = frame_index_of(o, o->tick);
            decision[MARO_VM_DEC_CPU] = a[6];
            decision[MARO_VM_DEC_MEMORY] = a[7];
            decision[MARO_VM_DEC_SUB_ID] = a[1];
            decision[MARO_VM_DEC_CATEGORY] = a[5];
            decision[MARO_VM_DEC_BUFFER_TIME] = o->vms[d->vm].budget;
            decision[MARO_VM_DEC_N_VALID] = valid_pms(o, d->vm, decision + MARO_VM_DEC_HEAD);
            decision[MARO_VM_DEC_STATUS] = MARO_STATUS_DECISION;
            decision[MARO_VM_DEC_STEP] = o->ep_step++;
            fill_metrics(o, metrics);
            return MARO_STATUS_DECISION;
        }
        if (be_post_step(o, o->tick)) break;
        o->tick += 1;
    }
    if ((o->tick + 1) % o->snap_res != 0) take_snapshot(o, frame_index_of(o, o->tick));
    o->gen_state = 2;
    decision[MARO_VM_DEC_TICK] = o->tick;
    decision[MARO_VM_DEC_STATUS] = MARO_STATUS_DONE;
    decision[MARO_VM_DEC_STEP] = o->ep_step++;
    fill_metrics(o, metrics);
    return MARO_STATUS_DONE;
}

int vm_oracle_frame_words(const VmOracle* o) { return o->frame_words; }
void vm_oracle_read_frame(const VmOracle* o, int32_t* out) { export_frame(o, out); }
int vm_oracle_tick(const VmOracle* o) { return o->tick; }
void vm_oracle_counters(const VmOracle* o, int64_t* out) { out[0] = o->n_steps; out[1] = o->n_ticks; out[2] = o->n_events; out[3] = o->n_snapshots; }
int vm_oracle_read_snapshot(const VmOracle* o, int frame_index, int32_t* out) {
    if (frame_index < 0) return 0;
    int row = frame_index % o->ring_rows;
    if (o->ring_frame[row] != frame_index) return 0;
    memcpy(out, o->ring + (size_t)row * o->frame_words, 4 * (size_t)o->frame_words);
    return 1;
}

/* best fit (examples/vm_scheduling/rule_based_algorithm/best_fit.py:27-64, metric "remaining_cpu_cores"): the valid PM
 * with the fewest remaining cpu cores in the pre-decision snapshot (= the live frame), first minimum wins */
void vm_policy_best_fit(const VmOracle* o, const int32_t* dec, int32_t* action) {
    int n = dec[MARO_VM_DEC_N_VALID], best = -1, best_rem = 0;
    for (int k = 0; k < n; k++) {
        int p = dec[MARO_VM_DEC_HEAD + k];
        int rem = o->pms[p].cpu_cap - o->pms[p].cpu_alloc;
        if (best < 0 || rem < best_rem) { best = p; best_rem = rem; }
    }
    action[0] = dec[MARO_VM_DEC_VM_ID]; action[1] = MARO_VM_ACTION_ALLOCATE; action[2] = best; action[3] = 0;
}
int64_t vm_oracle_run_episode(VmOracle* o, int policy, int64_t* final_metrics) {
    int32_t* dec = (int32_t*)malloc(sizeof(int32_t) * (size_t)vm_oracle_decision_words(o));
    int32_t act[4];
    int64_t met[MARO_VM_METRIC_WORDS], steps = 0;
    int st = vm_oracle_step(o, NULL, 0, dec, met);
    while (st == MARO_STATUS_DECISION) {
        steps++;
        if (policy == 1) { vm_policy_best_fit(o, dec, act); st = vm_oracle_step(o, act, 1, dec, met); }
        else { act[0] = dec[MARO_VM_DEC_VM_ID]; act[1] = 0; act[2] = dec[MARO_VM_DEC_HEAD]; act[3] = 0; st = vm_oracle_step(o, act, 1, dec, met); }
    }
    steps++;
    if (final_metrics) memcpy(final_metrics, met, sizeof(met));
    free(dec);
    return steps;
}
