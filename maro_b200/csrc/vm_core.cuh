// vm_core.cuh — per-replica vm_scheduling simulation core (device code; one lane group = one replica).
//
// From-scratch formulation of the reference's Env.step for the vm_scheduling scenario
// (maro/simulator/core.py:317-381, maro/simulator/scenarios/vm_scheduling/business_engine.py:449-525, 575-905,
//  physical_machine.py:54-63, virtual_machine.py:60-90).
//
// What the reference does with Python dicts and per-tick event lists is restated around three flat structures per
// replica, all resident in HBM / L2 (a replica of azure.2019.10k is ~70 KB, too large for shared memory at a useful
// occupancy):
//
//   * PM lists.  `pm.live_vms` + `_live_vms[vm_id]` become two parallel arrays of 16-byte entries per PM, kept in
//     allocation order and laid out [slot][pm] so that the lanes of a group (lane = pm mod G) read consecutive entries:
//     hot {series base, deletion tick, cores | request->creation delay << 16, utilisation (float32 bits)} is everything
//     the per-tick sweep needs (no per-VM table look-up), cold {vm, creation tick, memory, -} is read when a VM leaves.  The per-tick sweep (_process_finished_vm, _update_vm_workload,
//     _update_pm_workload: :575-592, :640-652, :770-781) is ONE pass over these lists: each lane walks its PMs' lists
//     in order, drops entries whose deletion tick is now, refreshes the utilisation from the trace and accumulates
//     `cpu_utilization * cores` in list order (the order fixes the float64 sum, hence round(x, 2), hence decisions).
//   * Request stream.  The REQUEST events of a tick are (a) requests postponed `delay_duration` ticks ago, then
//     (b) the trace's new requests.  Because the delay is a constant, (a) is a FIFO ordered by due tick: a ring of
//     {vm, remaining buffer time, due tick} replaces the event linked lists.  PENDING_DECISION / TAKE_ACTION are
//     immediate children of the request being executed, i.e. the kernel's return and resume points.
//   * Frame.  The canonical frame words (attr-major, alphabetical attributes; float attributes as float32 — what a
//     snapshot query returns, np_backend.pyx:547-560) live at the head of the replica block; cpu_utilization is kept
//     exactly as the integer k = round(100 x) next to it, energy comes from a host-built table energy[pm_type][k]
//     (:671-688 evaluated by the host libm for every k in 0..10000), so all float64 arithmetic that decides anything
//     is +,*,/ only and bit-identical to the CPU statement.
//
// total_incomes adds the unit price of every live VM each tick in dict order (:913-915); that float64 chain is replaced
// by a maintained sum of live prices (one add per tick).  It is a reporting metric only (no decision reads it) and
// agrees to ~1e-12 relative; everything else is bit-exact.
#pragma once
#include "cim_core.cuh"  // lane-group primitives

namespace maro {

enum VmPmAttr { VPA_CLUSTER, VPA_CPU_ALLOC, VPA_CPU_CAP, VPA_CPU_UTIL, VPA_DC, VPA_ENERGY, VPA_ID, VPA_MEM_ALLOC, VPA_MEM_CAP,
                VPA_OVERSUB, VPA_PM_TYPE, VPA_RACK, VPA_REGION, VPA_ZONE, VPA_COUNT };

enum VmCtrl {
    VC_STATE, VC_TICK, VC_EP_STEP, VC_CUR_VM, VC_CUR_BUDGET, VC_FIFO_HEAD, VC_FIFO_COUNT, VC_REQ_CUR,
    VC_REQ_END, VC_N_LIVE, VC_ERR, VC_PAD0,
    // 64-bit values from here on (even index)
    VC_NSTEPS, VC_NTICKS = VC_NSTEPS + 2, VC_NEVENTS = VC_NTICKS + 2, VC_NSNAPS = VC_NEVENTS + 2,
    VC_M_REQ = VC_NSNAPS + 2, VC_M_SUCC_ALLOC = VC_M_REQ + 2, VC_M_SUCC_COMP = VC_M_SUCC_ALLOC + 2,
    VC_M_FAIL_ALLOC = VC_M_SUCC_COMP + 2, VC_M_FAIL_COMP = VC_M_FAIL_ALLOC + 2, VC_M_LAT_AGENT = VC_M_FAIL_COMP + 2,
    VC_M_LAT_RES = VC_M_LAT_AGENT + 2, VC_M_OVERSUB = VC_M_LAT_RES + 2, VC_M_OVL_PMS = VC_M_OVERSUB + 2,
    VC_M_OVL_VMS = VC_M_OVL_PMS + 2,
    VC_D_INCOMES = VC_M_OVL_VMS + 2, VC_D_ENERGY_COST = VC_D_INCOMES + 2, VC_D_ENERGY = VC_D_ENERGY_COST + 2,
    VC_D_PROFIT = VC_D_ENERGY + 2, VC_D_LIVE_PRICE = VC_D_PROFIT + 2,
    VC_COUNT = VC_D_LIVE_PRICE + 2
};
enum { VM_ST_START = 0, VM_ST_AWAIT = 1, VM_ST_DONE = 2, VM_ST_FINISHED = 3 };
enum { VM_UTIL_STEPS = 10000 };  // energy table rows: k = 0..10000 (cpu_utilization 0.00 .. 100.00)

struct VmShape {
    int N, R, C, D, Z, RG, T, n_vm;
    int max_tick, start_tick, snap_res, ring_rows, delay, budget, kill_all, max_actions, n_replicas;
    double max_cpu_over, max_mem_over, max_util_rate, unit_energy_price, pue;
    int FW, FWp, CWp, K, FQ, SW, DW;
    int o_rack, o_cluster, o_dc, o_zone, o_region;  // frame word offsets of the upper node types
    // table blob (int32 word offsets; doubles 8-byte aligned)
    int t_frame0, t_energy, t_rack_range, t_cluster_range, t_dc_range, t_zone_range, t_region_range;
    int t_rec0, t_rec1, t_price, t_req_offset, t_val, t_has;
};

struct VmReplica {
    int32_t* f;    // frame words
    int32_t* c;    // ctrl
    int32_t* uk;   // [N] cpu_utilization * 100 as an integer
    int32_t* len;  // [N] PM list lengths
    int32_t* q;    // [FQ][4] postponed requests {vm, remaining buffer time, due tick, 0}
    int32_t* l;    // [K][N][4] PM lists, hot part  {series offset - creation, deletion, cores | (creation - request) << 16, util bits}
    int32_t* lc;   // [K][N][4] PM lists, cold part {vm, creation, memory, 0}
    const int32_t* t;
    int32_t* snap;
    int32_t* snap_frame;
};

// pointers of replica `i` inside the state / snapshot arrays
MARO_DEV VmReplica vm_replica_at(const VmShape& s, int32_t* state, const int32_t* tables, int32_t* snap, int32_t* snap_frame, size_t i) {
    VmReplica r;
    int32_t* base = state + i * (size_t)s.SW;
    r.f = base;
    r.c = base + s.FWp;
    r.uk = r.c + s.CWp;
    r.len = r.uk + s.N;
    r.q = r.uk + ((2 * s.N + 3) & ~3);
    r.l = r.q + 4 * (size_t)s.FQ;
    r.lc = r.l + 4 * (size_t)s.K * s.N;
    r.t = tables;
    r.snap = snap + i * (size_t)s.ring_rows * s.FWp;
    r.snap_frame = snap_frame + i * (size_t)s.ring_rows;
    return r;
}

struct I4 { int32_t x, y, z, w; };
#ifdef MARO_HOST_EMULATION
static inline I4 ld4(const int32_t* p) { I4 v; memcpy(&v, p, 16); return v; }
static inline void st4(int32_t* p, I4 v) { memcpy(p, &v, 16); }
static inline I4 ld4_ro(const int32_t* p) { return ld4(p); }
static inline double maro_rint(double x) { return rint(x); }
#else
__device__ __forceinline__ I4 ld4(const int32_t* p) { int4 v = *reinterpret_cast<const int4*>(p); return {v.x, v.y, v.z, v.w}; }
__device__ __forceinline__ void st4(int32_t* p, I4 v) { *reinterpret_cast<int4*>(p) = make_int4(v.x, v.y, v.z, v.w); }
__device__ __forceinline__ I4 ld4_ro(const int32_t* p) { int4 v = __ldg(reinterpret_cast<const int4*>(p)); return {v.x, v.y, v.z, v.w}; }
__device__ __forceinline__ double maro_rint(double x) { return rint(x); }
#endif

MARO_DEV int32_t& VPM(const VmShape& s, const VmReplica& r, int attr, int p) { return r.f[attr * s.N + p]; }
MARO_DEV int64_t vctrl_get64(const VmReplica& r, int i) { return *reinterpret_cast<const int64_t*>(r.c + i); }
MARO_DEV void vctrl_add64(const VmReplica& r, int i, int64_t d) { *reinterpret_cast<int64_t*>(r.c + i) += d; }
MARO_DEV double& vctrl_f64(const VmReplica& r, int i) { return *reinterpret_cast<double*>(r.c + i); }
MARO_DEV const double* vm_tab_f64(const VmReplica& r, int off) { return reinterpret_cast<const double*>(r.t + off); }

// sum of one double per lane with a fixed (butterfly) association — deterministic for a given state
template <int G>
MARO_DEV double sum_f64(const Grp<G>& g, double x) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        int64_t b;
        memcpy(&b, &x, 8);
        int lo = g.shfl((int)(uint32_t)((uint64_t)b & 0xffffffffu), g.lane ^ o);
        int hi = g.shfl((int)(uint32_t)((uint64_t)b >> 32), g.lane ^ o);
        int64_t ob = (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
        double y;
        memcpy(&y, &ob, 8);
        // both partners must add in the same operand order to stay in agreement
        x = (g.lane & o) ? y + x : x + y;
    }
    return x;
}

// PhysicalMachine.update_cpu_utilization (physical_machine.py:54-63): round(max(0, x), 2) of an np.float64 ->
// k = rint(100 x); the attribute value is k / 100.0
MARO_DEV int util_to_k(double x) { return x > 0 ? (int)maro_rint(x * 100.0) : 0; }
MARO_DEV double energy_of_k(const VmShape& s, const VmReplica& r, int pm_type, int k) {
    if (k > VM_UTIL_STEPS) k = VM_UTIL_STEPS;
    return vm_tab_f64(r, s.t_energy)[pm_type * (VM_UTIL_STEPS + 1) + k];
}
MARO_DEV void pm_store_util(const VmShape& s, const VmReplica& r, int p, int k) {
    r.uk[p] = k;
    VPM(s, r, VPA_CPU_UTIL, p) = maro_f2i(maro_d2f((double)k / 100.0));
    VPM(s, r, VPA_ENERGY, p) = maro_f2i(maro_d2f(energy_of_k(s, r, VPM(s, r, VPA_PM_TYPE, p), k)));
}

MARO_DEV int vm_frame_index(const VmShape& s, int tick) { return (tick - s.start_tick) / s.snap_res; }

// Snapshot rows hold the whole frame, but the static attributes are written once per row at reset (vm_replica_reset):
// take_snapshot copies only the words that can change — five PM attributes and the empty-machine counts.
template <int G>
MARO_DEV void vm_copy_dynamic(const VmShape& s, const Grp<G>& g, const int32_t* src, int32_t* dst) {
    const int N = s.N;
    for (int i = g.lane; i < N; i += G) {
        dst[VPA_CPU_ALLOC * N + i] = src[VPA_CPU_ALLOC * N + i];
        dst[VPA_CPU_UTIL * N + i] = src[VPA_CPU_UTIL * N + i];
        dst[VPA_ENERGY * N + i] = src[VPA_ENERGY * N + i];
        dst[VPA_MEM_ALLOC * N + i] = src[VPA_MEM_ALLOC * N + i];
        dst[VPA_OVERSUB * N + i] = src[VPA_OVERSUB * N + i];
    }
    for (int i = g.lane; i < s.R; i += G) dst[s.o_rack + 2 * s.R + i] = src[s.o_rack + 2 * s.R + i];
    for (int i = g.lane; i < s.C; i += G) dst[s.o_cluster + s.C + i] = src[s.o_cluster + s.C + i];
    for (int i = g.lane; i < s.D; i += G) dst[s.o_dc + i] = src[s.o_dc + i];
    for (int i = g.lane; i < s.Z; i += G) dst[s.o_zone + i] = src[s.o_zone + i];
    for (int i = g.lane; i < s.RG; i += G) dst[s.o_region + i] = src[s.o_region + i];
}

template <int G>
MARO_DEV void vm_snapshot(const VmShape& s, const Grp<G>& g, const VmReplica& r, int frame_index) {
    g.sync();
    int row = frame_index % s.ring_rows;
    vm_copy_dynamic(s, g, r.f, r.snap + (size_t)row * s.FWp);
    if (g.lane == 0) r.snap_frame[row] = frame_index;
    g.sync();
}

// _get_valid_pms (:715-768) for request `vm`: number of valid PMs; with `out` also their ids, ascending
template <int G>
MARO_DEV int vm_valid_pms(const VmShape& s, const Grp<G>& g, const VmReplica& r, int vm, int32_t* out) {
    I4 r0 = ld4_ro(r.t + s.t_rec0 + 4 * vm), r1 = ld4_ro(r.t + s.t_rec1 + 4 * vm);
    const int cores = r0.w, mem = r1.x, cat = r1.w;
    int n = 0;
    for (int base = 0; base < s.N; base += G) {
        int p = base + g.lane;
        bool ok = false;
        if (p < s.N) {
            int ov = VPM(s, r, VPA_OVERSUB, p), ca = VPM(s, r, VPA_CPU_ALLOC, p), cap = VPM(s, r, VPA_CPU_CAP, p);
            int ma = VPM(s, r, VPA_MEM_ALLOC, p), mcap = VPM(s, r, VPA_MEM_CAP, p);
            if (cat == 1 || cat == 2) {
                ok = (ov == 0 || ov == -1) && ca + cores <= cap && ma + mem <= mcap;
            } else {
                double util = (double)r.uk[p] / 100.0;
                ok = (ov == 0 || ov == 1) && (double)(ca + cores) <= s.max_cpu_over * (double)cap &&
                     (double)(ma + mem) <= s.max_mem_over * (double)mcap &&
                     util / 100 * (double)cap + (double)cores <= s.max_util_rate * (double)cap;
            }
        }
        uint32_t m = g.ballot(ok);
        if (out && ok) {
            const int at = n + maro_popc(m & ((1u << g.lane) - 1u));
            out[at] = p;
            // extension behind the id area: the PM's remaining CPU cores in the decision's (= the live) frame, what the
            // rule-based agents of the reference query the snapshot list for (rule_based_algorithm/best_fit.py:38-44)
            out[s.N + at] = VPM(s, r, VPA_CPU_CAP, p) - VPM(s, r, VPA_CPU_ALLOC, p);
        }
        n += maro_popc(m);
    }
    return n;
}

// _postpone_vm_request (:690-713); returns false when the ring is full
MARO_DEV bool vm_postpone(const VmShape& s, const VmReplica& r, int tick, bool resource, int vm, int budget,
                          int remaining_buffer_time, int& fifo_head, int& fifo_count) {
    if (remaining_buffer_time >= s.delay) {
        vctrl_add64(r, resource ? VC_M_LAT_RES : VC_M_LAT_AGENT, s.delay);
        if (fifo_count >= s.FQ) return false;
        int slot = fifo_head + fifo_count;
        if (slot >= s.FQ) slot -= s.FQ;
        st4(r.q + 4 * slot, I4{vm, budget - s.delay, tick + s.delay, 0});
        fifo_count++;
    } else {
        vctrl_add64(r, VC_M_FAIL_ALLOC, 1);
    }
    return true;
}

// _on_action_received (:828-905).  Restriction of this build: every action must name the VM of the decision being
// answered (the reference accepts any id in its pending-payload dict); anything else is MARO_STATUS_BAD_ACTION.
// Leader lane only.  Returns 0, -1 (bad action) or -2 (a PM list / the ring overflowed).
MARO_DEV int vm_on_actions(const VmShape& s, const VmReplica& r, int tick, const int32_t* act, int n_act,
                           int& fifo_head, int& fifo_count) {
    const int vm = r.c[VC_CUR_VM];
    if (n_act <= 0) return 0;  // empty action list: the pending request is dropped (:836-839)
    if (n_act > s.max_actions) n_act = s.max_actions;
    I4 r0 = ld4_ro(r.t + s.t_rec0 + 4 * vm), r1 = ld4_ro(r.t + s.t_rec1 + 4 * vm);
    for (int i = 0; i < n_act; i++) {
        I4 a = ld4(act + 4 * i);
        if (i > 0 || a.x != r1.y) return -1;  // "The VM id ... sent by agent is invalid."
        if (a.y == MARO_VM_ACTION_ALLOCATE) {
            int p = a.z;
            if (p < 0 || p >= s.N) return -1;
            int n = r.len[p];
            if (n >= s.K) return -2;
            float u = maro_i2f(r.t[s.t_val + r0.x]);  // get_utilization(cur_tick): series[0]
            const int delay = tick - r0.y;  // creation - request tick (0 .. buffer budget)
            if (delay < 0 || delay > 0x7fff || r0.w < 0 || r0.w > 0xffff) return -2;
            st4(r.l + 4 * ((size_t)n * s.N + p), I4{r0.x - tick, tick + r0.z, r0.w | (delay << 16), maro_f2i(u)});
            st4(r.lc + 4 * ((size_t)n * s.N + p), I4{vm, tick, r1.x, 0});
            r.len[p] = n + 1;
            r.c[VC_N_LIVE] += 1;
            if (VPM(s, r, VPA_OVERSUB, p) == 0) VPM(s, r, VPA_OVERSUB, p) = r1.w == 0 ? 1 : -1;
            int cap = VPM(s, r, VPA_CPU_CAP, p);
            VPM(s, r, VPA_CPU_ALLOC, p) += r0.w;
            VPM(s, r, VPA_MEM_ALLOC, p) += r1.x;
            double x = ((double)cap * ((double)r.uk[p] / 100.0) + (double)r0.w * (double)u) / (double)cap;
            pm_store_util(s, r, p, util_to_k(x));
            vctrl_add64(r, VC_M_SUCC_ALLOC, 1);
            vctrl_f64(r, VC_D_LIVE_PRICE) += vm_tab_f64(r, s.t_price)[vm];
        } else {
            int budget = r.c[VC_CUR_BUDGET];
            if (!vm_postpone(s, r, tick, false, vm, budget, budget - a.z * s.delay, fifo_head, fifo_count)) return -2;
        }
    }
    return 0;
}

// BusinessEngine.step (:449-493) minus the request insertion: finished VMs, VM / PM workloads, roll-ups
template <int G>
MARO_DEV void vm_tick_begin(const VmShape& s, const Grp<G>& g, const VmReplica& r, int tick) {
    const double* price = vm_tab_f64(r, s.t_price);
    const uint8_t* has = reinterpret_cast<const uint8_t*>(r.t + s.t_has);
    int fin = 0;
    double fin_price = 0.0;
    const int32_t* val = r.t + s.t_val;
    for (int p = g.lane; p < s.N; p += G) {
        const int n = r.len[p];
        int kept = 0, ca = 0, ma = 0;
        double used = 0.0;
        // four list slots per round: their entry loads, then their trace look-ups, are issued together (the sweep is
        // latency bound: one replica = one warp, ~14 warps per SM); the float64 sum still runs in list order
        for (int k0 = 0; k0 < n; k0 += 4) {
            I4 e[4];
            int32_t nb[4];
            uint8_t hs[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (k0 + j < n) e[j] = ld4(r.l + 4 * ((size_t)(k0 + j) * s.N + p));
#pragma unroll
            for (int j = 0; j < 4; j++) {
                hs[j] = 0; nb[j] = 0;
                if (k0 + j < n && e[j].y != tick) {
                    int idx = e[j].x + tick;  // series offset + (tick - creation)
                    hs[j] = has[idx + (e[j].z >> 16)];  // "+ (creation - request)": flags are indexed from the request tick
                    nb[j] = val[idx];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = k0 + j;
                if (k >= n) break;
                if (e[j].y == tick) {  // _process_finished_vm (:770-781)
                    I4 c = ld4(r.lc + 4 * ((size_t)k * s.N + p));
                    ca += e[j].z & 0xffff;
                    ma += c.z;
                    fin++;
                    fin_price += price[c.x];
                    continue;
                }
                bool dirty = false;
                if (hs[j]) {  // _update_vm_workload: a reading exists for this VM at this tick
                    dirty = nb[j] != e[j].w;
                    e[j].w = nb[j];
                }
                used += (double)maro_i2f(e[j].w) * (double)(e[j].z & 0xffff);
                if (kept != k) {
                    st4(r.l + 4 * ((size_t)kept * s.N + p), e[j]);
                    st4(r.lc + 4 * ((size_t)kept * s.N + p), ld4(r.lc + 4 * ((size_t)k * s.N + p)));
                } else if (dirty) {
                    r.l[4 * ((size_t)k * s.N + p) + 3] = e[j].w;
                }
                kept++;
            }
        }
        if (kept != n) {
            r.len[p] = kept;
            VPM(s, r, VPA_CPU_ALLOC, p) -= ca;
            VPM(s, r, VPA_MEM_ALLOC, p) -= ma;
            if (kept == 0) VPM(s, r, VPA_OVERSUB, p) = 0;
        }
        int k100 = util_to_k(used / (double)VPM(s, r, VPA_CPU_CAP, p));  // _update_pm_workload (:640-652)
        if (k100 != r.uk[p]) pm_store_util(s, r, p, k100);
    }
    fin = g.sum(fin);
    fin_price = sum_f64(g, fin_price);
    g.sync();
    if (g.lane == 0 && fin) {
        vctrl_add64(r, VC_M_SUCC_COMP, fin);
        int live = r.c[VC_N_LIVE] - fin;
        r.c[VC_N_LIVE] = live;
        vctrl_f64(r, VC_D_LIVE_PRICE) = live == 0 ? 0.0 : vctrl_f64(r, VC_D_LIVE_PRICE) - fin_price;
    }
    // _update_upper_level_metrics (:594-638)
    const int32_t* rr = r.t + s.t_rack_range;
    for (int i = g.lane; i < s.R; i += G) {
        int c = 0;
        for (int p = rr[2 * i]; p < rr[2 * i + 1]; p++) c += VPM(s, r, VPA_CPU_ALLOC, p) == 0;
        r.f[s.o_rack + 2 * s.R + i] = c;  // racks: cluster_id, data_center_id, empty_machine_num, ...
    }
    g.sync();
    if (g.lane == 0) {
        const int32_t* cr = r.t + s.t_cluster_range;
        for (int i = 0; i < s.C; i++) { int c = 0; for (int k = cr[2 * i]; k < cr[2 * i + 1]; k++) c += r.f[s.o_rack + 2 * s.R + k]; r.f[s.o_cluster + 1 * s.C + i] = c; }
        const int32_t* dr = r.t + s.t_dc_range;
        for (int i = 0; i < s.D; i++) { int c = 0; for (int k = dr[2 * i]; k < dr[2 * i + 1]; k++) c += r.f[s.o_cluster + 1 * s.C + k]; r.f[s.o_dc + 0 * s.D + i] = c; }
        const int32_t* zr = r.t + s.t_zone_range;
        for (int i = 0; i < s.Z; i++) { int c = 0; for (int k = zr[2 * i]; k < zr[2 * i + 1]; k++) c += r.f[s.o_dc + 0 * s.D + k]; r.f[s.o_zone + 0 * s.Z + i] = c; }
        const int32_t* gr = r.t + s.t_region_range;
        for (int i = 0; i < s.RG; i++) { int c = 0; for (int k = gr[2 * i]; k < gr[2 * i + 1]; k++) c += r.f[s.o_zone + 0 * s.Z + k]; r.f[s.o_region + 0 * s.RG + i] = c; }
    }
    g.sync();
}

// post_step (:495-525) up to the snapshot; `scratch` = 2 N doubles private to the group
template <int G>
MARO_DEV void vm_tick_end(const VmShape& s, const Grp<G>& g, const VmReplica& r, int tick, double* scratch) {
    const double* price = vm_tab_f64(r, s.t_price);
    int oversubs = 0;
    bool any_overload = false;
    for (int p = g.lane; p < s.N; p += G) {
        int ov = VPM(s, r, VPA_OVERSUB, p);
        if (ov != 0 && VPM(s, r, VPA_CPU_ALLOC, p) > VPM(s, r, VPA_CPU_CAP, p)) oversubs++;
        int k = r.uk[p];
        double e = energy_of_k(s, r, VPM(s, r, VPA_PM_TYPE, p), k);
        scratch[p] = e;
        scratch[s.N + p] = e * s.unit_energy_price * s.pue;
        any_overload |= k > VM_UTIL_STEPS;
    }
    oversubs = g.sum(oversubs);
    any_overload = g.ballot(any_overload) != 0;
    g.sync();
    if (g.lane == 0) {
        double total_energy = 0.0, total_cost = 0.0;
        for (int p = 0; p < s.N; p++) { total_energy += scratch[p]; total_cost += scratch[s.N + p]; }
        if (oversubs) vctrl_add64(r, VC_M_OVERSUB, oversubs);
        double incomes = vctrl_f64(r, VC_D_INCOMES);
        if (any_overload) {  // _overload (:654-669), PM order
            for (int p = 0; p < s.N; p++) {
                if (r.uk[p] <= VM_UTIL_STEPS) continue;
                int n = r.len[p];
                if (s.kill_all) {
                    double gone = 0.0;
                    for (int k = 0; k < n; k++) {
                        I4 c = ld4(r.lc + 4 * ((size_t)k * s.N + p));
                        incomes -= price[c.x] * (double)(tick - c.y);
                        gone += price[c.x];
                    }
                    r.len[p] = 0;
                    int live = r.c[VC_N_LIVE] - n;
                    r.c[VC_N_LIVE] = live;
                    vctrl_f64(r, VC_D_LIVE_PRICE) = live == 0 ? 0.0 : vctrl_f64(r, VC_D_LIVE_PRICE) - gone;
                    vctrl_add64(r, VC_M_FAIL_COMP, n);
                }
                vctrl_add64(r, VC_M_OVL_VMS, n);
            }
        }
        vctrl_f64(r, VC_D_ENERGY) += total_energy;
        double cost = vctrl_f64(r, VC_D_ENERGY_COST) + total_cost;
        vctrl_f64(r, VC_D_ENERGY_COST) = cost;
        incomes += vctrl_f64(r, VC_D_LIVE_PRICE);  // _update_incomes (:913-915)
        vctrl_f64(r, VC_D_INCOMES) = incomes;
        vctrl_f64(r, VC_D_PROFIT) = incomes - cost;
    }
    g.sync();
}

MARO_DEV void vm_fill_metrics(const VmReplica& r, int64_t* m) {
    m[0] = vctrl_get64(r, VC_M_REQ); m[1] = vctrl_get64(r, VC_D_INCOMES); m[2] = vctrl_get64(r, VC_D_ENERGY_COST);
    m[3] = vctrl_get64(r, VC_D_PROFIT); m[4] = vctrl_get64(r, VC_D_ENERGY); m[5] = vctrl_get64(r, VC_M_SUCC_ALLOC);
    m[6] = vctrl_get64(r, VC_M_SUCC_COMP); m[7] = vctrl_get64(r, VC_M_FAIL_ALLOC); m[8] = vctrl_get64(r, VC_M_FAIL_COMP);
    m[9] = vctrl_get64(r, VC_M_LAT_AGENT); m[10] = vctrl_get64(r, VC_M_LAT_RES); m[11] = vctrl_get64(r, VC_M_OVERSUB);
    m[12] = vctrl_get64(r, VC_M_OVL_PMS); m[13] = vctrl_get64(r, VC_M_OVL_VMS); m[14] = 0; m[15] = 0;
}

// One Env.step of one replica (core.py:92-133, 301-381).  `act` = [max_actions][4] int32 of this replica.
template <int G>
MARO_DEV void vm_replica_step(const VmShape& s, const Grp<G>& g, const VmReplica& r, const int32_t* act, int n_act,
                              int32_t* dec, int64_t* met, double* scratch) {
    int state = r.c[VC_STATE];
    g.sync();
    if (state >= VM_ST_DONE) {
        for (int i = g.lane; i < s.DW; i += G) dec[i] = i == MARO_VM_DEC_STATUS ? MARO_STATUS_FINISHED : 0;
        if (g.lane == 0) {
            r.c[VC_STATE] = VM_ST_FINISHED;
            for (int i = 0; i < MARO_VM_METRIC_WORDS; i++) met[i] = 0;
        }
        g.sync();
        return;
    }
    int tick = r.c[VC_TICK];
    int fifo_head = r.c[VC_FIFO_HEAD], fifo_count = r.c[VC_FIFO_COUNT];
    int req_cur = r.c[VC_REQ_CUR], req_end = r.c[VC_REQ_END];
    int nev = 0, nticks = 0, nsnaps = 0, err = 0;
    bool resume = state == VM_ST_AWAIT;
    if (resume) {
        if (g.lane == 0) err = vm_on_actions(s, r, tick, act, n_act, fifo_head, fifo_count);
        err = g.shfl(err, 0);
        fifo_head = g.shfl(fifo_head, 0);
        fifo_count = g.shfl(fifo_count, 0);
        nev += 2;  // the PENDING_DECISION event and its TAKE_ACTION child
        g.sync();
    }
    int status = MARO_STATUS_DONE, cur_vm = -1, cur_budget = 0;
    while (!err) {
        if (!resume) {
            vm_tick_begin(s, g, r, tick);
            nticks++;
            req_cur = r.t[s.t_req_offset + tick];
            req_end = r.t[s.t_req_offset + tick + 1];
            if (g.lane == 0 && req_end > req_cur) vctrl_add64(r, VC_M_REQ, req_end - req_cur);
        }
        resume = false;
        // EventBuffer.execute (event_buffer.py:190-247) over the tick's REQUEST events
        for (;;) {
            int vm, budget;
            I4 head = ld4(r.q + 4 * fifo_head);
            if (fifo_count > 0 && head.z == tick) {
                vm = head.x; budget = head.y;
                fifo_head = fifo_head + 1 == s.FQ ? 0 : fifo_head + 1;
                fifo_count--;
            } else if (req_cur < req_end) {
                vm = req_cur++; budget = s.budget;
            } else {
                break;
            }
            nev++;
            int n_valid = vm_valid_pms<G>(s, g, r, vm, nullptr);  // _on_vm_required (:783-826)
            if (n_valid > 0) { cur_vm = vm; cur_budget = budget; break; }
            int ok = 1;
            if (g.lane == 0) ok = vm_postpone(s, r, tick, true, vm, budget, budget, fifo_head, fifo_count);
            ok = g.shfl(ok, 0);
            fifo_count = g.shfl(fifo_count, 0);
            if (!ok) { err = -2; break; }
        }
        if (err || cur_vm >= 0) break;
        vm_tick_end(s, g, r, tick, scratch);
        if ((tick + 1) % s.snap_res == 0) { vm_snapshot(s, g, r, vm_frame_index(s, tick)); nsnaps++; }
        if (tick + 1 >= s.max_tick) break;
        tick++;
    }
    g.sync();
    if (err) {
        status = err == -2 ? MARO_STATUS_QUEUE_OVERFLOW : MARO_STATUS_BAD_ACTION;
        for (int i = g.lane; i < s.DW; i += G) dec[i] = i == MARO_VM_DEC_STATUS ? status : 0;
        if (g.lane == 0) { r.c[VC_STATE] = VM_ST_FINISHED; r.c[VC_ERR] = err; vm_fill_metrics(r, met); }
        g.sync();
        return;
    }
    if (cur_vm >= 0) {
        vm_snapshot(s, g, r, vm_frame_index(s, tick));  // core.py:345
        nsnaps++;
        int n_valid = vm_valid_pms<G>(s, g, r, cur_vm, dec + MARO_VM_DEC_HEAD);
        for (int i = n_valid + g.lane; i < s.DW - MARO_VM_DEC_HEAD; i += G)
            if (i < s.N || i >= s.N + n_valid) dec[MARO_VM_DEC_HEAD + i] = 0;  // (ids [0, n) and their extension [N, N + n) stay)
        status = MARO_STATUS_DECISION;
        if (g.lane == 0) {
            I4 r0 = ld4_ro(r.t + s.t_rec0 + 4 * cur_vm), r1 = ld4_ro(r.t + s.t_rec1 + 4 * cur_vm);
            dec[MARO_VM_DEC_TICK] = tick;
            dec[MARO_VM_DEC_VM_ID] = r1.y;
            dec[MARO_VM_DEC_FRAME_INDEX] = vm_frame_index(s, tick);
            dec[MARO_VM_DEC_CPU] = r0.w;
            dec[MARO_VM_DEC_MEMORY] = r1.x;
            dec[MARO_VM_DEC_SUB_ID] = r1.z;
            dec[MARO_VM_DEC_STATUS] = status;
            dec[MARO_VM_DEC_STEP] = r.c[VC_EP_STEP];
            dec[MARO_VM_DEC_CATEGORY] = r1.w;
            dec[MARO_VM_DEC_BUFFER_TIME] = cur_budget;
            dec[MARO_VM_DEC_N_VALID] = n_valid;
            dec[MARO_VM_DEC_EXT_OFFSET] = MARO_VM_DEC_HEAD + s.N;
            r.c[VC_CUR_VM] = cur_vm;
            r.c[VC_CUR_BUDGET] = cur_budget;
        }
    } else {
        if ((tick + 1) % s.snap_res != 0) { vm_snapshot(s, g, r, vm_frame_index(s, tick)); nsnaps++; }  // core.py:122-126
        for (int i = g.lane; i < s.DW; i += G)
            dec[i] = i == MARO_VM_DEC_TICK ? tick : (i == MARO_VM_DEC_STATUS ? status : (i == MARO_VM_DEC_STEP ? r.c[VC_EP_STEP] : 0));
    }
    g.sync();
    if (g.lane == 0) {
        r.c[VC_STATE] = cur_vm >= 0 ? VM_ST_AWAIT : VM_ST_DONE;
        r.c[VC_TICK] = tick;
        r.c[VC_EP_STEP] += 1;
        r.c[VC_FIFO_HEAD] = fifo_head; r.c[VC_FIFO_COUNT] = fifo_count;
        r.c[VC_REQ_CUR] = req_cur; r.c[VC_REQ_END] = req_end;
        vctrl_add64(r, VC_NSTEPS, 1); vctrl_add64(r, VC_NTICKS, nticks);
        vctrl_add64(r, VC_NEVENTS, nev); vctrl_add64(r, VC_NSNAPS, nsnaps);
        vm_fill_metrics(r, met);
    }
    g.sync();
}

// Env.reset (core.py:135-153) + BusinessEngine.reset (:527-563); the step / tick / event / snapshot counters persist
template <int G>
MARO_DEV void vm_replica_reset(const VmShape& s, const Grp<G>& g, const VmReplica& r, bool init_ring) {
    for (int i = g.lane; i < s.FWp; i += G) r.f[i] = i < s.FW ? r.t[s.t_frame0 + i] : 0;
    for (int i = g.lane; i < s.CWp; i += G)
        if (i < VC_NSTEPS || i >= VC_M_REQ) r.c[i] = 0;
    for (int i = g.lane; i < s.N; i += G) { r.uk[i] = 0; r.len[i] = 0; }
    for (int i = g.lane; i < s.ring_rows; i += G) r.snap_frame[i] = -1;
    // ring rows: the static attributes are written once, when the handle is created (snapshots copy only the dynamic words)
    if (init_ring)
        for (int row = 0; row < s.ring_rows; row++)
            for (int i = g.lane * 4; i < s.FWp; i += G * 4) st4(r.snap + (size_t)row * s.FWp + i, ld4(r.t + s.t_frame0 + i));
    g.sync();
    if (g.lane == 0) r.c[VC_TICK] = s.start_tick;
    g.sync();
}

}  // namespace maro
