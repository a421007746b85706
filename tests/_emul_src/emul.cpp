// TEST INFRASTRUCTURE ONLY — compiles maro_b200/csrc/cim_core.cuh for the host: every lane of a lane group is a
// host thread (warp_emul.hpp), so the cooperative shuffle / ballot / match / atomic phases run as written.  Lets the
// kernel's per-replica logic be debugged against the oracle without a GPU.  Never loaded by the package.
// Build: g++ -O1 -g -pthread -ffp-contract=off -DMARO_HOST_EMULATION -I. -shared -fPIC emul.cpp -o ../_emul/libmaro_emul.so
#define MARO_HOST_EMULATION 1
#include "../../maro_b200/csrc/cim_host.hpp"

using namespace maro;

struct Emul {
    CimShape s;
    std::vector<int32_t> tables, state, snap, snap_frame, topo;
    std::vector<uint32_t> mt;
    int mt_words = 0, B = 0, lanes = 8;
};

static Replica rep_of(Emul* e, int i) {
    Replica r;
    int32_t* st = e->state.data() + (size_t)i * e->s.SW;
    r.f = st; r.c = st + e->s.FWp; r.q = st + e->s.FWp + e->s.CWp;
    r.t = e->tables.data() + (size_t)e->topo[i] * e->s.table_words;
    r.mt = e->mt_words ? e->mt.data() + (size_t)i * e->mt_words : nullptr;
    r.snap = e->snap.data() + (size_t)i * e->s.ring_rows * e->s.FWp;
    r.snap_frame = e->snap_frame.data() + (size_t)i * e->s.ring_rows;
    return r;
}

template <int G>
static void reset_g(Emul* e, int i) {
    Replica r = rep_of(e, i);
    wemu::run_group(G, [&](int lane) { replica_reset<G>(e->s, Grp<G>(lane), r); });
}
template <int G>
static void step_g(Emul* e, int i, const int32_t* actp, int n, int32_t* dec, int64_t* met) {
    Replica r = rep_of(e, i);
    if (n > G) n = G;
    wemu::run_group(G, [&](int lane) {
        Act4 act = {0, 0, 0, 0};
        if (lane < n) { act.v = actp[4 * lane]; act.p = actp[4 * lane + 1]; act.qty = actp[4 * lane + 2]; act.type = actp[4 * lane + 3]; }
        if (e->s.order_table && !e->s.buffer_noise) replica_step<G, false>(e->s, Grp<G>(lane), r, act, n, dec, met);
        else replica_step<G, true>(e->s, Grp<G>(lane), r, act, n, dec, met); });
}
static void reset_one(Emul* e, int i) {
    switch (e->lanes) { case 1: reset_g<1>(e, i); break; case 8: reset_g<8>(e, i); break; case 16: reset_g<16>(e, i); break; default: reset_g<32>(e, i); }
}

extern "C" {
Emul* emul_create(const MaroCimTopology* topos, int n_topos, const MaroCimConfig* cfg, int lanes) {
    Emul* e = new Emul();
    int ms, mt, md;
    if (compute_shape_and_tables(topos, n_topos, cfg, e->s, e->tables, ms, mt, md)) { delete e; return nullptr; }
    e->B = cfg->n_replicas;
    e->lanes = lanes > 0 ? lanes : lanes_per_replica(e->s);
    e->state.assign((size_t)e->B * e->s.SW, 0);
    e->snap.assign((size_t)e->B * e->s.ring_rows * e->s.FWp, 0);
    e->snap_frame.assign((size_t)e->B * e->s.ring_rows, -1);
    e->topo.assign(e->B, 0);
    if (cfg->replica_topology) for (int i = 0; i < e->B; i++) e->topo[i] = cfg->replica_topology[i];
    if (e->s.order_noise || e->s.buffer_noise) {
        e->mt_words = mt_block_words(e->s);
        e->mt.assign((size_t)e->B * e->mt_words, 0);
    }
    for (int i = 0; i < e->B; i++) reset_one(e, i);
    return e;
}
void emul_destroy(Emul* e) { delete e; }
void emul_reset(Emul* e) { for (int i = 0; i < e->B; i++) reset_one(e, i); }
void emul_step(Emul* e, const int32_t* actions, const int32_t* n_actions, int32_t* decisions, int64_t* metrics) {
    for (int i = 0; i < e->B; i++) {
        int n = actions ? (n_actions ? n_actions[i] : 1) : 0;
        const int32_t* act = actions ? actions + (size_t)i * e->s.max_actions * 4 : nullptr;
        switch (e->lanes) {
            case 1: step_g<1>(e, i, act, n, decisions + (size_t)i * e->s.DW, metrics + i * 3); break;
            case 8: step_g<8>(e, i, act, n, decisions + (size_t)i * e->s.DW, metrics + i * 3); break;
            case 16: step_g<16>(e, i, act, n, decisions + (size_t)i * e->s.DW, metrics + i * 3); break;
            default: step_g<32>(e, i, act, n, decisions + (size_t)i * e->s.DW, metrics + i * 3); break;
        }
    }
}
int emul_frame_words(Emul* e) { return e->s.FW; }
int emul_dec_words(Emul* e) { return e->s.DW; }
int emul_max_actions(Emul* e) { return e->s.max_actions; }
int emul_lanes(Emul* e) { return e->lanes; }
void emul_read_frame(Emul* e, int rep, int32_t* out) { memcpy(out, e->state.data() + (size_t)rep * e->s.SW, 4 * e->s.FW); }
int emul_read_snapshot(Emul* e, int rep, int frame, int32_t* out) {
    int row = frame % e->s.ring_rows;
    if (frame < 0 || e->snap_frame[(size_t)rep * e->s.ring_rows + row] != frame) return 0;
    memcpy(out, e->snap.data() + ((size_t)rep * e->s.ring_rows + row) * e->s.FWp, 4 * e->s.FW);
    return 1;
}
int emul_tick(Emul* e, int rep) { return e->state[(size_t)rep * e->s.SW + e->s.FWp + C_TICK]; }
#ifdef MARO_TRACK_QPEAK
int maro_emul_qpeak = 0;
#endif
void emul_counters(Emul* e, int rep, int64_t* out) { memcpy(out, e->state.data() + (size_t)rep * e->s.SW + e->s.FWp + C_NSTEPS_LO, 32); }
}

// ------------------------------------------------------------------------------------------------ citi_bike
#include "../../maro_b200/csrc/bike_host.hpp"

struct BikeEmul {
    BikeShape s;
    std::vector<int32_t> tables, state, snap, snap_frame;
    std::vector<uint32_t> rng;
    std::vector<int64_t> seeds;  // per-replica transfer seeds (-1: the topology's)
    int B = 0, lanes = 8;
};
static BikeReplica bike_rep_of(BikeEmul* e, int i) {
    BikeReplica r;
    int32_t* st = e->state.data() + (size_t)i * e->s.SW;
    r.f = st; r.c = st + e->s.FWp; r.q = st + e->s.FWp + e->s.CWp;
    r.t = e->tables.data();
    r.rng = e->rng.data() + (size_t)i * e->s.rng_words;
    r.seed = e->seeds.empty() ? -1 : e->seeds[i];
    r.snap = e->snap.data() + (size_t)i * e->s.ring_rows * e->s.FWp;
    r.snap_frame = e->snap_frame.data() + (size_t)i * e->s.ring_rows;
    return r;
}
template <int G>
static void bike_reset_g(BikeEmul* e, int i) {
    BikeReplica r = bike_rep_of(e, i);
    wemu::run_group(G, [&](int lane) { bike_replica_reset<G>(e->s, Grp<G>(lane), r); });
}
template <int G>
static void bike_step_g(BikeEmul* e, int i, const int32_t* actp, int n, int32_t* dec, int64_t* met) {
    BikeReplica r = bike_rep_of(e, i);
    if (n > G) n = G;
    wemu::run_group(G, [&](int lane) {
        Act4 act = {0, 0, 0, 0};
        if (lane < n) { act.v = actp[4 * lane]; act.p = actp[4 * lane + 1]; act.qty = actp[4 * lane + 2]; act.type = actp[4 * lane + 3]; }
        bike_replica_step<G>(e->s, Grp<G>(lane), r, act, n, dec, met);
    });
}
extern "C" {
BikeEmul* bike_emul_create(const MaroBikeTopology* topo, const MaroCimConfig* cfg, int lanes) {
    BikeEmul* e = new BikeEmul();
    if (bike_compute_shape_and_tables(*topo, cfg, e->s, e->tables)) { delete e; return nullptr; }
    e->B = cfg->n_replicas;
    e->lanes = lanes > 0 ? lanes : bike_lanes_per_replica(e->s);
    e->state.assign((size_t)e->B * e->s.SW, 0);
    e->snap.assign((size_t)e->B * e->s.ring_rows * e->s.FWp, 0);
    e->snap_frame.assign((size_t)e->B * e->s.ring_rows, -1);
    e->rng.assign((size_t)e->B * e->s.rng_words, 0);
    for (int i = 0; i < e->B; i++) { if (e->lanes == 1) bike_reset_g<1>(e, i); else if (e->lanes == 8) bike_reset_g<8>(e, i); else bike_reset_g<32>(e, i); }
    return e;
}
void bike_emul_destroy(BikeEmul* e) { delete e; }
// maro_bike_set_transfer_seeds + reset of replica `rep`
void bike_emul_reseed(BikeEmul* e, int rep, int64_t seed) {
    if (e->seeds.empty()) e->seeds.assign(e->B, -1);
    e->seeds[rep] = seed;
    if (e->lanes == 1) bike_reset_g<1>(e, rep); else if (e->lanes == 8) bike_reset_g<8>(e, rep); else bike_reset_g<32>(e, rep);
}
int bike_emul_dec_words(BikeEmul* e) { return e->s.DW; }
int bike_emul_frame_words(BikeEmul* e) { return e->s.FW; }
void bike_emul_step(BikeEmul* e, const int32_t* actions, const int32_t* n_actions, int32_t* decisions, int64_t* metrics) {
    for (int i = 0; i < e->B; i++) {
        int n = actions ? (n_actions ? n_actions[i] : 1) : 0;
        const int32_t* act = actions ? actions + (size_t)i * e->s.max_actions * 4 : nullptr;
        if (e->lanes == 1) bike_step_g<1>(e, i, act, n, decisions + (size_t)i * e->s.DW, metrics + i * 3);
        else if (e->lanes == 8) bike_step_g<8>(e, i, act, n, decisions + (size_t)i * e->s.DW, metrics + i * 3);
        else bike_step_g<32>(e, i, act, n, decisions + (size_t)i * e->s.DW, metrics + i * 3);
    }
}
void bike_emul_read_frame(BikeEmul* e, int rep, int32_t* out) { memcpy(out, e->state.data() + (size_t)rep * e->s.SW, 4 * e->s.FW); }
int bike_emul_read_snapshot(BikeEmul* e, int rep, int frame, int32_t* out) {
    int row = frame % e->s.ring_rows;
    if (frame < 0 || e->snap_frame[(size_t)rep * e->s.ring_rows + row] != frame) return 0;
    memcpy(out, e->snap.data() + ((size_t)rep * e->s.ring_rows + row) * e->s.FWp, 4 * e->s.FW);
    return 1;
}
int bike_emul_tick(BikeEmul* e, int rep) { return e->state[(size_t)rep * e->s.SW + e->s.FWp + BC_TICK]; }
void bike_emul_counters(BikeEmul* e, int rep, int64_t* out) { memcpy(out, e->state.data() + (size_t)rep * e->s.SW + e->s.FWp + BC_NSTEPS_LO, 32); }
}

// ------------------------------------------------------------------------------------------------ vm_scheduling
#include "../../maro_b200/csrc/vm_host.hpp"

struct VmEmul {
    VmShape s;
    std::vector<int32_t> tables, state, snap, snap_frame;
    std::vector<double> scratch;
    int B = 0, lanes = 32;
    std::string err;
};
template <int G>
static void vm_reset_g(VmEmul* e, int i, bool init_ring = false) {
    VmReplica r = vm_replica_at(e->s, e->state.data(), e->tables.data(), e->snap.data(), e->snap_frame.data(), i);
    wemu::run_group(G, [&](int lane) { vm_replica_reset<G>(e->s, Grp<G>(lane), r, init_ring); });
}
template <int G>
static void vm_step_g(VmEmul* e, int i, const int32_t* actp, int n, int32_t* dec, int64_t* met) {
    VmReplica r = vm_replica_at(e->s, e->state.data(), e->tables.data(), e->snap.data(), e->snap_frame.data(), i);
    wemu::run_group(G, [&](int lane) { vm_replica_step<G>(e->s, Grp<G>(lane), r, actp, n, dec, met, e->scratch.data()); });
}
extern "C" {
VmEmul* vm_emul_create(const MaroVmTopology* topo, const MaroCimConfig* cfg, int lanes) {
    VmEmul* e = new VmEmul();
    e->err = vm_compute_shape_and_tables(*topo, cfg, e->s, e->tables);
    if (!e->err.empty()) { fprintf(stderr, "vm_emul_create: %s\n", e->err.c_str()); delete e; return nullptr; }
    e->B = cfg->n_replicas;
    e->lanes = lanes > 0 ? lanes : 32;
    e->state.assign((size_t)e->B * e->s.SW, 0);
    e->snap.assign((size_t)e->B * e->s.ring_rows * e->s.FWp, 0);
    e->snap_frame.assign((size_t)e->B * e->s.ring_rows, -1);
    e->scratch.assign(2 * (size_t)e->s.N, 0.0);
    for (int i = 0; i < e->B; i++) { if (e->lanes == 1) vm_reset_g<1>(e, i, true); else if (e->lanes == 8) vm_reset_g<8>(e, i, true); else vm_reset_g<32>(e, i, true); }
    return e;
}
void vm_emul_destroy(VmEmul* e) { delete e; }
int vm_emul_dec_words(VmEmul* e) { return e->s.DW; }
int vm_emul_frame_words(VmEmul* e) { return e->s.FW; }
void vm_emul_reset(VmEmul* e) {
    for (int i = 0; i < e->B; i++) { if (e->lanes == 1) vm_reset_g<1>(e, i); else if (e->lanes == 8) vm_reset_g<8>(e, i); else vm_reset_g<32>(e, i); }
}
void vm_emul_step(VmEmul* e, const int32_t* actions, const int32_t* n_actions, int32_t* decisions, int64_t* metrics) {
    for (int i = 0; i < e->B; i++) {
        int n = actions ? (n_actions ? n_actions[i] : 1) : 0;
        const int32_t* act = actions ? actions + (size_t)i * e->s.max_actions * 4 : nullptr;
        int32_t* dec = decisions + (size_t)i * e->s.DW;
        int64_t* met = metrics + (size_t)i * MARO_VM_METRIC_WORDS;
        if (e->lanes == 1) vm_step_g<1>(e, i, act, n, dec, met);
        else if (e->lanes == 8) vm_step_g<8>(e, i, act, n, dec, met);
        else vm_step_g<32>(e, i, act, n, dec, met);
    }
}
void vm_emul_read_frame(VmEmul* e, int rep, int32_t* out) { memcpy(out, e->state.data() + (size_t)rep * e->s.SW, 4 * e->s.FW); }
int vm_emul_read_snapshot(VmEmul* e, int rep, int frame, int32_t* out) {
    int row = frame % e->s.ring_rows;
    if (frame < 0 || e->snap_frame[(size_t)rep * e->s.ring_rows + row] != frame) return 0;
    memcpy(out, e->snap.data() + ((size_t)rep * e->s.ring_rows + row) * e->s.FWp, 4 * e->s.FW);
    return 1;
}
int vm_emul_tick(VmEmul* e, int rep) { return e->state[(size_t)rep * e->s.SW + e->s.FWp + VC_TICK]; }
void vm_emul_counters(VmEmul* e, int rep, int64_t* out) { memcpy(out, e->state.data() + (size_t)rep * e->s.SW + e->s.FWp + VC_NSTEPS, 32); }
}
