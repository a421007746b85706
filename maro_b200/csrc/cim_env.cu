// cim_env.cu — kernels + C ABI (include/maro_b200.h) of the batched CIM discrete-event core for sm_100a.
//
// Kernels
//   cim_step_kernel    one warp = one replica.  The replica's state block (frame | control | event queue) is
//                      staged HBM -> shared memory with one TMA bulk copy (cp.async.bulk + mbarrier), the step
//                      runs out of shared memory (cim_core.cuh), snapshot rows stream to the ring with 128-bit
//                      coalesced stores, and the block is written back with 128-bit stores.
//   cim_reset_kernel   Env.reset for masked replicas.
//   cim_query_kernel   snapshot_list[node][ticks:nodes:attrs] gather -> float64 (env_common.cuh, shared by the scenarios).
//   cim_policy_kernel  hashed random agent (bench helper).
//   cim_rl_*_kernel    RL state / action / reward shaping over the snapshot ring.
// The citi_bike and vm_scheduling scenarios are bike_env.cu / vm_env.cu (same handle layout, env_common.cuh).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
//   (-fmad=false: CPython never contracts a*b+c; order generation must round like the reference.)
#include "env_common.cuh"

// =====================================================================================================
// Kernels
// =====================================================================================================
struct StepArgs {
    int32_t* state;        // [B][SW]
    int32_t* snap;         // [B][ring][FWp]
    int32_t* snap_frame;   // [B][ring]
    uint32_t* mt;          // [B][MTW] or nullptr
    const int32_t* tables; // [K][table_words]
    const int32_t* replica_topology;  // [B]
    const uint8_t* active;            // [B] or nullptr
    const int32_t* actions;           // [B][A][4] or nullptr
    const int32_t* n_actions;         // [B] or nullptr
    int32_t* decisions;               // [B][8]
    int64_t* metrics;                 // [B][3]
    uint8_t* light;                   // [B] 1: the next step only applies an action and yields the tick's next decision
    int mt_words;
};

__device__ __forceinline__ Replica make_replica(const CimShape& s, const StepArgs& a, int rep, int32_t* st) {
    Replica r;
    r.f = st;
    r.c = st + s.FWp;
    r.q = st + s.FWp + s.CWp;
    r.t = a.tables + (int64_t)a.replica_topology[rep] * s.table_words;
    r.mt = a.mt ? a.mt + (int64_t)rep * a.mt_words : nullptr;
    r.snap = a.snap + (int64_t)rep * s.ring_rows * s.FWp;
    r.snap_frame = a.snap_frame + (int64_t)rep * s.ring_rows;
    return r;
}

// kSpread (small batches, G < 32): one replica per WARP, only its first G lanes work.  Packing 32/G replicas into a warp
// makes the warp issue the union of their control paths; with fewer replicas than the GPU has warp slots it is faster to
// give every replica its own warp (same lane-group code, the other lanes exit).
template <int kWarps, int G, bool kGeneral, bool kSpread = false>
__global__ void __launch_bounds__(kWarps * 32, kGeneral ? 1 : 32 / kWarps) cim_step_kernel(const __grid_constant__ CimShape s,
                                                               const __grid_constant__ StepArgs a) {
    constexpr int kGroups = kSpread ? kWarps : kWarps * 32 / G;  // replicas in flight per CTA
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // one mbarrier per lane group (first 256 B)
    if (kSpread && (threadIdx.x & 31) >= G) return;
    const int gid = kSpread ? threadIdx.x >> 5 : threadIdx.x / G;
    const Grp<G> g(threadIdx.x & 31);
    int32_t* st = reinterpret_cast<int32_t*>(smem_raw + 256) + (size_t)gid * s.SW;
    uint64_t* bar = bars + gid;
    if (g.lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    g.sync();
    uint32_t phase = 0;
    for (int rep = blockIdx.x * kGroups + gid; rep < s.n_replicas; rep += gridDim.x * kGroups) {
        if (a.active && !a.active[rep]) {
            if (g.lane == 0) a.decisions[rep * 8 + 6] = MARO_STATUS_INACTIVE;
            continue;
        }
        int32_t* gstate = a.state + (int64_t)rep * s.SW;
        // ---- stage in: one TMA bulk copy, completion on the group's mbarrier.  A step that is known to stay inside
        // the current tick (another decision of the same tick is pending: it applies the action, snapshots and returns)
        // never touches the event queue, so only [frame | control] travels, both ways.
        const bool light = a.light[rep] != 0;
        const uint32_t bytes = light ? (uint32_t)(s.FWp + s.CWp) * 4u : (uint32_t)s.SW * 4u;
        if (g.lane == 0) {
            fence_proxy_async();  // order earlier generic-proxy accesses to this smem before the async write
            mbar_expect_tx(bar, bytes);
            bulk_g2s(st, gstate, bytes, bar);
        }
        while (!mbar_try_wait(bar, phase)) {}
        phase ^= 1u;
        Replica r = make_replica(s, a, rep, st);
        const int n_act = a.actions ? (a.n_actions ? min(max(a.n_actions[rep], 0), min(s.max_actions, G)) : 1) : 0;
        Act4 act = {0, 0, 0, 0};
        if (g.lane < n_act) {  // lane k fetches action k with one 128-bit load (actions may live in mapped host memory)
            int4 v = reinterpret_cast<const int4*>(a.actions + (int64_t)rep * s.max_actions * 4)[g.lane];
            act.v = v.x; act.p = v.y; act.qty = v.z; act.type = v.w;
        }
        replica_step<G, kGeneral>(s, g, r, act, n_act, a.decisions + (int64_t)rep * 8, a.metrics + (int64_t)rep * 3);
        // ---- write back (128-bit coalesced) what this step could have changed
        const int4* src4 = reinterpret_cast<const int4*>(st);
        int4* dst4 = reinterpret_cast<int4*>(gstate);
        const int n4 = (int)(bytes >> 4);
        for (int i = g.lane; i < n4; i += G) dst4[i] = src4[i];
        if (g.lane == 0) {  // hint for the next step: awaiting an action with another arrival of this tick still to decide
            const int32_t* c = st + s.FWp;
            const uint64_t arr = ((uint64_t)(uint32_t)c[C_ARR_HI] << 32) | (uint32_t)c[C_ARR_LO];
            const int dp = c[C_DEC_POS];
            a.light[rep] = (c[C_STATE] == ST_AWAIT && dp < 64 && (arr >> dp) != 0) ? 1 : 0;
        }
        g.sync();
    }
}

__global__ void cim_reset_kernel(const __grid_constant__ CimShape s, const __grid_constant__ StepArgs a) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const Grp<32> g(threadIdx.x & 31);
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int rep = warp_global; rep < s.n_replicas; rep += n_warps) {
        if (a.active && !a.active[rep]) continue;
        Replica r = make_replica(s, a, rep, a.state + (int64_t)rep * s.SW);  // operate directly on global memory
        replica_reset<32>(s, g, r);
        if (g.lane == 0) a.light[rep] = 0;
    }
}

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// hello-world random agent (examples/hello_world/cim/hello.py:24-32) as a counter hash of (replica, step)
__global__ void cim_policy_kernel(const int32_t* __restrict__ dec, int32_t* __restrict__ act, int n, int max_actions,
                                  uint32_t seed, uint32_t replica_base) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* d = dec + i * 8;
    const uint32_t step = (uint32_t)d[7];
    uint32_t h1 = hash_u32(seed ^ hash_u32((uint32_t)(i + replica_base) * 0x9e3779b9u + step * 0x85ebca6bu + 0x1234567u));
    uint32_t h2 = hash_u32(h1 + 0x68bc21ebu);
    int load = d[3], dis = d[4];
    bool to_discharge = dis > 0 && (h1 & 1u);
    int scope = to_discharge ? dis : load;
    int qty = scope > 0 ? (int)(h2 % (uint32_t)(scope + 1)) : 0;
    int4 o = make_int4(d[2], d[1], qty, to_discharge ? 1 : 0);
    *reinterpret_cast<int4*>(act + (int64_t)i * max_actions * 4) = o;
}

struct MaroCimEnv : EnvCommon {
    CimShape s;
    int K = 0, mt_words = 0, warps_per_cta = 4, lanes = 32, grid = 0, max_stops = 0, max_targets = 0, max_distinct = 0;
    bool spread = false;  // one replica per warp (cim_step_kernel kSpread)
    size_t smem_bytes = 0;
    int32_t *d_tables = nullptr, *d_topo = nullptr;
    uint32_t* d_mt = nullptr;
    uint8_t* d_light = nullptr;
    std::vector<int32_t> h_tables;
};

// =====================================================================================================
// RL state / reward shaping on the snapshot ring (examples/cim/rl/env_sampler.py:15-36, 66-80)
// =====================================================================================================
struct ShapeArgs {
    const int32_t* snap;
    const int32_t* snap_frame;
    int ring_rows, FWp, B;
    // state
    const int32_t* decisions;  // [B][8]
    int look_back_ticks;       // look_back - 1 frames: max(0, tick - rt), rt = 0..look_back-2
    int n_ports_per_state;     // 1 + future_stop_number
    int npa, nva;              // attribute counts
    int port_attr_off[16], port_attr_isf[16], vessel_attr_off[16], vessel_attr_isf[16];
    int o_fut, fut;            // future_stop_list: word offset, slots per vessel
    int P, V;
    double* state_out;         // [B][look_back_ticks * n_ports_per_state * npa + nva]
    // reward
    const int32_t* ticks;      // [B] tick of the action
    const int32_t* ports;      // [B] port that acted
    const double* decay;       // [time_window] time_decay ** i
    int time_window, off_fulfillment, off_shortage;
    double fulfillment_factor, shortage_factor;
    float* reward_out;         // [B]
    // action translation
    const int32_t* model_actions;  // [B] index into action_space
    const double* action_space;    // [n_action_space]
    int n_action_space, finite_vessel_space, has_early_discharge, max_actions, off_remaining_space, off_early_discharge;
    int32_t* actions_out;          // [B][max_actions][4]
};

// word `w` of snapshot `frame` of replica `rep`; frames not in the ring read as 0 (np_backend.pyx:543-549)
__device__ __forceinline__ bool snap_row(const ShapeArgs& q, int rep, int frame, const int32_t*& row) {
    if (frame < 0) return false;
    int r = frame % q.ring_rows;
    if (q.snap_frame[(int64_t)rep * q.ring_rows + r] != frame) return false;
    row = q.snap + ((int64_t)rep * q.ring_rows + r) * q.FWp;
    return true;
}

// state[rep] = concat(ports[ticks : [port] + future_stop_list : port_attrs], vessels[tick : vessel : vessel_attrs]) as float64
__global__ void cim_rl_state_kernel(const __grid_constant__ ShapeArgs q) {
    const int per_tick = q.n_ports_per_state * q.npa;
    const int dim = q.look_back_ticks * per_tick + q.nva;
    const int64_t total = (int64_t)q.B * dim;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int rep = (int)(i / dim), e = (int)(i % dim);
        const int32_t* d = q.decisions + (int64_t)rep * 8;
        double v = 0.0;
        if (d[MARO_DEC_STATUS] == MARO_STATUS_DECISION) {
            const int tick = d[MARO_DEC_TICK], vessel = d[MARO_DEC_VESSEL];
            const int32_t* now = nullptr;
            const bool have_now = snap_row(q, rep, tick, now);
            if (e >= q.look_back_ticks * per_tick) {
                const int a = e - q.look_back_ticks * per_tick;
                if (have_now) {
                    int w = now[q.vessel_attr_off[a] + vessel];
                    v = q.vessel_attr_isf[a] ? (double)__int_as_float(w) : (double)w;
                }
            } else {
                const int k = e / per_tick, j = (e % per_tick) / q.npa, a = e % q.npa;
                int port = d[MARO_DEC_PORT];
                if (j > 0) port = have_now ? now[q.o_fut + vessel * q.fut + (j - 1)] : 0;  // .astype("int") of a 0-padded query
                const int frame = tick - k > 0 ? tick - k : 0;
                const int32_t* row = nullptr;
                if (port >= 0 && port < q.P && snap_row(q, rep, frame, row)) {
                    int w = row[q.port_attr_off[a] + port];
                    v = q.port_attr_isf[a] ? (double)__int_as_float(w) : (double)w;
                }
            }
        }
        q.state_out[i] = v;
    }
}

// reward[rep] = float32(ff * sum_k decay[k] * fulfillment[tick+1+k, port] - sf * sum_k decay[k] * shortage[tick+1+k, port])
__global__ void cim_rl_reward_kernel(const __grid_constant__ ShapeArgs q) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int rep = warp_global; rep < q.B; rep += n_warps) {
        const int tick = q.ticks[rep], port = q.ports[rep];
        double f = 0.0, sh = 0.0;
        if (port >= 0 && port < q.P) {
            for (int k = lane; k < q.time_window; k += 32) {
                const int32_t* row = nullptr;
                if (snap_row(q, rep, tick + 1 + k, row)) {
                    f += q.decay[k] * (double)row[q.off_fulfillment + port];
                    sh += q.decay[k] * (double)row[q.off_shortage + port];
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            f += __shfl_xor_sync(0xffffffffu, f, o);
            sh += __shfl_xor_sync(0xffffffffu, sh, o);
        }
        if (lane == 0) q.reward_out[rep] = (float)(q.fulfillment_factor * f - q.shortage_factor * sh);
    }
}

static void register_attrs(MaroCimEnv* e) {
    const CimShape& s = e->s;
    static const char* pn[] = {"acc_booking", "acc_fulfillment", "acc_shortage", "booking", "capacity", "empty",
                               "fulfillment", "full", "on_consignee", "on_shipper", "shortage", "transfer_cost"};
    for (int a = 0; a < 12; a++) e->attrs[0].push_back({pn[a], a * s.P, 1, a == 11, s.P});
    static const char* vn[] = {"capacity", "early_discharge", "empty", "full", "is_parking", "last_loc_idx",
                               "loc_port_idx", "next_loc_idx", "remaining_space", "route_idx"};
    for (int a = 0; a < 10; a++) e->attrs[1].push_back({vn[a], s.o_vs + a * s.V, 1, 0, s.V});
    e->attrs[1].push_back({"past_stop_list", s.o_past, s.past, 0, s.V});
    e->attrs[1].push_back({"past_stop_tick_list", s.o_past_tick, s.past, 0, s.V});
    e->attrs[1].push_back({"future_stop_list", s.o_fut, s.fut, 0, s.V});
    e->attrs[1].push_back({"future_stop_tick_list", s.o_fut_tick, s.fut, 0, s.V});
    e->attrs[2].push_back({"full_on_ports", s.o_fop, s.P * s.P, 0, 1});
    e->attrs[2].push_back({"full_on_vessels", s.o_fov, s.V * s.P, 0, 1});
    e->attrs[2].push_back({"vessel_plans", s.o_vp, s.V * s.P, 0, 1});
}

static StepArgs base_args(MaroCimEnv* e) {
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.state = e->d_state;
    a.snap = e->d_snap;
    a.snap_frame = e->d_snap_frame;
    a.mt = e->d_mt;
    a.tables = e->d_tables;
    a.replica_topology = e->d_topo;
    a.light = e->d_light;
    a.mt_words = e->mt_words;
    return a;
}

template <int W, int G, bool kGeneral>
static cudaError_t launch_step_wgn(MaroCimEnv* e, const StepArgs& a) {
    if (G < 32 && W == 4 && e->spread) {  // one replica per warp (small batches), instantiated for 4 warps per CTA only
        cudaError_t err = cudaFuncSetAttribute(cim_step_kernel<W, G, kGeneral, (G < 32 && W == 4)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes);
        if (err != cudaSuccess) return err;
        cim_step_kernel<W, G, kGeneral, (G < 32 && W == 4)><<<e->grid, W * 32, e->smem_bytes, e->stream>>>(e->s, a);
        return cudaGetLastError();
    }
    cudaError_t err = cudaFuncSetAttribute(cim_step_kernel<W, G, kGeneral>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes);
    if (err != cudaSuccess) return err;
    cim_step_kernel<W, G, kGeneral><<<e->grid, W * 32, e->smem_bytes, e->stream>>>(e->s, a);
    return cudaGetLastError();
}

template <int W, int G>
static cudaError_t launch_step_wg(MaroCimEnv* e, const StepArgs& a) {
    // noise-free fixed-mode topologies run the specialised kernel (no MT19937 / float64 paths compiled in)
    const bool general = !(e->s.order_table && !e->s.buffer_noise);
    return general ? launch_step_wgn<W, G, true>(e, a) : launch_step_wgn<W, G, false>(e, a);
}

template <int G>
static cudaError_t launch_step_g(MaroCimEnv* e, const StepArgs& a) {
    switch (e->warps_per_cta) {
        case 1: return launch_step_wg<1, G>(e, a);
        case 2: return launch_step_wg<2, G>(e, a);
        case 4: return launch_step_wg<4, G>(e, a);
        default: return launch_step_wg<8, G>(e, a);
    }
}

static cudaError_t launch_step(MaroCimEnv* e, const StepArgs& a) {
    switch (e->lanes) {
        case 8: return launch_step_g<8>(e, a);
        case 16: return launch_step_g<16>(e, a);
        default: return launch_step_g<32>(e, a);
    }
}

extern "C" {

const char* maro_last_error(void) { return g_err.c_str(); }
int maro_abi_version(void) { return MARO_B200_ABI_VERSION; }

int maro_cim_destroy(MaroCimEnv* e) {
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaFree(e->d_tables); cudaFree(e->d_topo); cudaFree(e->d_mt); cudaFree(e->d_light);
    common_free(e);
    delete e;
    return 0;
}

int maro_cim_create(const MaroCimTopology* topos, int32_t n_topos, const MaroCimConfig* cfg, MaroCimEnv** out) {
    if (!topos || n_topos < 1 || !cfg || !out || cfg->n_replicas < 1) return fail("maro_cim_create: bad arguments");
    const MaroCimTopology& t0 = topos[0];
    if (t0.n_ports < 1 || t0.n_ports > 255 || t0.n_vessels < 1 || t0.n_vessels > 64)
        return fail("maro_cim_create: supported sizes are 1..255 ports and 1..64 vessels");
    for (int k = 1; k < n_topos; k++)
        if (check_same_shape(t0, topos[k])) return fail("maro_cim_create: all topologies of one handle must share a shape");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("maro_cim_create: no CUDA device — this library has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("maro_cim_create: bad device ordinal");
    CK(cudaSetDevice(cfg->device));

    MaroCimEnv* e = new MaroCimEnv();
    e->device = cfg->device;
    e->B = cfg->n_replicas;
    e->K = n_topos;
    CimShape& s = e->s;
    if (compute_shape_and_tables(topos, n_topos, cfg, s, e->h_tables, e->max_stops, e->max_targets, e->max_distinct)) {
        delete e;
        return fail("maro_cim_create: inconsistent topology tables / durations must be positive");
    }
    const char* ln = getenv("MARO_B200_LANES");  // tuning override: lanes per replica (8 / 16 / 32, >= the topology's minimum)
    const int cfg_lanes = ln ? std::max(atoi(ln), lanes_per_replica(s)) : 0;
    register_attrs(e);

    // launch geometry: G lanes per replica, as many warps per CTA as shared memory allows (<= 8), persistent grid
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, e->device));
    e->lanes = cfg_lanes > 0 ? cfg_lanes : lanes_per_replica(s);
    const int gpw = 32 / e->lanes;  // replicas per warp
    const size_t per_warp = (size_t)s.SW * 4 * gpw;
    const size_t max_smem = prop.sharedMemPerBlockOptin;
    if (256 + per_warp > max_smem) { delete e; return fail("maro_cim_create: replica state does not fit in shared memory"); }
    // warps per CTA: the choice that keeps the most warps resident per SM (shared memory vs the 64-register budget
    // of the specialised kernel), then fewer for small batches so that replicas spread over all SMs
    const size_t sm_smem = prop.sharedMemPerMultiprocessor;
    int w = 1, best = 0;
    for (int cand = 8; cand >= 1; cand >>= 1) {
        size_t cta = 256 + per_warp * cand;
        if (cta > max_smem) continue;
        int blocks = (int)std::min<size_t>(sm_smem / (cta + 1024), (size_t)(64 / cand));
        if (blocks * cand > best) { best = blocks * cand; w = cand; }
    }
    while (w > 1 && (e->B + w * gpw - 1) / (w * gpw) < prop.multiProcessorCount) w >>= 1;
    e->warps_per_cta = w;
    e->smem_bytes = 256 + per_warp * w;
    int ctas_needed = (e->B + w * gpw - 1) / (w * gpw);
    int resident = std::max<int>(1, (int)std::min<size_t>(64 / w, sm_smem / (e->smem_bytes + 1024)));
    e->grid = std::min(ctas_needed, prop.multiProcessorCount * resident);
    // small batch, sub-warp groups: one replica per warp while the replicas fit the resident warp slots (<= 32 per SM)
    const char* sp = getenv("MARO_B200_SPREAD");
    const bool want_spread = sp ? atoi(sp) != 0 : e->B <= prop.multiProcessorCount * 32;  // measured crossover 4 k .. 8 k replicas
    if (gpw > 1 && want_spread && 256 + (size_t)s.SW * 4 * 4 <= max_smem) {
        e->spread = true;
        e->warps_per_cta = 4;
        e->smem_bytes = 256 + (size_t)s.SW * 4 * 4;
        e->grid = (e->B + 3) / 4;
    }

    e->ring_rows = s.ring_rows; e->FW = s.FW; e->FWp = s.FWp; e->SW = s.SW;
    e->off_tick = s.FWp + C_TICK; e->off_counters = s.FWp + C_NSTEPS_LO;
    e->dec_words = MARO_CIM_DECISION_WORDS; e->max_actions = s.max_actions;
    if (common_alloc(e)) { maro_cim_destroy(e); return 1; }
    const int B = e->B;
    CK(cudaMalloc(&e->d_tables, e->h_tables.size() * 4));
    CK(cudaMalloc(&e->d_topo, (size_t)B * 4));
    CK(cudaMalloc(&e->d_light, (size_t)B));
    CK(cudaMemset(e->d_light, 0, (size_t)B));
    CK(cudaMemcpy(e->d_tables, e->h_tables.data(), e->h_tables.size() * 4, cudaMemcpyHostToDevice));
    std::vector<int32_t> topo(B, 0);
    if (cfg->replica_topology)
        for (int i = 0; i < B; i++) {
            if (cfg->replica_topology[i] < 0 || cfg->replica_topology[i] >= n_topos) { maro_cim_destroy(e); return fail("maro_cim_create: replica_topology out of range"); }
            topo[i] = cfg->replica_topology[i];
        }
    CK(cudaMemcpy(e->d_topo, topo.data(), (size_t)B * 4, cudaMemcpyHostToDevice));
    if (s.order_noise || s.buffer_noise) {
        e->mt_words = mt_block_words(s);
        CK(cudaMalloc(&e->d_mt, (size_t)B * e->mt_words * 4));
    }
    *out = e;
    int rc = maro_cim_reset(e, nullptr);
    if (rc) { maro_cim_destroy(e); *out = nullptr; return rc; }
    return 0;
}

int maro_cim_set_stream(MaroCimEnv* e, void* cuda_stream, int32_t external) {
    if (!e) return fail("null handle");
    e->stream = external ? (cudaStream_t)cuda_stream : e->own_stream;
    return 0;
}

int maro_cim_reset(MaroCimEnv* e, const uint8_t* mask) {
    if (!e) return fail("null handle");
    CK(cudaSetDevice(e->device));
    StepArgs a = base_args(e);
    if (mask) {
        uint8_t* d_active = e->d_in + (size_t)e->B * e->s.max_actions * 16 + (size_t)e->B * 4;
        memcpy(e->h_in, mask, e->B);
        CK(cudaMemcpyAsync(d_active, e->h_in, e->B, cudaMemcpyHostToDevice, e->stream));
        a.active = d_active;
    }
    int threads = 128, blocks = std::min((e->B * 32 + threads - 1) / threads, 148 * 16);
    cim_reset_kernel<<<blocks, threads, 0, e->stream>>>(e->s, a);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

int maro_cim_set_topology(MaroCimEnv* e, int32_t index, const MaroCimTopology* topo) {
    if (!e || !topo || index < 0 || index >= e->K) return fail("maro_cim_set_topology: bad arguments");
    CK(cudaSetDevice(e->device));
    std::vector<int32_t> blob;
    if (topo->n_ports != e->s.P || topo->n_vessels != e->s.V || topo->max_tick != e->s.max_tick)
        return fail("maro_cim_set_topology: shape differs from the handle's");
    if (topo->stop_offset[topo->n_vessels] > e->max_stops || topo->target_offset[topo->n_ports] > e->max_targets)
        return fail("maro_cim_set_topology: more stops/targets than the handle was sized for");
    CimShape probe = e->s;  // rebuild with identical padding; offsets must come out the same
    if ((e->s.order_table && count_distinct_orders(*topo) > e->max_distinct) ||
        build_blob(*topo, probe, blob, e->max_stops, e->max_targets, true, e->max_distinct) || probe.table_words != e->s.table_words ||
        probe.t_mt_buffer != e->s.t_mt_buffer || probe.t_order_proportion != e->s.t_order_proportion)
        return fail("maro_cim_set_topology: shape differs from the handle's");
    memcpy(e->h_tables.data() + (size_t)index * e->s.table_words, blob.data(), blob.size() * 4);
    CK(cudaMemcpy(e->d_tables + (size_t)index * e->s.table_words, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

int maro_cim_step_device(MaroCimEnv* e, const uint8_t* d_active, const int32_t* d_actions, const int32_t* d_n_actions,
                         int32_t* d_decisions, int64_t* d_metrics) {
    if (!e || !d_decisions || !d_metrics) return fail("maro_cim_step_device: bad arguments");
    CK(cudaSetDevice(e->device));
    StepArgs a = base_args(e);
    a.active = d_active; a.actions = d_actions; a.n_actions = d_n_actions;
    a.decisions = d_decisions; a.metrics = d_metrics;
    CK(launch_step(e, a));
    return 0;
}

int maro_cim_step(MaroCimEnv* e, const uint8_t* active, const int32_t* actions, const int32_t* n_actions,
                  int32_t* decisions, int64_t* metrics) {
    if (!e || !decisions || !metrics) return fail("maro_cim_step: bad arguments");
    CK(cudaSetDevice(e->device));
    return common_host_step(e, active, actions, n_actions, decisions, metrics,
                            [&](const uint8_t* a, const int32_t* ac, const int32_t* na, int32_t* d, int64_t* m) {
                                return maro_cim_step_device(e, a, ac, na, d, m);
                            });
}

int maro_cim_pinned_buffers(MaroCimEnv* e, void** actions, void** n_actions, void** active, void** decisions, void** metrics) {
    return common_pinned_buffers(e, actions, n_actions, active, decisions, metrics);
}
int maro_cim_step_pinned(MaroCimEnv* e, int32_t use_actions, int32_t use_n_actions, int32_t use_active) {
    if (!e) return fail("maro_cim_step_pinned: null handle");
    CK(cudaSetDevice(e->device));
    const uint8_t* f = reinterpret_cast<const uint8_t*>(1);  // presence flags only
    return common_host_step(e, use_active ? f : nullptr, use_actions ? reinterpret_cast<const int32_t*>(f) : nullptr,
                            use_n_actions ? reinterpret_cast<const int32_t*>(f) : nullptr, nullptr, nullptr,
                            [&](const uint8_t* a, const int32_t* ac, const int32_t* na, int32_t* d, int64_t* m) {
                                return maro_cim_step_device(e, a, ac, na, d, m);
                            }, true);
}
int32_t maro_cim_frame_words(MaroCimEnv* e) { return e ? e->s.FW : -1; }

int maro_cim_query(MaroCimEnv* e, const int32_t* replicas, int32_t n_replicas, int32_t node_type, const int32_t* frame_indices,
                   int32_t n_frames, const int32_t* nodes, int32_t n_nodes, const int32_t* attrs, int32_t n_attrs, double* out,
                   int64_t* out_per_replica) {
    if (!out) return fail("maro_cim_query: null output");
    return query_impl(e, replicas, n_replicas, node_type, frame_indices, n_frames, nodes, n_nodes, attrs, n_attrs, nullptr, out, out_per_replica);
}

int maro_cim_query_device(MaroCimEnv* e, const int32_t* replicas, int32_t n_replicas, int32_t node_type, const int32_t* frame_indices,
                          int32_t n_frames, const int32_t* nodes, int32_t n_nodes, const int32_t* attrs, int32_t n_attrs, double* d_out,
                          int64_t* out_per_replica) {
    if (!d_out) return fail("maro_cim_query_device: null output");
    return query_impl(e, replicas, n_replicas, node_type, frame_indices, n_frames, nodes, n_nodes, attrs, n_attrs, d_out, nullptr, out_per_replica);
}

int32_t maro_cim_attr_id(MaroCimEnv* e, int32_t node_type, const char* name) { return common_attr_id(e, node_type, name); }
int32_t maro_cim_attr_slots(MaroCimEnv* e, int32_t node_type, int32_t attr_id) { return common_attr_slots(e, node_type, attr_id); }
int maro_cim_read_frame(MaroCimEnv* e, int32_t replica, int32_t* out_words, int32_t n_words) { return common_read_frame(e, replica, out_words, n_words); }
int maro_cim_ticks(MaroCimEnv* e, int32_t* out_ticks) { return common_ticks(e, out_ticks); }
int maro_cim_counters(MaroCimEnv* e, int64_t* out) { return common_counters(e, out); }
int maro_cim_snapshot_frames(MaroCimEnv* e, int32_t replica, int32_t* out, int32_t cap, int32_t* n_out) {
    return common_snapshot_frames(e, replica, out, cap, n_out);
}

int maro_cim_random_policy_device(MaroCimEnv* e, const int32_t* d_decisions, int32_t* d_actions, uint32_t seed,
                                  uint32_t replica_base) {
    if (!e || !d_decisions || !d_actions) return fail("maro_cim_random_policy_device: bad arguments");
    CK(cudaSetDevice(e->device));
    int threads = 256, blocks = (e->B + threads - 1) / threads;
    cim_policy_kernel<<<blocks, threads, 0, e->stream>>>(d_decisions, d_actions, e->B, e->s.max_actions, seed, replica_base);
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"

// _translate_to_env_action (examples/cim/rl/env_sampler.py:38-64): model action index -> {vessel, port, quantity, type}
__global__ void cim_rl_action_kernel(const __grid_constant__ ShapeArgs q) {
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= q.B) return;
    const int32_t* d = q.decisions + (int64_t)rep * 8;
    int4 out = make_int4(0, 0, 0, 0);
    if (d[MARO_DEC_STATUS] == MARO_STATUS_DECISION) {
        const int tick = d[MARO_DEC_TICK], vessel = d[MARO_DEC_VESSEL];
        int m = q.model_actions[rep];
        m = m < 0 ? 0 : (m >= q.n_action_space ? q.n_action_space - 1 : m);
        const int32_t* now = nullptr;
        const bool have_now = snap_row(q, rep, tick, now);
        const double percent = fabs(q.action_space[m]);
        const double zero_action_idx = (double)q.n_action_space / 2.0;
        double quantity;
        int type;
        if ((double)m < zero_action_idx) {
            type = 0;  // ActionType.LOAD
            quantity = rint(percent * (double)d[MARO_DEC_SCOPE_LOAD]);  // python round(): half to even
            if (q.finite_vessel_space) {
                const double space = have_now ? (double)now[q.off_remaining_space + vessel] : 0.0;
                quantity = quantity <= space ? quantity : space;
            }
        } else {
            type = 1;  // ActionType.DISCHARGE ((double)m == zero_action_idx cannot happen for an odd-sized space either way)
            const double early = q.has_early_discharge && have_now ? (double)now[q.off_early_discharge + vessel] : 0.0;
            const double plan = percent * ((double)d[MARO_DEC_SCOPE_DISCHARGE] + early) - early;
            quantity = plan > 0 ? rint(plan) : rint(percent * (double)d[MARO_DEC_SCOPE_DISCHARGE]);
        }
        out = make_int4(vessel, d[MARO_DEC_PORT], (int)quantity, type);
    }
    *reinterpret_cast<int4*>(q.actions_out + (int64_t)rep * q.max_actions * 4) = out;
}

static int shape_common(MaroCimEnv* e, ShapeArgs& q) {
    memset(&q, 0, sizeof(q));
    q.snap = e->d_snap; q.snap_frame = e->d_snap_frame; q.ring_rows = e->ring_rows; q.FWp = e->FWp; q.B = e->B;
    q.P = e->s.P; q.V = e->s.V; q.o_fut = e->s.o_fut; q.fut = e->s.fut;
    return 0;
}

extern "C" {

int32_t maro_cim_rl_state_dim(MaroCimEnv* e, int32_t look_back, int32_t n_port_attrs, int32_t n_vessel_attrs) {
    if (!e || look_back < 2) return -1;
    return (look_back - 1) * (1 + e->s.fut) * n_port_attrs + n_vessel_attrs;
}

int maro_cim_rl_state_device(MaroCimEnv* e, const int32_t* d_decisions, int32_t look_back, const int32_t* port_attrs,
                             int32_t n_port_attrs, const int32_t* vessel_attrs, int32_t n_vessel_attrs, double* d_out) {
    if (!e || !d_decisions || !d_out || !port_attrs || !vessel_attrs || look_back < 2 || n_port_attrs < 1 || n_port_attrs > 16 ||
        n_vessel_attrs < 0 || n_vessel_attrs > 16)
        return fail("maro_cim_rl_state_device: bad arguments");
    CK(cudaSetDevice(e->device));
    ShapeArgs q;
    shape_common(e, q);
    for (int i = 0; i < n_port_attrs; i++) {
        int a = port_attrs[i];
        if (a < 0 || a >= (int)e->attrs[0].size() || e->attrs[0][a].slots != 1) return fail("maro_cim_rl_state_device: bad port attribute");
        q.port_attr_off[i] = e->attrs[0][a].off; q.port_attr_isf[i] = e->attrs[0][a].isf;
    }
    for (int i = 0; i < n_vessel_attrs; i++) {
        int a = vessel_attrs[i];
        if (a < 0 || a >= (int)e->attrs[1].size() || e->attrs[1][a].slots != 1) return fail("maro_cim_rl_state_device: bad vessel attribute");
        q.vessel_attr_off[i] = e->attrs[1][a].off; q.vessel_attr_isf[i] = e->attrs[1][a].isf;
    }
    q.decisions = d_decisions; q.look_back_ticks = look_back - 1; q.n_ports_per_state = 1 + e->s.fut;
    q.npa = n_port_attrs; q.nva = n_vessel_attrs; q.state_out = d_out;
    const int64_t total = (int64_t)e->B * maro_cim_rl_state_dim(e, look_back, n_port_attrs, n_vessel_attrs);
    int threads = 256, blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 8);
    cim_rl_state_kernel<<<blocks, threads, 0, e->stream>>>(q);
    CK(cudaGetLastError());
    return 0;
}

int maro_cim_rl_action_device(MaroCimEnv* e, const int32_t* d_decisions, const int32_t* d_model_actions, const double* d_action_space,
                              int32_t n_action_space, int32_t finite_vessel_space, int32_t has_early_discharge, int32_t* d_actions) {
    if (!e || !d_decisions || !d_model_actions || !d_action_space || !d_actions || n_action_space < 1)
        return fail("maro_cim_rl_action_device: bad arguments");
    CK(cudaSetDevice(e->device));
    ShapeArgs q;
    shape_common(e, q);
    q.decisions = d_decisions; q.model_actions = d_model_actions; q.action_space = d_action_space; q.n_action_space = n_action_space;
    q.finite_vessel_space = finite_vessel_space; q.has_early_discharge = has_early_discharge; q.max_actions = e->s.max_actions;
    q.off_remaining_space = e->attrs[1][common_attr_id(e, 1, "remaining_space")].off;
    q.off_early_discharge = e->attrs[1][common_attr_id(e, 1, "early_discharge")].off;
    q.actions_out = d_actions;
    int threads = 256, blocks = (e->B + threads - 1) / threads;
    cim_rl_action_kernel<<<blocks, threads, 0, e->stream>>>(q);
    CK(cudaGetLastError());
    return 0;
}

int maro_cim_rl_reward_device(MaroCimEnv* e, const int32_t* d_ticks, const int32_t* d_ports, const double* d_decay,
                              int32_t time_window, double fulfillment_factor, double shortage_factor, float* d_out) {
    if (!e || !d_ticks || !d_ports || !d_decay || !d_out || time_window < 1) return fail("maro_cim_rl_reward_device: bad arguments");
    CK(cudaSetDevice(e->device));
    ShapeArgs q;
    shape_common(e, q);
    q.ticks = d_ticks; q.ports = d_ports; q.decay = d_decay; q.time_window = time_window;
    q.off_fulfillment = e->attrs[0][common_attr_id(e, 0, "fulfillment")].off;
    q.off_shortage = e->attrs[0][common_attr_id(e, 0, "shortage")].off;
    q.fulfillment_factor = fulfillment_factor; q.shortage_factor = shortage_factor; q.reward_out = d_out;
    int threads = 128, blocks = std::min((e->B * 32 + threads - 1) / threads, 148 * 16);
    cim_rl_reward_kernel<<<blocks, threads, 0, e->stream>>>(q);
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"

