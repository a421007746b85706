// bike_host.hpp — host-side (plain C++) shape + static-table serialisation for the citi_bike scenario, shared by
// bike_env.cu and the test-only host-emulation harness.  No CUDA here.
#pragma once
#include "bike_core.cuh"
#include "cim_host.hpp"

namespace maro {

// numpy legacy seeding for an integer seed: mt19937_seed (numpy/random/src/mt19937/mt19937.c)
inline void np_mt_seed(uint32_t seed, uint32_t* mt) {
    for (int i = 0; i < 624; i++) {
        mt[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
}

inline int bike_compute_shape_and_tables(const MaroBikeTopology& t, const MaroCimConfig* cfg, BikeShape& s,
                                         std::vector<int32_t>& tables) {
    memset(&s, 0, sizeof(s));
    const int S = t.n_stations;
    if (S < 1 || S > 255) return 1;
    s.S = S; s.max_tick = t.max_tick; s.start_tick = cfg->start_tick;
    s.snap_res = cfg->snapshot_resolution > 0 ? cfg->snapshot_resolution : 1;
    s.res_is_one = s.snap_res == 1;
    const int durations = s.max_tick - s.start_tick;
    if (durations < 1) return 1;
    const int total_frames = (durations + s.snap_res - 1) / s.snap_res;
    s.ring_rows = cfg->max_snapshots > 0 ? std::min(cfg->max_snapshots, total_frames) : total_frames;
    s.resolution = t.resolution; s.extra_cost_mode = t.extra_cost_mode;
    s.max_actions = cfg->max_actions > 0 ? cfg->max_actions : 1;
    s.n_replicas = cfg->n_replicas;
    s.time_mean = t.time_mean; s.time_std = t.time_std; s.supply_ratio = t.supply_ratio; s.demand_ratio = t.demand_ratio;
    s.scope_low = t.scope_low; s.scope_high = t.scope_high;
    s.FW = BA_COUNT * S + S * S;
    s.FWp = round_up(s.FW, 4);
    s.CWp = round_up(BC_COUNT, 4);
    const int ntrips = t.trip_offset[t.max_tick];
    int max_dur = 1, max_per_tick = 1;
    for (int k = 0; k < ntrips; k++) max_dur = std::max(max_dur, t.trip_dur[k]);
    for (int k = 0; k < t.max_tick; k++) max_per_tick = std::max(max_per_tick, t.trip_offset[k + 1] - t.trip_offset[k]);
    int max_delay = std::max(max_dur, (int)ceil(t.time_mean + 8.0 * fabs(t.time_std)) + 1) + 1;
    int qh = 16;
    while (qh < max_delay + 1) qh <<= 1;
    s.QH = qh;
    // outstanding events: every trip of the last max_dur ticks may still be out + deliveries + a tick's decisions
    int qn = cfg->queue_capacity > 0 ? cfg->queue_capacity
                                     : std::min(4096, std::max(32, max_per_tick * std::min(max_dur + 1, 64) + 3 * S + 8));
    s.QN = round_up(qn, 4);
    s.o_scratch = round_up(s.FWp + s.CWp + s.QN * 2 + s.QH + s.QN, 4);
    s.SW = s.o_scratch + round_up(4 * S, 4);  // (+ the action-scope scratch: it used to sit in the per-replica global block)
    s.DW = MARO_BIKE_DEC_HEAD + 2 * S;
    // action-scope filter chain
    if (t.n_filters < 0 || t.n_filters > MARO_BIKE_MAX_FILTERS) return 1;
    s.n_filters = t.n_filters;
    s.tw_windows = 1;
    bool dropped = false;  // has an earlier filter been able to drop a neighbour?
    for (int f = 0; f < t.n_filters; f++) {
        s.filter_type[f] = t.filter_type[f]; s.filter_num[f] = t.filter_num[f]; s.filter_windows[f] = t.filter_windows[f];
        if (t.filter_type[f] < 0 || t.filter_type[f] > 2 || t.filter_num[f] < 0) return 1;
        if (t.filter_type[f] == MARO_BIKE_FILTER_TRIP_WINDOW) {
            if (t.filter_windows[f] < 1) return 1;
            s.tw_windows = std::max(s.tw_windows, t.filter_windows[f]);
        }
        // a distance filter behind a dropping filter indexes the dropped neighbour: KeyError in the reference
        // (decision_strategy.py:45-48) — such a configuration cannot run there either
        if (t.filter_type[f] == MARO_BIKE_FILTER_DISTANCE && dropped) return 2;
        if (t.filter_num[f] < S - 1) { dropped = true; s.drops = 1; }
    }
    s.scope_off = 632;
    s.rng_words = round_up(632 + s.tw_windows + s.tw_windows * S + 4 * S, 4);
    tables.clear();
    BlobBuilder b(tables);
    s.t_bikes = b.put_i(t.station_bikes, S, S);
    s.t_capacity = b.put_i(t.station_capacity, S, S);
    s.t_id = b.put_i(t.station_id, S, S);
    s.t_nbr_offset = b.put_i(t.nbr_offset, S + 1, S + 1);
    s.t_nbr_idx = b.put_i(t.nbr_idx, t.nbr_offset[S], std::max(1, t.nbr_offset[S]));
    s.t_trip_offset = b.put_i(t.trip_offset, t.max_tick + 1, t.max_tick + 1);
    s.t_trip_src = b.put_i(t.trip_src, ntrips, std::max(1, ntrips));
    s.t_trip_dst = b.put_i(t.trip_dst, ntrips, std::max(1, ntrips));
    s.t_trip_dur = b.put_i(t.trip_dur, ntrips, std::max(1, ntrips));
    s.t_day_of_tick = b.put_i(t.day_of_tick, t.max_tick, t.max_tick);
    s.t_day_feat = b.put_i(t.day_feat, 4 * t.n_days, 4 * t.n_days);
    uint32_t mt[624];
    np_mt_seed(t.transfer_seed, mt);
    s.t_mt = b.put_i(reinterpret_cast<int32_t*>(mt), 624, 624);
    tables.resize(round_up((int)tables.size(), 4), 0);
    return 0;
}

inline int bike_lanes_per_replica(const BikeShape& s) { return s.S <= 8 ? 8 : (s.S <= 16 ? 16 : 32); }

}  // namespace maro
