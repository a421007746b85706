// bike_core.cuh — per-replica citi_bike simulation core (device code; one lane group = one replica).
//
// From-scratch formulation of the reference's Env.step for the citi_bike scenario
// (maro/simulator/core.py:317-381, maro/simulator/scenarios/citi_bike/business_engine.py:101-147, 398-559,
//  decision_strategy.py:166-397, station.py:70-75).
//
// Unlike CIM, the citi_bike handlers are order dependent across stations (RequireBike tests `bikes < 1`, a full
// station overflows into its distance-sorted neighbours), so the event chain of a tick runs on the leader lane out of
// shared memory; the lanes cooperate on the wide parts (stage-in/out, snapshots, per-station sweeps).
// Tick T's execution order (reference: insertion order of `_pending_events[T]`):
//   1. events queued by earlier ticks (ReturnBike / DeliverBike) — calendar-queue bucket T, FIFO, they push nothing;
//   2. this tick's trips (RequireBike) in trace order — a trip with duration 0 appends its ReturnBike to bucket T;
//   3. RebalanceBike on decision ticks — appends one decision event per station over/under the water marks;
//   4. bucket T again (FIFO): late ReturnBikes, the decision events (each stops the kernel: Sequential mode), and
//      DeliverBikes of actions whose transfer time rounds to 0.
#pragma once
#include "cim_core.cuh"  // lane-group primitives, ctrl_get64/ctrl_add64 style helpers are re-declared below

namespace maro {

enum BikeAttr { BA_BIKES, BA_CAPACITY, BA_EXTRA_COST, BA_FAILED_RETURN, BA_FULFILLMENT, BA_HOLIDAY, BA_ID, BA_MIN_BIKES,
                BA_SHORTAGE, BA_TEMPERATURE, BA_TRANSFER_COST, BA_TRIP_REQUIREMENT, BA_WEATHER, BA_WEEKDAY, BA_COUNT };

enum BikeCtrl {
    BC_STATE, BC_TICK, BC_PHASE, BC_PEND_STATION, BC_PEND_TYPE, BC_FREE_TOP, BC_Q_COUNT, BC_ERR,
    BC_TRIPS_LO, BC_TRIPS_HI, BC_SHORT_LO, BC_SHORT_HI, BC_OPNUM_LO, BC_OPNUM_HI, BC_LAST_DAY, BC_EP_STEP,
    BC_NSTEPS_LO, BC_NSTEPS_HI, BC_NTICKS_LO, BC_NTICKS_HI, BC_NEVENTS_LO, BC_NEVENTS_HI, BC_NSNAPS_LO, BC_NSNAPS_HI,
    BC_COUNT
};
enum BikeEv { BE_RETURN = 0, BE_DELIVER = 1, BE_DECISION = 2 };

struct BikeShape {
    int S, max_tick, start_tick, snap_res, ring_rows, resolution, extra_cost_mode, max_actions, n_replicas;
    int res_is_one;
    double time_mean, time_std, supply_ratio, demand_ratio, scope_low, scope_high;
    int FW, FWp, CWp, QN, QH, SW, DW;  // DW = decision row words
    // table blob offsets
    int t_bikes, t_capacity, t_id, t_nbr_offset, t_nbr_idx, t_trip_offset, t_trip_src, t_trip_dst, t_trip_dur;
    int t_day_of_tick, t_day_feat, t_mt;
    int rng_words;  // per-replica RNG block: 624 state + idx + has_gauss + gauss (2 words), then the scope-filter area
    // decision.action_scope.filters (decision_strategy.py:15-163); scope-filter area inside the per-replica global block:
    // [tw_frame W][tw_cache W*S][scratch 4*S] at word offset `scope_off` (W = tw_windows = the widest trip-window filter)
    int n_filters, filter_type[4], filter_num[4], filter_windows[4];
    int drops;  // some filter can drop a neighbour (num < S - 1): the chain has to be run (otherwise it only permutes a dict)
    int scope_off, tw_windows;
    int o_scratch;  // 4 x S words of action-scope scratch inside the replica block (word offset from the frame)
};

struct BikeReplica {
    int32_t* f;
    int32_t* c;
    int32_t* q;
    const int32_t* t;
    uint32_t* rng;
    int32_t* snap;
    int32_t* snap_frame;
    int64_t seed;  // >= 0: this replica's own np.random seed for transfer_time (set at reset); < 0: the topology's transfer_seed
};

MARO_DEV int32_t& BA(const BikeShape& s, const BikeReplica& r, int attr, int i) { return r.f[attr * s.S + i]; }
MARO_DEV int64_t bctrl_get64(const BikeReplica& r, int lo) { return (int64_t)(((uint64_t)(uint32_t)r.c[lo + 1] << 32) | (uint32_t)r.c[lo]); }
MARO_DEV void bctrl_add64(const BikeReplica& r, int lo, int64_t d) {
    int64_t v = bctrl_get64(r, lo) + d;
    r.c[lo] = (int32_t)(uint32_t)((uint64_t)v & 0xffffffffu);
    r.c[lo + 1] = (int32_t)(uint32_t)((uint64_t)v >> 32);
}

// station.bikes = v  +  Station._on_bikes_changed (station.py:70-75)
MARO_DEV void set_bikes(const BikeShape& s, const BikeReplica& r, int i, int v) {
    BA(s, r, BA_BIKES, i) = v;
    if (v < BA(s, r, BA_MIN_BIKES, i)) BA(s, r, BA_MIN_BIKES, i) = v;
}

// queue layout: ev[QN][2] | bucket[QH] | nxt[QN] u16 | free[QN] u16  (same as the CIM calendar queue)
MARO_DEV int32_t* bq_bucket(const BikeShape& s, const BikeReplica& r) { return r.q + s.QN * 2; }
MARO_DEV uint16_t* bq_next(const BikeShape& s, const BikeReplica& r) { return reinterpret_cast<uint16_t*>(r.q + s.QN * 2 + s.QH); }
MARO_DEV uint16_t* bq_free(const BikeShape& s, const BikeReplica& r) { return bq_next(s, r) + s.QN; }

MARO_DEV void bike_push(const BikeShape& s, const BikeReplica& r, int tick, int type, int a, int b, int number) {
    if (tick < 0 || tick >= s.max_tick) return;  // never visited by the Env
    int top = r.c[BC_FREE_TOP];
    if (top <= 0) { r.c[BC_ERR] = -2; return; }
    int slot = bq_free(s, r)[top - 1];
    r.c[BC_FREE_TOP] = top - 1;
    r.c[BC_Q_COUNT] += 1;
    r.q[slot * 2] = type | (a << 8) | (b << 16);
    r.q[slot * 2 + 1] = number;
    bq_next(s, r)[slot] = Q_NIL;
    int32_t* bk = bq_bucket(s, r) + (tick & (s.QH - 1));
    int hb = *bk, tail = (hb >> 16) & 0xffff;
    if (tail == Q_NIL) *bk = slot | (slot << 16);
    else { bq_next(s, r)[tail] = (uint16_t)slot; *bk = (hb & 0xffff) | (slot << 16); }
}

// ---- numpy legacy RandomState: mt19937 + legacy_gauss (numpy/random/src/legacy/legacy-distributions.c)
MARO_DEV uint32_t bike_mt_next(const BikeReplica& r) {
    uint32_t* mt = r.rng;
    int idx = (int)r.rng[624];
    if (idx >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) mt[kk] = mt_mix(mt[kk], mt[kk + 1], mt[kk + 397]);
        for (; kk < 623; kk++) mt[kk] = mt_mix(mt[kk], mt[kk + 1], mt[kk - 227]);
        mt[623] = mt_mix(mt[623], mt[0], mt[396]);
        idx = 0;
    }
    uint32_t y = mt[idx++];
    r.rng[624] = (uint32_t)idx;
    return mt_temper(y);
}
// One iteration of legacy_gauss's rejection loop takes four words.  The stream position stays a multiple of four (624 = 4 x 156;
// seeding leaves it at 624), so an iteration never straddles a regeneration: either four loads that do not depend on each other
// (one L2 round trip on the leader's chain instead of four position-load -> word-load pairs), or the regeneration first.
MARO_DEV double bike_gauss(const BikeReplica& r) {
    // header words 624..627 = position | has_gauss | cached gauss (double), one 128-bit load
#ifdef MARO_HOST_EMULATION
    const uint32_t h0 = r.rng[624], h1 = r.rng[625], h2 = r.rng[626], h3 = r.rng[627];
#else
    const uint4 hv = *reinterpret_cast<const uint4*>(r.rng + 624);
    const uint32_t h0 = hv.x, h1 = hv.y, h2 = hv.z, h3 = hv.w;
#endif
    double* cache = reinterpret_cast<double*>(r.rng + 626);
    if (h1) {
        r.rng[625] = 0;
        *cache = 0.0;
        const uint64_t bits = ((uint64_t)h3 << 32) | h2;
        double t;
        memcpy(&t, &bits, 8);
        return t;
    }
    int idx = (int)h0;
    double f, x1, x2, r2;
    do {
        uint32_t y0, y1, y2, y3;
        if ((idx & 3) == 0 && idx + 4 <= 624) {
            const uint32_t* mt = r.rng + idx;
            y0 = mt[0]; y1 = mt[1]; y2 = mt[2]; y3 = mt[3];
            y0 = mt_temper(y0); y1 = mt_temper(y1); y2 = mt_temper(y2); y3 = mt_temper(y3);
            idx += 4;
        } else {  // regeneration due (or a foreign stream position): word by word through the generic path
            r.rng[624] = (uint32_t)idx;
            y0 = bike_mt_next(r); y1 = bike_mt_next(r); y2 = bike_mt_next(r); y3 = bike_mt_next(r);
            idx = (int)r.rng[624];
        }
        x1 = 2.0 * (((double)(y0 >> 5) * 67108864.0 + (double)(y1 >> 6)) / 9007199254740992.0) - 1.0;
        x2 = 2.0 * (((double)(y2 >> 5) * 67108864.0 + (double)(y3 >> 6)) / 9007199254740992.0) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    *cache = f * x1;
    r.rng[624] = (uint32_t)idx;
    r.rng[625] = 1;
    return f * x2;
}

// the regeneration of the 624 state words by all lanes of the group (same recurrence, G words at a time: the first 227 words mix
// old words only, later ones reach back 227 words to new ones — further than G), when the stream stands exactly at its end
template <int G>
MARO_DEV void bike_mt_regenerate(const Grp<G>& g, const BikeReplica& r) {
    if (r.rng[624] != 624u) return;  // (group-uniform: every lane reads the same word)
    uint32_t* mt = r.rng;
    g.sync();
    for (int base = 0; base < 623; base += G) {
        const int kk = base + g.lane;
        uint32_t nv = 0;
        if (kk < 623) nv = mt_mix(mt[kk], mt[kk + 1], kk < 227 ? mt[kk + 397] : mt[kk - 227]);
        g.sync();  // every lane has read its (old) right-hand neighbour before anyone stores
        if (kk < 623) mt[kk] = nv;
        g.sync();
    }
    if (g.lane == 0) {
        mt[623] = mt_mix(mt[623], mt[0], mt[396]);
        r.rng[624] = 0;
    }
    g.sync();
}

// BikeDecisionStrategy.move_to_neighbor (decision_strategy.py:295-343)
MARO_DEV void move_to_neighbor(const BikeShape& s, const BikeReplica& r, int src, int cur, int bike_number) {
    int lo = r.t[s.t_nbr_offset + cur], hi = r.t[s.t_nbr_offset + cur + 1];
    for (int k = lo, order = 0; k < hi; k++, order++) {
        int n = r.t[s.t_nbr_idx + k];
        int nb_bikes = BA(s, r, BA_BIKES, n);
        int accept = BA(s, r, BA_CAPACITY, n) - nb_bikes;
        if (accept > bike_number) accept = bike_number;
        set_bikes(s, r, n, nb_bikes + accept);
        int cost = accept * (order + 1);
        int who = s.extra_cost_mode == 0 ? src : (s.extra_cost_mode == 1 ? cur : n);
        BA(s, r, BA_EXTRA_COST, who) += cost;
        bike_number -= accept;
        if (bike_number == 0) break;
    }
}

MARO_DEV void on_required_bike(const BikeShape& s, const BikeReplica& r, int tick, int src, int dst, int dur) {  // :398-437
    int bikes = BA(s, r, BA_BIKES, src);
    BA(s, r, BA_TRIP_REQUIREMENT, src) += 1;
    bctrl_add64(r, BC_TRIPS_LO, 1);
    r.f[BA_COUNT * s.S + src * s.S + dst] += 1;
    if (bikes < 1) {
        BA(s, r, BA_SHORTAGE, src) += 1;
        bctrl_add64(r, BC_SHORT_LO, 1);
    } else {
        BA(s, r, BA_FULFILLMENT, src) += 1;
        set_bikes(s, r, src, bikes - 1);
        bike_push(s, r, tick + dur, BE_RETURN, src, dst, 1);
    }
}
MARO_DEV void on_bike_returned(const BikeShape& s, const BikeReplica& r, int from, int to, int n) {  // :439-466
    int bikes = BA(s, r, BA_BIKES, to);
    int empty_docks = BA(s, r, BA_CAPACITY, to) - bikes;
    int acc = empty_docks < n ? empty_docks : n;
    if (acc < n) {
        BA(s, r, BA_FAILED_RETURN, to) += n - acc;
        move_to_neighbor(s, r, from, to, n - acc);
    }
    set_bikes(s, r, to, bikes + acc);
}
MARO_DEV void on_bike_deliver(const BikeShape& s, const BikeReplica& r, int from, int to, int n) {  // :494-519
    int bikes = BA(s, r, BA_BIKES, to);
    int empty_docks = BA(s, r, BA_CAPACITY, to) - bikes;
    int acc = empty_docks < n ? empty_docks : n;
    if (acc < n) move_to_neighbor(s, r, from, to, n - acc);
    if (acc > 0) {
        BA(s, r, BA_TRANSFER_COST, to) += acc;
        bctrl_add64(r, BC_OPNUM_LO, acc);
    }
    set_bikes(s, r, to, bikes + acc);
}
// _on_rebalance_bikes (:468-492) + get_stations_need_decision (decision_strategy.py:229-251)
MARO_DEV void on_rebalance(const BikeShape& s, const BikeReplica& r, int tick) {
    for (int i = 0; i < s.S; i++) {
        double ratio = (double)BA(s, r, BA_BIKES, i) / (double)BA(s, r, BA_CAPACITY, i);
        int type = -1;
        if (ratio >= s.supply_ratio) type = 0;
        else if (ratio <= s.demand_ratio) type = 1;
        if (type >= 0) bike_push(s, r, tick, BE_DECISION, i, type, 0);
    }
}
// _on_action_received (:521-559)
template <int G>
MARO_DEV void bike_on_actions(const BikeShape& s, const Grp<G>& g, const BikeReplica& r, int tick, const Act4& mine, int n) {
    // The decision event being answered has left this tick's list already.  If it was the LAST element, the reference's list
    // keeps its tail pointer on the removed event (EventLinkedList._extract_sub_events, event_linked_list.py:86-92, does not
    // update `_tail`), so events appended to this tick's list from now on — deliveries with transfer time 0 — are lost.
    const bool tail_lost = (*(bq_bucket(s, r) + (tick & (s.QH - 1))) & 0xffff) == Q_NIL;
    if (n > 0) bike_mt_regenerate(g, r);  // (a stream that stands at its end: regenerate with all lanes before the leader draws)
    for (int i = 0; i < n; i++) {
        int from = g.shfl(mine.v, i), to = g.shfl(mine.p, i), number = g.shfl(mine.qty, i);
        if (g.lane != 0) continue;
        if (from < 0 || to < 0 || from >= s.S || to >= s.S) continue;
        int bikes = BA(s, r, BA_BIKES, from);
        int executed = bikes < number ? bikes : number;
        if (executed > 0) {
            set_bikes(s, r, from, bikes - executed);
            // transfer_time = round(np.random.normal(mean, scale=std))   (decision_strategy.py:213-216)
            double x = s.time_mean + s.time_std * bike_gauss(r);
            int tt = (int)rint(x);  // python round(): half to even
            // a negative transfer time (normal(20, 5) can produce one) files the event under a tick that has already
            // been executed: the reference never runs it and the bikes are lost (event_buffer.py:166-175)
            if (tt > 0 || (tt == 0 && !tail_lost)) bike_push(s, r, tick + tt, BE_DELIVER, from, to, executed);
        }
    }
}

// sorted(items, key=(value, index), reverse)[:out] by selection: position j takes the best remaining entry (keys are unique)
MARO_DEV void bike_select_top(int n, int out, int32_t* idx, int32_t* val, int32_t* key, bool reverse) {
    for (int j = 0; j < out; j++) {
        int best = j;
        for (int k = j + 1; k < n; k++) {
            bool less = key[k] < key[best] || (key[k] == key[best] && idx[k] < idx[best]);
            if (reverse ? !less : less) best = k;
        }
        if (best != j) {
            int t = idx[j]; idx[j] = idx[best]; idx[best] = t;
            t = val[j]; val[j] = val[best]; val[best] = t;
            t = key[j]; key[j] = key[best]; key[best] = t;
        }
    }
}

// BikeDecisionStrategy.action_scope (decision_strategy.py:253-293): neighbour scope -> filter chain (DistanceFilter :15-52,
// RequirementsFilter :55-88, TripsWindowFilter :91-166 incl. its per-frame cache) -> the station itself; pairs in station
// order (the reference's dict carries no order).  Leader lane only; the pre-decision snapshot of this tick has been taken.
MARO_DEV int bike_action_scope(const BikeShape& s, const BikeReplica& r, int station, int type, int frame_now, int32_t* pairs) {
    const int lo = r.t[s.t_nbr_offset + station], hi = r.t[s.t_nbr_offset + station + 1];
    const int S = s.S;
    int32_t* area = reinterpret_cast<int32_t*>(r.rng) + s.scope_off;
    int32_t* tw_frame = area;
    int32_t* tw_cache = area + s.tw_windows;
    int32_t* idx = r.f + s.o_scratch;  // (shared memory, behind the event queue)
    int32_t* val = idx + S;
    int32_t* key = val + S;
    int32_t* res = key + S;  // value per station, -1 = not in the scope
    for (int i = 0; i < S; i++) res[i] = -1;
    int n = 0;
    for (int k = lo; k < hi; k++) {
        const int nb = r.t[s.t_nbr_idx + k];
        const int bikes = BA(s, r, BA_BIKES, nb), cap = BA(s, r, BA_CAPACITY, nb);
        idx[n] = nb;
        val[n] = type == 0 ? cap - bikes : (int)maro_floor((double)bikes * s.scope_high);
        n++;
    }
    if (s.drops) {
        for (int f = 0; f < s.n_filters; f++) {
            const int out = s.filter_num[f] < n ? s.filter_num[f] : n;
            if (s.filter_type[f] == 0) {
                // the `out` nearest of the station's FULL neighbour list, with their values from the current scope
                for (int k = 0; k < out; k++) {
                    const int nb = r.t[s.t_nbr_idx + lo + k];
                    int at = -1;
                    for (int j = k; j < n; j++) if (idx[j] == nb) { at = j; break; }
                    if (at < 0) { r.c[BC_ERR] = -3; break; }  // KeyError in the reference (a filter before it dropped `nb`)
                    int t = idx[k]; idx[k] = idx[at]; idx[at] = t;
                    t = val[k]; val[k] = val[at]; val[at] = t;
                }
            } else if (s.filter_type[f] == 1) {
                for (int j = 0; j < n; j++) key[j] = val[j];
                bike_select_top(n, out, idx, val, key, true);
            } else {
                // the latest `windows` frames the ring holds; frames are consecutive, the newest is the pre-decision snapshot
                const int W = s.filter_windows[f];
                int first = frame_now - (s.ring_rows - 1);
                if (first < 0) first = 0;
                int held = 0;
                for (int fr = first; fr <= frame_now; fr++) held += r.snap_frame[fr < s.ring_rows ? fr : fr % s.ring_rows] == fr;
                const int avail = W < held ? W : held;
                for (int j = 0; j < n; j++) key[j] = 0;
                int taken = 0;
                for (int fr = frame_now; fr >= first && taken < avail; fr--) {
                    const int row = fr < s.ring_rows ? fr : fr % s.ring_rows;
                    if (r.snap_frame[row] != fr) continue;
                    taken++;
                    int32_t* c = tw_cache + (fr % s.tw_windows) * S;
                    if (fr == frame_now || tw_frame[fr % s.tw_windows] != fr) {  // latest: always re-read; others: once
                        const int32_t* src = r.snap + (int64_t)row * s.FWp + BA_TRIP_REQUIREMENT * S;
                        for (int i = 0; i < S; i++) c[i] = src[i];
                        tw_frame[fr % s.tw_windows] = fr;
                    }
                    for (int j = 0; j < n; j++) key[j] += c[idx[j]];
                }
                bike_select_top(n, out, idx, val, key, type == 1);
            }
            n = out;
        }
    }
    for (int j = 0; j < n; j++) res[idx[j]] = val[j];
    {
        const int bikes = BA(s, r, BA_BIKES, station), cap = BA(s, r, BA_CAPACITY, station);
        res[station] = type == 0 ? (int)maro_floor((double)bikes * (1.0 - s.scope_low)) : cap - bikes;
    }
    int m = 0;
    for (int i = 0; i < S; i++)
        if (res[i] >= 0 || i == station) { pairs[2 * m] = i; pairs[2 * m + 1] = res[i]; m++; }
    return m;
}

template <int G>
MARO_DEV void bike_snapshot(const BikeShape& s, const Grp<G>& g, const BikeReplica& r, int frame_index) {
    g.sync();
    int row = frame_index < s.ring_rows ? frame_index : frame_index % s.ring_rows;
    int32_t* dst = r.snap + (int64_t)row * s.FWp;
#ifdef MARO_HOST_EMULATION
    LANE_LOOP(i, s.FWp) dst[i] = r.f[i];
#else
    const int4* src4 = reinterpret_cast<const int4*>(r.f);
    int4* dst4 = reinterpret_cast<int4*>(dst);
    LANE_LOOP(i, s.FWp / 4) dst4[i] = src4[i];
#endif
    if (g.lane == 0) { r.snap_frame[row] = frame_index; bctrl_add64(r, BC_NSNAPS_LO, 1); }
    g.sync();
}
MARO_DEV int bike_frame_index(const BikeShape& s, int tick) { return s.res_is_one ? tick - s.start_tick : (tick - s.start_tick) / s.snap_res; }

// Drain bucket `tick` on the leader lane.  Returns true when it stopped at a decision event (left in BC_PEND_*).
MARO_DEV bool bike_drain(const BikeShape& s, const BikeReplica& r, int tick, int& nev) {
    int32_t* bk = bq_bucket(s, r) + (tick & (s.QH - 1));
    for (;;) {
        int hb = *bk, slot = hb & 0xffff;
        if (slot == Q_NIL) return false;
        int nxt = bq_next(s, r)[slot];
        *bk = nxt == Q_NIL ? (Q_NIL | (Q_NIL << 16)) : (nxt | (hb & 0xffff0000));
        int w0 = r.q[slot * 2], number = r.q[slot * 2 + 1];
        int top = r.c[BC_FREE_TOP];
        bq_free(s, r)[top] = (uint16_t)slot;
        r.c[BC_FREE_TOP] = top + 1;
        r.c[BC_Q_COUNT] -= 1;
        int type = w0 & 0xff, a = (w0 >> 8) & 0xff, b = (w0 >> 16) & 0xff;
        if (type == BE_DECISION) { r.c[BC_PEND_STATION] = a; r.c[BC_PEND_TYPE] = b; return true; }
        if (type == BE_RETURN) on_bike_returned(s, r, a, b, number);
        else on_bike_deliver(s, r, a, b, number);
        nev++;
    }
}

template <int G>
MARO_DEV void bike_replica_step(const BikeShape& s, const Grp<G>& g, const BikeReplica& r, const Act4& act, int n_act,
                                int32_t* dec, int64_t* met) {
    int state = r.c[BC_STATE];
    if (state >= ST_DONE) {
        g.sync();
        if (g.lane == 0) {
            if (state == ST_DONE) r.c[BC_STATE] = ST_FINISHED;
            for (int i = 0; i < s.DW; i++) dec[i] = 0;
            dec[6] = 2;
            met[0] = met[1] = met[2] = 0;
        }
        g.sync();
        return;
    }
    int nev = 0, nticks = 0;
    int tick = r.c[BC_TICK];
    if (state == ST_AWAIT) {
        g.sync();
        bike_on_actions(s, g, r, tick, act, n_act);
        nev += 2;  // the decision event + TAKE_ACTION
        g.sync();
        state = ST_DECISIONS;  // continue draining the tick's list
    } else if (state == ST_START) {
        state = ST_TICK_BEGIN;
    }
    int status = 0;
    bool decided = false;
    for (;;) {
        if (state == ST_TICK_BEGIN) {
            nticks++;
            // _update_station_extra_features (:370-396): once per day, before the tick's events
            int day = r.t[s.t_day_of_tick + tick];
            if (day != r.c[BC_LAST_DAY]) {
                const int32_t* f = r.t + s.t_day_feat + 4 * day;
                LANE_LOOP(i, s.S) {
                    BA(s, r, BA_WEEKDAY, i) = f[0]; BA(s, r, BA_HOLIDAY, i) = f[1];
                    BA(s, r, BA_WEATHER, i) = f[2]; BA(s, r, BA_TEMPERATURE, i) = f[3];
                }
            }
            g.sync();
            if (g.lane == 0) {
                r.c[BC_LAST_DAY] = day;
                bike_drain(s, r, tick, nev);                                  // 1. queued by earlier ticks
                int lo = r.t[s.t_trip_offset + tick], hi = r.t[s.t_trip_offset + tick + 1];
                for (int k = lo; k < hi; k++) {                               // 2. this tick's trips
                    on_required_bike(s, r, tick, r.t[s.t_trip_src + k], r.t[s.t_trip_dst + k], r.t[s.t_trip_dur + k]);
                    nev++;
                }
                if ((tick + 1) % s.resolution == 0) { on_rebalance(s, r, tick); nev++; }  // 3.
            }
            g.sync();
            state = ST_DECISIONS;
        }
        // 4. rest of the tick's list; stops at a decision event
        g.sync();
        if (g.lane == 0) r.c[BC_PHASE] = bike_drain(s, r, tick, nev) ? 1 : 0;
        g.sync();
        if (r.c[BC_PHASE]) {
            bike_snapshot(s, g, r, bike_frame_index(s, tick));  // core.py:345
            decided = true;
            state = ST_AWAIT;
            break;
        }
        // post_step (:131-147)
        if (s.res_is_one || (tick + 1) % s.snap_res == 0) {
            bike_snapshot(s, g, r, bike_frame_index(s, tick));
            LANE_LOOP(i, s.S) {
                BA(s, r, BA_SHORTAGE, i) = 0; BA(s, r, BA_TRIP_REQUIREMENT, i) = 0; BA(s, r, BA_EXTRA_COST, i) = 0;
                BA(s, r, BA_TRANSFER_COST, i) = 0; BA(s, r, BA_FULFILLMENT, i) = 0; BA(s, r, BA_FAILED_RETURN, i) = 0;
                BA(s, r, BA_MIN_BIKES, i) = BA(s, r, BA_BIKES, i);
            }
            g.sync();
        }
        if (tick + 1 == s.max_tick) {
            if (!s.res_is_one && (tick + 1) % s.snap_res != 0) bike_snapshot(s, g, r, bike_frame_index(s, tick));
            state = ST_DONE;
            status = 1;
            break;
        }
        tick += 1;
        state = ST_TICK_BEGIN;
    }
    g.sync();
    if (g.lane == 0) {
        for (int i = 0; i < s.DW; i++) dec[i] = 0;
        dec[0] = tick;
        if (decided) {
            int st_i = r.c[BC_PEND_STATION], ty = r.c[BC_PEND_TYPE];
            dec[1] = st_i; dec[2] = bike_frame_index(s, tick); dec[3] = ty;
            dec[4] = bike_action_scope(s, r, st_i, ty, bike_frame_index(s, tick), dec + 8);
        }
        if (r.c[BC_ERR] == -2) { state = ST_ERROR; status = -2; }
        dec[6] = status;
        dec[7] = r.c[BC_EP_STEP];
        r.c[BC_EP_STEP] += 1;
        r.c[BC_STATE] = state;
        r.c[BC_TICK] = tick;
        bctrl_add64(r, BC_NSTEPS_LO, 1);
        bctrl_add64(r, BC_NTICKS_LO, nticks);
        bctrl_add64(r, BC_NEVENTS_LO, nev);
        met[0] = bctrl_get64(r, BC_TRIPS_LO); met[1] = bctrl_get64(r, BC_SHORT_LO); met[2] = bctrl_get64(r, BC_OPNUM_LO);
    }
    g.sync();
}

template <int G>
MARO_DEV void bike_replica_reset(const BikeShape& s, const Grp<G>& g, const BikeReplica& r) {
    LANE_LOOP(i, s.FWp) r.f[i] = 0;
    LANE_LOOP(i, s.CWp) if (i < BC_NSTEPS_LO || i > BC_NSNAPS_HI) r.c[i] = 0;
    g.sync();
    LANE_LOOP(i, s.S) {
        int b = r.t[s.t_bikes + i];
        BA(s, r, BA_CAPACITY, i) = r.t[s.t_capacity + i];
        BA(s, r, BA_BIKES, i) = b;
        BA(s, r, BA_MIN_BIKES, i) = b;
        BA(s, r, BA_ID, i) = r.t[s.t_id + i];
    }
    uint16_t* fs = bq_free(s, r);
    uint16_t* nx = bq_next(s, r);
    LANE_LOOP(i, s.QN) { r.q[2 * i] = 0; r.q[2 * i + 1] = 0; nx[i] = Q_NIL; fs[i] = (uint16_t)(s.QN - 1 - i); }
    LANE_LOOP(i, s.QH) bq_bucket(s, r)[i] = Q_NIL | (Q_NIL << 16);
    LANE_LOOP(i, s.ring_rows) r.snap_frame[i] = -1;
    if (r.seed < 0) {
        LANE_LOOP(i, 624) r.rng[i] = (uint32_t)r.t[s.t_mt + i];
    } else if (g.lane == 0) {  // numpy legacy seeding (mt19937_seed): a serial recurrence, once per reset
        uint32_t x = (uint32_t)r.seed;
        for (int i = 0; i < 624; i++) {
            r.rng[i] = x;
            x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i + 1u;
        }
    }
    g.sync();
    LANE_LOOP(i, s.tw_windows) (reinterpret_cast<int32_t*>(r.rng) + s.scope_off)[i] = -1;  // TripsWindowFilter.reset (:165-166)
    if (g.lane == 0) {
        r.rng[624] = 624; r.rng[625] = 0; r.rng[626] = 0; r.rng[627] = 0;
        r.c[BC_STATE] = ST_START;
        r.c[BC_TICK] = s.start_tick;
        r.c[BC_FREE_TOP] = s.QN;
        r.c[BC_LAST_DAY] = -1;
    }
    g.sync();
}

}  // namespace maro
