/*
 * bike_oracle.c — TEST INFRASTRUCTURE.  CPU restatement (plain C, one replica) of the reference's citi_bike
 * `Env.step` hot path.  Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Parity status: PINNED — tests/test_bike_oracle_golden.py checks it against traces of the unmodified reference
 * (tests/golden/gen_bike_golden.py; frozen toy.3s_4t dataset in tests/golden/bike_toy and the reference's own test
 * fixtures tests/data/citi_bike/case_{1,2}) and the known answers of tests/citi_bike/test_bike_scenario.py:65-211.
 *
 * Follows the reference's structure (per-tick linked lists of events, one handler per event type):
 *   Env._simulate / step / _assign_action        maro/simulator/core.py:92-133, 301-381
 *   EventBuffer.execute, EventLinkedList          maro/event_buffer/event_buffer.py:177-247, event_linked_list.py:53-137
 *   CitibikeBusinessEngine.step / post_step       maro/simulator/scenarios/citi_bike/business_engine.py:101-147
 *   handlers                                      business_engine.py:398-559
 *   BikeDecisionStrategy                          decision_strategy.py:166-397
 *   Station._on_bikes_changed                     station.py:70-75
 *   np.random.normal (legacy RandomState)         numpy/random/src/legacy/legacy-distributions.c (legacy_gauss),
 *                                                 mt19937 seeding numpy/random/src/mt19937/mt19937.c (mt19937_seed)
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -I../include bike_oracle.c -o _build/libbike_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "maro_b200.h"

/* ------------------------------------------------------------------ numpy legacy RandomState */
typedef struct { uint32_t mt[624]; int idx; int has_gauss; double gauss; } NpRng;

static void np_seed(NpRng* r, uint32_t seed) { /* mt19937_seed */
    for (int i = 0; i < 624; i++) {
        r->mt[i] = seed;
        seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
    }
    r->idx = 624; r->has_gauss = 0; r->gauss = 0.0;
}
static uint32_t np_u32(NpRng* r) {
    if (r->idx >= 624) {
        uint32_t* mt = r->mt;
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < 623; kk++) {
            uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk - 227] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        r->idx = 0;
    }
    uint32_t y = r->mt[r->idx++];
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
    return y;
}
static double np_double(NpRng* r) {
    uint32_t a = np_u32(r) >> 5, b = np_u32(r) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}
static double np_gauss(NpRng* r) { /* legacy_gauss */
    if (r->has_gauss) { r->has_gauss = 0; double t = r->gauss; r->gauss = 0.0; return t; }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * np_double(r) - 1.0;
        x2 = 2.0 * np_double(r) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    r->gauss = f * x1; r->has_gauss = 1;
    return f * x2;
}

/* ------------------------------------------------------------------ events */
enum { EV_REQUIRE, EV_RETURN, EV_REBALANCE, EV_DELIVER, EV_PENDING_DECISION, EV_TAKE_ACTION };
enum { ST_PENDING, ST_EXECUTING, ST_FINISHED };
typedef struct Event {
    int tick, type, state, cascade;
    int a, b, c; /* payload: REQUIRE src,dst,dur; RETURN/DELIVER from,to,number; DECISION station,type */
    int32_t actions[16][4];
    int n_actions;
    struct Event* next;
    struct Event* imm_head;
    struct Event* imm_tail;
    int imm_count;
} Event;
typedef struct { Event* head; Event* tail; int count; int tail_lost; } EvList;

typedef struct { int32_t bikes, shortage, trip_requirement, fulfillment, capacity, id, weekday, temperature, weather,
                 holiday, extra_cost, transfer_cost, failed_return, min_bikes; } Station;

typedef struct BikeOracle {
    MaroBikeTopology t;
    int S, start_tick, max_tick, snap_res, ring_rows;
    Station* st;
    int32_t* trips_adj;
    EvList* pending;
    int n_lists;
    NpRng rng;
    int64_t total_trips, total_shortages, total_operate;
    int last_day;
    int tick, gen_state, ep_step;
    Event* pending_decision;
    int64_t n_events, n_ticks, n_snapshots, n_steps;
    int frame_words;
    int32_t* ring;
    int32_t* ring_frame;
    /* TripsWindowFilter._window_states_cache (decision_strategy.py:106, 136-142): frame index -> trip_requirement of every
     * station as it stood when the frame was last queried as the LATEST one (or first queried at all) */
    int total_frames;
    int32_t* tw_cache;   /* [total_frames][S] */
    char* tw_has;        /* [total_frames] */
    int scope_error;     /* a filter asked for a neighbour an earlier filter had dropped: KeyError in the reference */
} BikeOracle;

static void* dup_arr(const void* p, size_t bytes) { void* q = malloc(bytes ? bytes : 1); if (bytes) memcpy(q, p, bytes); return q; }

static void set_bikes(Station* s, int v) { /* station.bikes = v  +  _on_bikes_changed (station.py:70-75) */
    s->bikes = v;
    if (v < s->min_bikes) s->min_bikes = v;
}

static Event* gen_event(int tick, int type, int cascade) {
    Event* e = (Event*)calloc(1, sizeof(Event));
    e->tick = tick; e->type = type; e->cascade = cascade;
    return e;
}
static void insert_event(BikeOracle* o, Event* e) {
    if (e->tick < 0 || e->tick >= o->n_lists) { free(e); return; }
    EvList* l = &o->pending[e->tick];
    e->next = NULL;
    if (l->tail_lost) {
        /* the reference's list keeps its `_tail` on a cascade event that has already been removed (see clear_finished):
         * whatever is appended now hangs off that dead node, is counted, and is never reached from the head again */
        l->count++;
        free(e);
        return;
    }
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
            /* EventLinkedList._extract_sub_events (event_linked_list.py:86-92) splices the immediate events in front of the
             * rest but does not touch `_tail`: when the cascade event was the LAST element of the list, `_tail` keeps pointing
             * at it after its removal, and events appended to this tick's list from then on (a DeliverBike with transfer time
             * 0 issued by the action of a tick's last decision event) are lost.  Found by tools/fuzz_cim_bike_parity.py. */
            if (!l->head) l->tail_lost = 1;
            e->imm_tail->next = l->head;
            if (!l->head) l->tail = e->imm_tail;
            l->head = e->imm_head;
            l->count += e->imm_count;
        }
        free(e);
    }
}

/* BikeDecisionStrategy.move_to_neighbor (decision_strategy.py:295-343) */
static void move_to_neighbor(BikeOracle* o, int src, int cur, int bike_number) {
    for (int k = o->t.nbr_offset[cur], order = 0; k < o->t.nbr_offset[cur + 1]; k++, order++) {
        int n = o->t.nbr_idx[k];
        Station* nb = &o->st[n];
        int nb_bikes = nb->bikes;
        int accept = nb->capacity - nb_bikes;
        if (accept > bike_number) accept = bike_number;
        set_bikes(nb, nb_bikes + accept);
        int cost = accept * (order + 1);
        if (o->t.extra_cost_mode == 0) o->st[src].extra_cost += cost;
        else if (o->t.extra_cost_mode == 1) o->st[cur].extra_cost += cost;
        else nb->extra_cost += cost;
        bike_number -= accept;
        if (bike_number == 0) break;
    }
}

static void on_required_bike(BikeOracle* o, Event* ev) { /* :398-437 */
    Station* s = &o->st[ev->a];
    int bikes = s->bikes;
    s->trip_requirement += 1;
    o->total_trips += 1;
    o->trips_adj[ev->a * o->S + ev->b] += 1;
    if (bikes < 1) { s->shortage += 1; o->total_shortages += 1; }
    else {
        s->fulfillment += 1;
        set_bikes(s, bikes - 1);
        Event* e = gen_event(ev->tick + ev->c, EV_RETURN, 0);
        e->a = ev->a; e->b = ev->b; e->c = 1;
        insert_event(o, e);
    }
}
static void on_bike_returned(BikeOracle* o, Event* ev) { /* :439-466 */
    Station* s = &o->st[ev->b];
    int bikes = s->bikes, n = ev->c;
    int empty_docks = s->capacity - bikes;
    int acc = empty_docks < n ? empty_docks : n;
    if (acc < n) {
        s->failed_return += n - acc;
        move_to_neighbor(o, ev->a, ev->b, n - acc);
    }
    set_bikes(s, bikes + acc);
}
static void on_rebalance(BikeOracle* o, Event* ev) { /* :468-492, decision_strategy.py:229-251 */
    if ((ev->tick + 1) % o->t.resolution != 0) return;
    for (int i = 0; i < o->S; i++) {
        double ratio = (double)o->st[i].bikes / (double)o->st[i].capacity;
        int type = -1;
        if (ratio >= o->t.supply_ratio) type = 0;
        else if (ratio <= o->t.demand_ratio) type = 1;
        if (type >= 0) {
            Event* d = gen_event(ev->tick, EV_PENDING_DECISION, 1);
            d->a = i; d->b = type;
            insert_event(o, d);
        }
    }
}
static void on_bike_deliver(BikeOracle* o, Event* ev) { /* :494-519 */
    Station* s = &o->st[ev->b];
    int bikes = s->bikes, n = ev->c;
    int empty_docks = s->capacity - bikes;
    int acc = empty_docks < n ? empty_docks : n;
    if (acc < n) move_to_neighbor(o, ev->a, ev->b, n - acc);
    if (acc > 0) { s->transfer_cost += acc; o->total_operate += acc; }
    set_bikes(s, bikes + acc);
}
static void on_action_received(BikeOracle* o, Event* ev) { /* :521-559 */
    for (int i = 0; i < ev->n_actions; i++) {
        int from = ev->actions[i][0], to = ev->actions[i][1], number = ev->actions[i][2];
        if (from < 0 || to < 0 || from >= o->S || to >= o->S) continue;
        Station* s = &o->st[from];
        int bikes = s->bikes;
        int executed = bikes < number ? bikes : number;
        if (executed > 0) {
            set_bikes(s, bikes - executed);
            /* transfer_time = round(np.random.normal(mean, scale=std))  (decision_strategy.py:213-216) */
            double x = o->t.time_mean + o->t.time_std * np_gauss(&o->rng);
            int transfer_time = (int)nearbyint(x); /* python round(): half to even */
            Event* e = gen_event(ev->tick + transfer_time, EV_DELIVER, 0);
            e->a = from; e->b = to; e->c = executed;
            insert_event(o, e);
        }
    }
}
static void dispatch(BikeOracle* o, Event* e) {
    switch (e->type) {
        case EV_REQUIRE: on_required_bike(o, e); break;
        case EV_RETURN: on_bike_returned(o, e); break;
        case EV_REBALANCE: on_rebalance(o, e); break;
        case EV_DELIVER: on_bike_deliver(o, e); break;
        case EV_TAKE_ACTION: on_action_received(o, e); break;
        default: break;
    }
}
static Event* execute(BikeOracle* o, int tick) {
    EvList* l = &o->pending[tick];
    while (l->count) {
        clear_finished(l);
        Event* e = l->head;
        if (!e) break;
        if (e->type == EV_PENDING_DECISION && e->state != ST_EXECUTING) return e;
        e->state = ST_EXECUTING;
        dispatch(o, e);
        e->state = ST_FINISHED;
        o->n_events++;
    }
    return NULL;
}

/* canonical frame words: stations attr-major, 14 attrs alphabetical x S, then trips_adj[S*S] */
static void export_frame(const BikeOracle* o, int32_t* w) {
    int S = o->S;
    for (int i = 0; i < S; i++) {
        const Station* s = &o->st[i];
        int32_t v[14] = {s->bikes, s->capacity, s->extra_cost, s->failed_return, s->fulfillment, s->holiday, s->id,
                         s->min_bikes, s->shortage, s->temperature, s->transfer_cost, s->trip_requirement, s->weather,
                         s->weekday};
        for (int a = 0; a < 14; a++) w[a * S + i] = v[a];
    }
    memcpy(w + 14 * S, o->trips_adj, sizeof(int32_t) * S * S);
}
static void take_snapshot(BikeOracle* o, int frame_index) {
    int row = frame_index % o->ring_rows;
    export_frame(o, o->ring + (size_t)row * o->frame_words);
    o->ring_frame[row] = frame_index;
    o->n_snapshots++;
}
static int frame_index_of(const BikeOracle* o, int tick) { return (tick - o->start_tick) / o->snap_res; }

static void be_step(BikeOracle* o, int tick) { /* :101-129 */
    for (int k = o->t.trip_offset[tick]; k < o->t.trip_offset[tick + 1]; k++) {
        Event* e = gen_event(tick, EV_REQUIRE, 0);
        e->a = o->t.trip_src[k]; e->b = o->t.trip_dst[k]; e->c = o->t.trip_dur[k];
        insert_event(o, e);
    }
    if ((tick + 1) % o->t.resolution == 0) insert_event(o, gen_event(tick, EV_REBALANCE, 0));
    int day = o->t.day_of_tick[tick];
    if (day != o->last_day) { /* _update_station_extra_features :370-396 */
        o->last_day = day;
        const int32_t* f = o->t.day_feat + 4 * day;
        for (int i = 0; i < o->S; i++) { o->st[i].weekday = f[0]; o->st[i].holiday = f[1]; o->st[i].weather = f[2]; o->st[i].temperature = f[3]; }
    }
}
static int be_post_step(BikeOracle* o, int tick) { /* :131-147 */
    if ((tick + 1) % o->snap_res == 0) {
        take_snapshot(o, frame_index_of(o, tick));
        for (int i = 0; i < o->S; i++) {
            Station* s = &o->st[i];
            s->shortage = 0; s->trip_requirement = 0; s->extra_cost = 0; s->transfer_cost = 0; s->fulfillment = 0;
            s->failed_return = 0; s->min_bikes = s->bikes;
        }
    }
    return tick + 1 == o->max_tick;
}

static void free_events(BikeOracle* o) {
    for (int i = 0; i < o->n_lists; i++) {
        Event* e = o->pending[i].head;
        while (e) {
            Event* n = e->next;
            Event* s = e->imm_head;
            for (int k = 0; k < e->imm_count && s; k++) { Event* sn = s->next; free(s); s = sn; }
            free(e);
            e = n;
        }
        o->pending[i].head = o->pending[i].tail = NULL; o->pending[i].count = 0; o->pending[i].tail_lost = 0;
    }
}

void bike_oracle_reset(BikeOracle* o) { /* core.py:143-170, business_engine.py:171-196 */
    free_events(o);
    for (int i = 0; i < o->S; i++) {
        memset(&o->st[i], 0, sizeof(Station));
        o->st[i].capacity = o->t.station_capacity[i];
        o->st[i].min_bikes = 0; /* frame.reset() zeroes, then Station.reset(): capacity, bikes (cb: min(bikes, 0)=0), min_bikes, id */
        o->st[i].bikes = o->t.station_bikes[i];
        if (o->st[i].bikes < o->st[i].min_bikes) o->st[i].min_bikes = o->st[i].bikes;
        o->st[i].min_bikes = o->t.station_bikes[i];
        o->st[i].id = o->t.station_id[i];
    }
    memset(o->trips_adj, 0, sizeof(int32_t) * o->S * o->S);
    for (int i = 0; i < o->ring_rows; i++) o->ring_frame[i] = -1;
    memset(o->tw_has, 0, (size_t)o->total_frames); /* BikeDecisionStrategy.reset -> TripsWindowFilter.reset (:165-166) */
    o->scope_error = 0;
    np_seed(&o->rng, o->t.transfer_seed);
    o->total_trips = o->total_shortages = o->total_operate = 0;
    o->last_day = -1;
    o->tick = o->start_tick; o->gen_state = 0; o->ep_step = 0; o->pending_decision = NULL;
}

BikeOracle* bike_oracle_create(const MaroBikeTopology* t, int start_tick, int snapshot_resolution, int max_snapshots) {
    BikeOracle* o = (BikeOracle*)calloc(1, sizeof(BikeOracle));
    o->t = *t;
    int S = t->n_stations;
    o->S = S;
    int ntrips = t->trip_offset[t->max_tick];
#define DUP(f, n) o->t.f = (const int32_t*)dup_arr(t->f, sizeof(int32_t) * (size_t)(n))
    DUP(station_bikes, S); DUP(station_capacity, S); DUP(station_id, S); DUP(nbr_offset, S + 1); DUP(nbr_idx, t->nbr_offset[S]);
    DUP(trip_offset, t->max_tick + 1); DUP(trip_src, ntrips); DUP(trip_dst, ntrips); DUP(trip_dur, ntrips);
    DUP(day_of_tick, t->max_tick); DUP(day_feat, 4 * t->n_days);
#undef DUP
    o->start_tick = start_tick; o->max_tick = t->max_tick; o->snap_res = snapshot_resolution;
    int total_frames = (o->max_tick - start_tick + snapshot_resolution - 1) / snapshot_resolution;
    o->ring_rows = max_snapshots > 0 && max_snapshots < total_frames ? max_snapshots : total_frames;
    if (o->ring_rows < 1) o->ring_rows = 1;
    o->st = (Station*)calloc(S, sizeof(Station));
    o->trips_adj = (int32_t*)calloc(S * S, 4);
    o->n_lists = o->max_tick + 1;
    o->pending = (EvList*)calloc(o->n_lists, sizeof(EvList));
    o->frame_words = 14 * S + S * S;
    o->ring = (int32_t*)calloc((size_t)o->ring_rows * o->frame_words, 4);
    o->ring_frame = (int32_t*)calloc(o->ring_rows, 4);
    o->total_frames = total_frames > 0 ? total_frames : 1;
    o->tw_cache = (int32_t*)calloc((size_t)o->total_frames * S, 4);
    o->tw_has = (char*)calloc((size_t)o->total_frames, 1);
    bike_oracle_reset(o);
    return o;
}
void bike_oracle_destroy(BikeOracle* o) {
    if (!o) return;
    free_events(o);
    free(o->pending); free(o->ring); free(o->ring_frame); free(o->st); free(o->trips_adj); free(o->tw_cache); free(o->tw_has);
    free((void*)o->t.station_bikes); free((void*)o->t.station_capacity); free((void*)o->t.station_id);
    free((void*)o->t.nbr_offset); free((void*)o->t.nbr_idx); free((void*)o->t.trip_offset); free((void*)o->t.trip_src);
    free((void*)o->t.trip_dst); free((void*)o->t.trip_dur); free((void*)o->t.day_of_tick); free((void*)o->t.day_feat);
    free(o);
}

/* sorted(items, key=lambda kv: (kv[1], kv[0]), reverse=rev)[:out]  on parallel arrays (idx, val, key); keys are unique pairs */
static void sort_by_key(int n, int32_t* idx, int32_t* val, int64_t* key, int reverse) {
    for (int i = 1; i < n; i++) { /* insertion sort: n <= stations */
        int32_t ii = idx[i], vv = val[i];
        int64_t kk = key[i];
        int j = i - 1;
        while (j >= 0) {
            /* order of (key[j], idx[j]) vs (kk, ii) */
            int less = key[j] < kk || (key[j] == kk && idx[j] < ii);
            int wrong = reverse ? less : !less;
            if (!wrong) break;
            idx[j + 1] = idx[j]; val[j + 1] = val[j]; key[j + 1] = key[j];
            j--;
        }
        idx[j + 1] = ii; val[j + 1] = vv; key[j + 1] = kk;
    }
}

/* BikeDecisionStrategy.action_scope (decision_strategy.py:253-293): neighbour scope, the filter chain
 * (DistanceFilter :15-52, RequirementsFilter :55-88, TripsWindowFilter :91-166), then the station itself.
 * `pairs` = (station, value) in ascending station order (the reference returns a dict: order carries no meaning). */
static int action_scope(BikeOracle* o, int station, int type, int32_t* pairs) {
    const int S = o->S;
    int32_t idx[4096], val[4096];
    int64_t key[4096];
    int n = 0;
    for (int k = o->t.nbr_offset[station]; k < o->t.nbr_offset[station + 1]; k++) {
        int nb = o->t.nbr_idx[k];
        const Station* s = &o->st[nb];
        idx[n] = nb;
        val[n] = type == 0 ? s->capacity - s->bikes : (int)floor(s->bikes * o->t.scope_high);
        n++;
    }
    for (int f = 0; f < o->t.n_filters; f++) {
        int out = o->t.filter_num[f] < n ? o->t.filter_num[f] : n; /* output_num = min(num, len(source)) */
        if (o->t.filter_type[f] == MARO_BIKE_FILTER_DISTANCE) {
            /* result[n] = source[n] for the first `out` entries of the station's FULL neighbour list */
            int32_t nidx[4096], nval[4096];
            for (int k = 0; k < out; k++) {
                int nb = o->t.nbr_idx[o->t.nbr_offset[station] + k], at = -1;
                for (int j = 0; j < n; j++) if (idx[j] == nb) at = j;
                if (at < 0) { o->scope_error = 1; nval[k] = 0; } else nval[k] = val[at];
                nidx[k] = nb;
            }
            memcpy(idx, nidx, sizeof(int32_t) * (size_t)out); memcpy(val, nval, sizeof(int32_t) * (size_t)out);
        } else if (o->t.filter_type[f] == MARO_BIKE_FILTER_REQUIREMENTS) {
            for (int j = 0; j < n; j++) key[j] = val[j];
            sort_by_key(n, idx, val, key, 1);
        } else {
            /* frames held by the snapshot list in insertion (= ascending) order; the latest `available_windows` of them */
            int frames[4096], nf = 0;
            int first = o->total_frames, last = -1;
            for (int r = 0; r < o->ring_rows; r++) if (o->ring_frame[r] >= 0) { if (o->ring_frame[r] < first) first = o->ring_frame[r]; if (o->ring_frame[r] > last) last = o->ring_frame[r]; }
            for (int fr = first; fr <= last && nf < 4096; fr++) if (o->ring_frame[fr % o->ring_rows] == fr) frames[nf++] = fr;
            int avail = o->t.filter_windows[f] < nf ? o->t.filter_windows[f] : nf;
            for (int j = 0; j < n; j++) key[j] = 0;
            for (int i = 0; i < avail; i++) {
                int fr = frames[nf - avail + i];
                if (i == avail - 1 || !o->tw_has[fr]) { /* the latest frame may still change: always re-read; others once */
                    const int32_t* row = o->ring + (size_t)(fr % o->ring_rows) * o->frame_words;
                    memcpy(o->tw_cache + (size_t)fr * S, row + 11 * S, sizeof(int32_t) * (size_t)S); /* trip_requirement */
                    o->tw_has[fr] = 1;
                }
                for (int j = 0; j < n; j++) key[j] += o->tw_cache[(size_t)fr * S + idx[j]];
            }
            sort_by_key(n, idx, val, key, type == 1); /* Demand: most trips first; Supply: fewest first */
        }
        n = out;
    }
    int32_t res[4096];
    char has[4096];
    memset(has, 0, (size_t)S);
    for (int j = 0; j < n; j++) { res[idx[j]] = val[j]; has[idx[j]] = 1; }
    const Station* s = &o->st[station];
    res[station] = type == 0 ? (int)floor(s->bikes * (1 - o->t.scope_low)) : s->capacity - s->bikes;
    has[station] = 1;
    int m = 0;
    for (int i = 0; i < S; i++) if (has[i]) { pairs[2 * m] = i; pairs[2 * m + 1] = res[i]; m++; }
    return m;
}

static void fill_metrics(const BikeOracle* o, int64_t* m) { m[0] = o->total_trips; m[1] = o->total_shortages; m[2] = o->total_operate; }

int bike_oracle_decision_words(const BikeOracle* o) { return MARO_BIKE_DEC_HEAD + 2 * o->S; }

int bike_oracle_step(BikeOracle* o, const int32_t* actions, int n_actions, int32_t* decision, int64_t* metrics) {
    memset(decision, 0, sizeof(int32_t) * (size_t)bike_oracle_decision_words(o));
    if (o->gen_state >= 2) {
        o->gen_state = 3;
        decision[MARO_BIKE_DEC_STATUS] = MARO_STATUS_FINISHED;
        metrics[0] = metrics[1] = metrics[2] = 0;
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
        if (d) {
            take_snapshot(o, frame_index_of(o, o->tick));
            o->pending_decision = d;
            decision[MARO_BIKE_DEC_TICK] = o->tick;
            decision[MARO_BIKE_DEC_STATION] = d->a;
            decision[MARO_BIKE_DEC_FRAME_INDEX] = frame_index_of(o, o->tick);
            decision[MARO_BIKE_DEC_TYPE] = d->b;
            decision[MARO_BIKE_DEC_N_SCOPE] = action_scope(o, d->a, d->b, decision + MARO_BIKE_DEC_HEAD);
            decision[MARO_BIKE_DEC_STATUS] = MARO_STATUS_DECISION;
            decision[MARO_BIKE_DEC_STEP] = o->ep_step++;
            fill_metrics(o, metrics);
            return MARO_STATUS_DECISION;
        }
        if (be_post_step(o, o->tick)) break;
        o->tick += 1;
    }
    if ((o->tick + 1) % o->snap_res != 0) take_snapshot(o, frame_index_of(o, o->tick));
    o->gen_state = 2;
    decision[MARO_BIKE_DEC_TICK] = o->tick;
    decision[MARO_BIKE_DEC_STATUS] = MARO_STATUS_DONE;
    decision[MARO_BIKE_DEC_STEP] = o->ep_step++;
    fill_metrics(o, metrics);
    return MARO_STATUS_DONE;
}

int bike_oracle_frame_words(const BikeOracle* o) { return o->frame_words; }
void bike_oracle_read_frame(const BikeOracle* o, int32_t* out) { export_frame(o, out); }
int bike_oracle_tick(const BikeOracle* o) { return o->tick; }
void bike_oracle_counters(const BikeOracle* o, int64_t* out) { out[0] = o->n_steps; out[1] = o->n_ticks; out[2] = o->n_events; out[3] = o->n_snapshots; }
int bike_oracle_read_snapshot(const BikeOracle* o, int frame_index, int32_t* out) {
    if (frame_index < 0) return 0;
    int row = frame_index % o->ring_rows;
    if (o->ring_frame[row] != frame_index) return 0;
    memcpy(out, o->ring + (size_t)row * o->frame_words, 4 * (size_t)o->frame_words);
    return 1;
}

/* greedy top-1 policy (examples/citi_bike/greedy/launcher.py:35-65 with supply_top_k = demand_top_k = 1): the candidate
 * with the largest (value, station index) — heapq keeps the max tuple when the heap size is 1 */
void bike_policy_greedy(const int32_t* dec, int32_t* action) {
    int station = dec[MARO_BIKE_DEC_STATION], n = dec[MARO_BIKE_DEC_N_SCOPE];
    int best = -1, best_v = 0;
    for (int k = 0; k < n; k++) {
        int idx = dec[MARO_BIKE_DEC_HEAD + 2 * k], v = dec[MARO_BIKE_DEC_HEAD + 2 * k + 1];
        if (idx == station) continue;
        if (best < 0 || v > best_v || (v == best_v && idx > best)) { best = idx; best_v = v; }
    }
    if (best < 0) { action[0] = action[1] = -1; action[2] = action[3] = 0; return; }
    if (dec[MARO_BIKE_DEC_TYPE] == 0) { action[0] = station; action[1] = best; }
    else { action[0] = best; action[1] = station; }
    action[2] = best_v; action[3] = 0;
}
int64_t bike_oracle_run_episode(BikeOracle* o, int policy, int64_t* final_metrics) {
    int32_t dec[MARO_BIKE_DEC_HEAD + 2 * 4096], act[4];
    int64_t met[3] = {0, 0, 0}, steps = 0;
    int st = bike_oracle_step(o, NULL, 0, dec, met);
    while (st == MARO_STATUS_DECISION) {
        steps++;
        if (policy == 1) { bike_policy_greedy(dec, act); st = bike_oracle_step(o, act, 1, dec, met); }
        else st = bike_oracle_step(o, NULL, 0, dec, met);
    }
    steps++;
    if (final_metrics) memcpy(final_metrics, met, sizeof(met));
    return steps;
}
