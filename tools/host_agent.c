#define _GNU_SOURCE
/* Bench infrastructure: the hello-world random agent (examples/hello_world/cim/hello.py:24-32) on the host, as the
 * same counter hash of (replica, decision ordinal) that cim_policy_kernel evaluates on the device.  This is the
 * *agent* of the end-to-end measurement (user code outside the library), compiled by bench.py with gcc. */
#include <stdint.h>
#include <unistd.h>

static inline uint32_t hash_u32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

void agent_random(const int32_t* dec, int32_t* act, int n, int max_actions, uint32_t seed, uint32_t replica_base) {
#pragma omp parallel for schedule(static) num_threads(n >= 4096 ? 16 : 4) if (n >= 1024)
    for (int i = 0; i < n; i++) {
        const int32_t* d = dec + 8 * i;
        uint32_t h1 = hash_u32(seed ^ hash_u32((uint32_t)(i + replica_base) * 0x9e3779b9u + (uint32_t)d[7] * 0x85ebca6bu + 0x1234567u));
        uint32_t h2 = hash_u32(h1 + 0x68bc21ebu);
        int load = d[3], dis = d[4];
        int to_discharge = dis > 0 && (h1 & 1u);
        int scope = to_discharge ? dis : load;
        int32_t* a = act + (int64_t)i * max_actions * 4;
        a[0] = d[2]; a[1] = d[1];
        a[2] = scope > 0 ? (int32_t)(h2 % (uint32_t)(scope + 1)) : 0;
        a[3] = to_discharge;
    }
}

/* greedy top-1 citi_bike agent (examples/citi_bike/greedy/launcher.py:35-65): the candidate with the largest
 * (value, station) pair; decision rows are 8 + 2 * S words */
void agent_greedy(const int32_t* dec, int32_t* act, int n, int max_actions, int dec_words) {
#pragma omp parallel for schedule(static) num_threads(16) if (n >= 4096)
    for (int i = 0; i < n; i++) {
        const int32_t* d = dec + (int64_t)i * dec_words;
        int station = d[1], ns = d[4], best = -1, best_v = 0;
        for (int k = 0; k < ns; k++) {
            int idx = d[8 + 2 * k], v = d[9 + 2 * k];
            if (idx == station) continue;
            if (best < 0 || v > best_v || (v == best_v && idx > best)) { best = idx; best_v = v; }
        }
        int32_t* a = act + (int64_t)i * max_actions * 4;
        if (best < 0) { a[0] = a[1] = -1; a[2] = a[3] = 0; }
        else if (d[3] == 0) { a[0] = station; a[1] = best; a[2] = best_v; a[3] = 0; }
        else { a[0] = best; a[1] = station; a[2] = best_v; a[3] = 0; }
    }
}

/* best-fit vm_scheduling agent (examples/vm_scheduling/rule_based_algorithm/best_fit.py:27-64, "remaining_cpu_cores"):
 * `q` is the snapshot query the reference agent makes, as returned by maro_vm_query for every replica:
 * q[replica][frame][pm][{cpu_cores_capacity, cpu_cores_allocated}] (float64), `frames` the queried frame indices.
 * Decision rows: 12-word header (frame index at [2], status at [6], n_valid at [10]) + valid PM ids. */
void agent_best_fit(const int32_t* dec, int32_t* act, int n, int max_actions, int dec_words, const double* q,
                    const int32_t* frames, int n_frames, int n_pm) {
#pragma omp parallel for schedule(static) num_threads(16) if (n >= 1024)
    for (int i = 0; i < n; i++) {
        const int32_t* d = dec + (int64_t)i * dec_words;
        int32_t* a = act + (int64_t)i * max_actions * 4;
        int nv = d[6] == 0 ? d[10] : 0, f = 0;
        while (f < n_frames - 1 && frames[f] != d[2]) f++;
        const double* qq = q + ((int64_t)i * n_frames + f) * n_pm * 2;
        int best = -1;
        double best_rem = 0;
        for (int k = 0; k < nv; k++) {
            int p = d[12 + k];
            double rem = qq[2 * p] - qq[2 * p + 1];
            if (best < 0 || rem < best_rem) { best = p; best_rem = rem; }
        }
        if (best < 0) { a[0] = a[1] = -1; a[2] = a[3] = 0; }
        else { a[0] = d[1]; a[1] = 0; a[2] = best; a[3] = 0; }
    }
}

/* the same agent reading the decision row's extension area (remaining CPU cores per valid PM, include/maro_b200.h
 * MARO_VM_DEC_EXT_OFFSET) instead of a snapshot query: same choice, no 3 MB query per step */
void agent_best_fit_row(const int32_t* dec, int32_t* act, int n, int max_actions, int dec_words) {
#pragma omp parallel for schedule(static) num_threads(16) if (n >= 1024)
    for (int i = 0; i < n; i++) {
        const int32_t* d = dec + (int64_t)i * dec_words;
        int32_t* a = act + (int64_t)i * max_actions * 4;
        const int nv = d[6] == 0 ? d[10] : 0, ext = d[11];
        int best = -1, best_rem = 0;
        for (int k = 0; k < nv; k++) {
            const int rem = d[ext + k];
            if (best < 0 || rem < best_rem) { best = d[12 + k]; best_rem = rem; }
        }
        if (best < 0) { a[0] = a[1] = -1; a[2] = a[3] = 0; }
        else { a[0] = d[1]; a[1] = 0; a[2] = best; a[3] = 0; }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * End-to-end host loop of the CIM bench (user code on top of the C ABI, include/maro_b200.h): the batch is cut into
 * `n_sub` contiguous sub-batches; for each one the loop waits for its decision rows, runs the agent on them and submits
 * the actions (maro_cim_wait_pinned / maro_cim_submit_pinned), so the agent's work on one sub-batch overlaps the
 * device's work and the PCIe latency of the others.  n_sub = 1 is the plain lock-step loop.  Episodes restart
 * (maro_cim_reset with a mask) when every replica of a sub-batch reports DONE.
 * `dec` / `act` / `mask` are the library's pinned buffers (maro_cim_pinned_buffers). */
#include <time.h>

typedef int (*submit_fn)(void*, int32_t, int32_t, int32_t, int32_t, int32_t);
typedef int (*wait_fn)(void*, int32_t, int32_t);
typedef int (*reset_fn)(void*, const uint8_t*);

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void agent_random_range(const int32_t* dec, int32_t* act, int first, int count, uint32_t seed, uint32_t replica_base) {
    for (int i = first; i < first + count; i++) {
        const int32_t* d = dec + 8 * i;
        uint32_t h1 = hash_u32(seed ^ hash_u32((uint32_t)(i + replica_base) * 0x9e3779b9u + (uint32_t)d[7] * 0x85ebca6bu + 0x1234567u));
        uint32_t h2 = hash_u32(h1 + 0x68bc21ebu);
        int load = d[3], dis = d[4];
        int to_discharge = dis > 0 && (h1 & 1u);
        int scope = to_discharge ? dis : load;
        int32_t* a = act + (int64_t)i * 4;
        a[0] = d[2]; a[1] = d[1];
        a[2] = scope > 0 ? (int32_t)(h2 % (uint32_t)(scope + 1)) : 0;
        a[3] = to_discharge;
    }
}

/* returns 0 on success; out[0] = seconds of the timed loop, out[1] = seconds inside the agent, out[2] = resets */
int e2e_loop_cim(void* env, void* submit_p, void* wait_p, void* reset_p, const int32_t* dec, int32_t* act, uint8_t* mask,
                 int B, int gran, int n_sub, int n_steps, uint32_t seed, uint32_t replica_base, double* out) {
    submit_fn submit = (submit_fn)submit_p;
    wait_fn wait = (wait_fn)wait_p;
    reset_fn reset = (reset_fn)reset_p;
    int first[65], n = 0;
    if (n_sub < 1) n_sub = 1;
    if (n_sub > 64) n_sub = 64;
    int blocks = (B + gran - 1) / gran, per = (blocks + n_sub - 1) / n_sub;
    for (int b = 0; b < blocks; b += per) first[n++] = b * gran;
    first[n] = B;
    double t_agent = 0.0, resets = 0.0;
    const double t0 = now_s();
    for (int k = 0; k < n; k++)  /* generator start: the first step of an episode takes no action */
        if (submit(env, first[k], first[k + 1] - first[k], 0, 0, 0)) return 1;
    for (int step = 0; step < n_steps; step++) {
        for (int k = 0; k < n; k++) {
            const int f = first[k], c = first[k + 1] - first[k];
            if (wait(env, f, c)) return 1;
            if (step == n_steps - 1) continue;
            int done = 1;
            for (int i = f; i < f + c; i++) done &= dec[8 * i + 6] != 0;
            if (done) {  /* Env.reset for this sub-batch, then the episode's first step */
                for (int i = 0; i < B; i++) mask[i] = (uint8_t)(i >= f && i < f + c);
                if (reset(env, mask)) return 1;
                resets += 1.0;
                if (submit(env, f, c, 0, 0, 0)) return 1;
                continue;
            }
            const double ta = now_s();
            agent_random_range(dec, act, f, c, seed, replica_base);
            t_agent += now_s() - ta;
            if (submit(env, f, c, 1, 0, 0)) return 1;
        }
    }
    out[0] = now_s() - t0;
    out[1] = t_agent;
    out[2] = resets;
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * The same loop on `n_threads` host threads: the sub-batches are dealt out in contiguous runs, every thread drives its
 * own (wait -> agent -> submit) pipeline.  maro_cim_submit_pinned / maro_cim_wait_pinned / maro_cim_reset accept concurrent
 * callers on disjoint replica ranges (include/maro_b200.h).  out[1] = agent seconds averaged over the threads. */
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    void* env; submit_fn submit; wait_fn wait; reset_fn reset;
    const int32_t* dec; int32_t* act; int B, n_steps, k0, k1; const int* first; uint32_t seed, replica_base;
    double t_agent, resets; int rc, cpu;
} e2e_worker;

static void* e2e_worker_main(void* p) {
    e2e_worker* w = (e2e_worker*)p;
    if (w->cpu >= 0) {  /* next to the thread that created the handle (same socket as its pinned buffers) */
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(w->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    uint8_t* mask = (uint8_t*)malloc((size_t)w->B);
    w->rc = 0; w->t_agent = 0.0; w->resets = 0.0;
    for (int k = w->k0; k < w->k1 && !w->rc; k++)
        w->rc = w->submit(w->env, w->first[k], w->first[k + 1] - w->first[k], 0, 0, 0);
    for (int step = 0; step < w->n_steps && !w->rc; step++) {
        for (int k = w->k0; k < w->k1 && !w->rc; k++) {
            const int f = w->first[k], c = w->first[k + 1] - w->first[k];
            if ((w->rc = w->wait(w->env, f, c))) break;
            if (step == w->n_steps - 1) continue;
            int done = 1;
            for (int i = f; i < f + c; i++) done &= w->dec[8 * i + 6] != 0;
            if (done) {
                memset(mask, 0, (size_t)w->B);
                memset(mask + f, 1, (size_t)c);
                if ((w->rc = w->reset(w->env, mask))) break;
                w->resets += 1.0;
                w->rc = w->submit(w->env, f, c, 0, 0, 0);
                continue;
            }
            const double ta = now_s();
            agent_random_range(w->dec, w->act, f, c, w->seed, w->replica_base);
            w->t_agent += now_s() - ta;
            w->rc = w->submit(w->env, f, c, 1, 0, 0);
        }
    }
    free(mask);
    return 0;
}

int e2e_loop_cim_mt(void* env, void* submit_p, void* wait_p, void* reset_p, const int32_t* dec, int32_t* act, int B, int gran,
                    int n_sub, int n_threads, int n_steps, uint32_t seed, uint32_t replica_base, double* out) {
    int first[65], n = 0;
    if (n_sub < 1) n_sub = 1;
    if (n_sub > 64) n_sub = 64;
    int blocks = (B + gran - 1) / gran, per = (blocks + n_sub - 1) / n_sub;
    for (int b = 0; b < blocks; b += per) first[n++] = b * gran;
    first[n] = B;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = n;
    e2e_worker w[64];
    pthread_t th[64];
    const double t0 = now_s();
    for (int t = 0; t < n_threads; t++) {
        e2e_worker* x = &w[t];
        x->env = env; x->submit = (submit_fn)submit_p; x->wait = (wait_fn)wait_p; x->reset = (reset_fn)reset_p;
        x->dec = dec; x->act = act; x->B = B; x->n_steps = n_steps; x->first = first; x->seed = seed; x->replica_base = replica_base;
        x->k0 = (int)((long long)n * t / n_threads); x->k1 = (int)((long long)n * (t + 1) / n_threads);
        x->cpu = -1;
        if (t > 0) {
            const int cpu0 = sched_getcpu(), ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN);
            if (cpu0 >= 0 && ncpu > 0) x->cpu = (cpu0 + t) % ncpu;
            pthread_create(&th[t], 0, e2e_worker_main, x);
        }
    }
    e2e_worker_main(&w[0]);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], 0);
    out[0] = now_s() - t0; out[1] = 0.0; out[2] = 0.0;
    int rc = 0;
    for (int t = 0; t < n_threads; t++) { out[1] += w[t].t_agent / n_threads; out[2] += w[t].resets; rc |= w[t].rc; }
    return rc;
}
