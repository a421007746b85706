// TEST INFRASTRUCTURE ONLY — a tiny SIMT emulator: one host thread per lane of a lane-group, collectives implemented
// with a spin barrier + exchange array.  Lets the device source (maro_b200/csrc/cim_core.cuh) run unmodified on the
// CPU, including its shuffle / ballot / match / atomic based cooperative phases.
#pragma once
#include <stdint.h>

#include <atomic>
#include <functional>
#include <thread>
#include <vector>

namespace wemu {

struct Group {
    int width = 1;
    std::atomic<int> arrived{0};
    std::atomic<int> generation{0};
    uint64_t xchg[32];
};

struct LaneCtx {
    Group* g = nullptr;
    int lane = 0;
};

inline thread_local LaneCtx tl;

inline void barrier() {
    Group* g = tl.g;
    if (g->width == 1) return;
    int gen = g->generation.load(std::memory_order_acquire);
    if (g->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == g->width) {
        g->arrived.store(0, std::memory_order_relaxed);
        g->generation.fetch_add(1, std::memory_order_acq_rel);
    } else {
        int spins = 0;
        while (g->generation.load(std::memory_order_acquire) == gen) {
            if (++spins > 64) std::this_thread::yield();
        }
    }
}

// every lane contributes `v`; returns a snapshot of all lanes' values
inline void exchange(uint64_t v, uint64_t* out) {
    Group* g = tl.g;
    g->xchg[tl.lane] = v;
    barrier();
    for (int i = 0; i < g->width; i++) out[i] = g->xchg[i];
    barrier();
}

// run fn(lane) on `width` threads as one lane group
inline void run_group(int width, const std::function<void(int)>& fn) {
    Group g;
    g.width = width;
    if (width == 1) {
        tl.g = &g;
        tl.lane = 0;
        fn(0);
        return;
    }
    std::vector<std::thread> th;
    for (int l = 0; l < width; l++)
        th.emplace_back([&g, l, &fn]() {
            tl.g = &g;
            tl.lane = l;
            fn(l);
        });
    for (auto& t : th) t.join();
}

}  // namespace wemu
