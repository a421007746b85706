// vm_host.hpp — host-side (plain C++) shape + static-table serialisation for the vm_scheduling scenario, shared by
// vm_env.cu and the test-only host-emulation harness.  No CUDA here.
#pragma once
#include <string>

#include "cim_host.hpp"
#include "vm_core.cuh"

namespace maro {

// _cpu_utilization_to_energy_consumption (business_engine.py:671-688) for cpu_utilization = k / 100
inline double vm_energy_host(const double* power, double ticks_per_hour, double cpu_utilization) {
    cpu_utilization /= 100;
    if (1 < cpu_utilization) cpu_utilization = 1;
    double per_hour = power[2] + (power[1] - power[2]) * (2 * cpu_utilization - pow(cpu_utilization, power[0]));
    return (per_hour / ticks_per_hour) / 1000;
}

// returns "" or the reason the topology cannot be represented
inline std::string vm_compute_shape_and_tables(const MaroVmTopology& t, const MaroCimConfig* cfg, VmShape& s,
                                               std::vector<int32_t>& tables) {
    memset(&s, 0, sizeof(s));
    if (t.n_pm < 1 || t.n_rack < 1 || t.n_cluster < 1 || t.n_dc < 1 || t.n_zone < 1 || t.n_region < 1 || t.n_pm_types < 1)
        return "empty hierarchy";
    if (t.n_vm < 0) return "negative VM count";
    if (t.delay_duration < 1) return "DELAY_DURATION must be >= 1";
    s.N = t.n_pm; s.R = t.n_rack; s.C = t.n_cluster; s.D = t.n_dc; s.Z = t.n_zone; s.RG = t.n_region; s.T = t.n_pm_types;
    s.n_vm = t.n_vm;
    s.max_tick = t.max_tick; s.start_tick = cfg->start_tick;
    s.snap_res = cfg->snapshot_resolution > 0 ? cfg->snapshot_resolution : 1;
    const int durations = s.max_tick - s.start_tick;
    if (durations < 1) return "no ticks to simulate";
    const int total_frames = (durations + s.snap_res - 1) / s.snap_res;
    s.ring_rows = cfg->max_snapshots > 0 ? std::min(cfg->max_snapshots, total_frames) : total_frames;
    s.delay = t.delay_duration; s.budget = t.buffer_budget; s.kill_all = t.kill_all;
    s.max_actions = cfg->max_actions > 0 ? cfg->max_actions : 1;
    s.n_replicas = cfg->n_replicas;
    s.max_cpu_over = t.max_cpu_over; s.max_mem_over = t.max_mem_over; s.max_util_rate = t.max_util_rate;
    s.unit_energy_price = t.unit_energy_price; s.pue = t.pue;
    const int N = s.N;
    s.o_rack = VPA_COUNT * N;
    s.o_cluster = s.o_rack + 7 * s.R;
    s.o_dc = s.o_cluster + 6 * s.C;
    s.o_zone = s.o_dc + 5 * s.D;
    s.o_region = s.o_zone + 4 * s.Z;
    s.FW = s.o_region + 3 * s.RG;
    s.FWp = round_up(s.FW, 4);
    s.CWp = round_up(VC_COUNT, 4);
    s.DW = round_up(MARO_VM_DEC_HEAD + 2 * N, 4);  // header | valid PM ids (<= N) | their remaining CPU cores (<= N)  // rows stay 16-byte aligned; the metrics block behind them 8-byte aligned

    // PM list slots: the most VMs the valid-PM rules let one PM hold (cpu / memory over-subscription bounds with the
    // smallest request of the trace), plus slack for agents that allocate outside the valid list
    int min_cores = 1 << 30, min_mem = 1 << 30;
    for (int v = 0; v < t.n_vm; v++) {
        min_cores = std::min(min_cores, std::max(1, t.vm_attr[8 * v + 6]));
        min_mem = std::min(min_mem, std::max(1, t.vm_attr[8 * v + 7]));
    }
    if (t.n_vm == 0) min_cores = min_mem = 1;
    int K = 1;
    for (int p = 0; p < N; p++) {
        double by_cpu = std::max(1.0, t.max_cpu_over) * t.pm_attr[8 * p + 0] / min_cores;
        double by_mem = std::max(1.0, t.max_mem_over) * t.pm_attr[8 * p + 1] / min_mem;
        K = std::max(K, (int)std::min(by_cpu, by_mem) + 1);
    }
    K = std::min(K + 4, std::max(t.n_vm, 1));
    s.K = std::max(K, 1);

    // postponed-request ring: requests of a (budget + 2 delays) window can be pending at once
    const int window = std::max(0, t.buffer_budget) + 2 * t.delay_duration + 1;
    int fq = 0;
    for (int k = 0; k < t.max_tick; k++) {
        int lo = std::max(0, k - window);
        fq = std::max(fq, t.req_offset[k + 1] - t.req_offset[lo]);
    }
    fq = cfg->queue_capacity > 0 ? cfg->queue_capacity : std::min(fq + 4, std::max(t.n_vm, 1) + 1);
    s.FQ = std::max(fq, 4);
    const long long sw = (long long)s.FWp + s.CWp + round_up(2 * N, 4) + 4LL * s.FQ + 8LL * s.K * N;
    if (sw > 0x7fffffffLL) return "replica block too large";
    s.SW = (int)sw;

    tables.clear();
    BlobBuilder b(tables);
    // initial frame: static attributes, idle energy, every machine empty
    std::vector<int32_t> f0(s.FWp, 0);
    auto rack_total = [&](int i) { return t.rack_range[2 * i + 1] - t.rack_range[2 * i]; };
    for (int p = 0; p < N; p++) {
        const int32_t* a = t.pm_attr + 8 * p;
        float e = (float)t.pm_idle_energy[p];
        int32_t eb;
        memcpy(&eb, &e, 4);
        const int32_t v[VPA_COUNT] = {a[6], 0, a[0], 0, a[5], eb, p, 0, a[1], 0, a[2], a[7], a[3], a[4]};
        for (int k = 0; k < VPA_COUNT; k++) f0[k * N + p] = v[k];
    }
    std::vector<int> rt(s.R), ct(s.C), dt(s.D), zt(s.Z), gt(s.RG);
    for (int i = 0; i < s.R; i++) rt[i] = rack_total(i);
    for (int i = 0; i < s.C; i++) { ct[i] = 0; for (int k = t.cluster_range[2 * i]; k < t.cluster_range[2 * i + 1]; k++) ct[i] += rt[k]; }
    for (int i = 0; i < s.D; i++) { dt[i] = 0; for (int k = t.dc_range[2 * i]; k < t.dc_range[2 * i + 1]; k++) dt[i] += ct[k]; }
    for (int i = 0; i < s.Z; i++) { zt[i] = 0; for (int k = t.zone_range[2 * i]; k < t.zone_range[2 * i + 1]; k++) zt[i] += dt[k]; }
    for (int i = 0; i < s.RG; i++) { gt[i] = 0; for (int k = t.region_range[2 * i]; k < t.region_range[2 * i + 1]; k++) gt[i] += zt[k]; }
    for (int i = 0; i < s.R; i++) {
        const int32_t* ids = t.rack_ids + 4 * i;
        const int32_t v[7] = {ids[3], ids[2], rt[i], i, ids[0], rt[i], ids[1]};
        for (int k = 0; k < 7; k++) f0[s.o_rack + k * s.R + i] = v[k];
    }
    for (int i = 0; i < s.C; i++) {
        const int32_t* ids = t.cluster_ids + 3 * i;
        const int32_t v[6] = {ids[2], ct[i], i, ids[0], ct[i], ids[1]};
        for (int k = 0; k < 6; k++) f0[s.o_cluster + k * s.C + i] = v[k];
    }
    for (int i = 0; i < s.D; i++) {
        const int32_t v[5] = {dt[i], i, t.dc_ids[2 * i], dt[i], t.dc_ids[2 * i + 1]};
        for (int k = 0; k < 5; k++) f0[s.o_dc + k * s.D + i] = v[k];
    }
    for (int i = 0; i < s.Z; i++) {
        const int32_t v[4] = {zt[i], i, t.zone_ids[i], zt[i]};
        for (int k = 0; k < 4; k++) f0[s.o_zone + k * s.Z + i] = v[k];
    }
    for (int i = 0; i < s.RG; i++) {
        const int32_t v[3] = {gt[i], i, gt[i]};
        for (int k = 0; k < 3; k++) f0[s.o_region + k * s.RG + i] = v[k];
    }
    s.t_frame0 = b.put_i(f0.data(), s.FWp, s.FWp);

    std::vector<double> energy((size_t)s.T * (VM_UTIL_STEPS + 1));
    for (int ty = 0; ty < s.T; ty++)
        for (int k = 0; k <= VM_UTIL_STEPS; k++)
            energy[(size_t)ty * (VM_UTIL_STEPS + 1) + k] = vm_energy_host(t.pmtype_power + 3 * ty, t.ticks_per_hour, (double)k / 100.0);
    s.t_energy = b.put_d(energy.data(), (int)energy.size(), (int)energy.size());
    for (int p = 0; p < N; p++)
        if (t.pm_attr[8 * p + 2] < 0 || t.pm_attr[8 * p + 2] >= s.T) return "pm_type out of range";

    s.t_rack_range = b.put_i(t.rack_range, 2 * s.R, 2 * s.R);
    s.t_cluster_range = b.put_i(t.cluster_range, 2 * s.C, 2 * s.C);
    s.t_dc_range = b.put_i(t.dc_range, 2 * s.D, 2 * s.D);
    s.t_zone_range = b.put_i(t.zone_range, 2 * s.Z, 2 * s.Z);
    s.t_region_range = b.put_i(t.region_range, 2 * s.RG, 2 * s.RG);

    const int nv = std::max(t.n_vm, 1);
    std::vector<int32_t> rec0(4 * (size_t)nv, 0), rec1(4 * (size_t)nv, 0);
    for (int v = 0; v < t.n_vm; v++) {
        const int32_t* a = t.vm_attr + 8 * v;
        rec0[4 * v + 0] = t.util_offset[v]; rec0[4 * v + 1] = a[3]; rec0[4 * v + 2] = a[4]; rec0[4 * v + 3] = a[6];
        rec1[4 * v + 0] = a[7]; rec1[4 * v + 1] = a[0]; rec1[4 * v + 2] = a[1]; rec1[4 * v + 3] = a[5];
        if (t.util_offset[v + 1] <= t.util_offset[v]) return "a VM has an empty utilisation series";
    }
    tables.resize(round_up((int)tables.size(), 4), 0);  // 16-byte records
    s.t_rec0 = b.put_i(rec0.data(), (int)rec0.size(), (int)rec0.size());
    s.t_rec1 = b.put_i(rec1.data(), (int)rec1.size(), (int)rec1.size());
    std::vector<double> price(nv, 0.0);
    for (int v = 0; v < t.n_vm; v++) price[v] = t.vm_price[v];
    s.t_price = b.put_d(price.data(), nv, nv);
    s.t_req_offset = b.put_i(t.req_offset, t.max_tick + 1, t.max_tick + 1);
    // utilisation readings: float32 in the trace files (cpu_reader.py), clamped like VirtualMachine.get_utilization
    const int nu = t.n_vm > 0 ? t.util_offset[t.n_vm] : 0;
    std::vector<int32_t> val(std::max(nu, 1), 0);
    std::vector<uint8_t> has(round_up(std::max(nu, 1), 4), 0);
    for (int k = 0; k < nu; k++) {
        double x = t.util_val[k];
        x = x < 0 ? 0 : (x > 100 ? 100 : x);
        float f = (float)x;
        if ((double)f != x) return "utilisation reading is not a float32 value";
        memcpy(&val[k], &f, 4);
        has[k] = t.util_has[k] != 0;
    }
    s.t_val = b.put_i(val.data(), (int)val.size(), (int)val.size());
    s.t_has = b.put_i(reinterpret_cast<const int32_t*>(has.data()), (int)has.size() / 4, (int)has.size() / 4);
    tables.resize(round_up((int)tables.size(), 4), 0);
    return "";
}

}  // namespace maro
