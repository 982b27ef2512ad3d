// Host side of the traveller exchange: the reference's sequential bookkeeping (proportional allotment of migrators,
// occupancy heaps, agent slots) around the device kernels of travel.cu.  C ABI: epi_travel_pack / epi_travel_unpack /
// epi_finish_hour (include/epi.h).
#include <algorithm>
#include <cstring>

#include "engine.h"
#include "kernels.h"
#include "philox.cuh"

using namespace epi;

#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t _err = (call);                                                                                 \
        if (_err != cudaSuccess)                                                                                   \
            return engine_fail(e, EPI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_err));              \
    } while (0)

namespace {

enum { KK_TRAVEL = 6 };

template <class T>
int grow(epi_engine* e, T** p, size_t count) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    CU(cudaMalloc((void**)p, count * sizeof(T)));
    e->device_bytes += count * sizeof(T);
    return EPI_OK;
}

// lists of up to `n` travellers
int ensure_lists(epi_engine* e, size_t n) {
    if (n <= e->t_list_capacity) return EPI_OK;
    const size_t cap = std::max<size_t>(n + n / 2, 4096);
    int rc;
    if ((rc = grow(e, &e->t_out_slots, cap))) return rc;
    if ((rc = grow(e, &e->t_out_dest, cap))) return rc;
    if ((rc = grow(e, &e->t_idx, 3 * cap))) return rc;  // send_slots | in_slot, in_home, in_work
    if ((rc = grow(e, &e->t_placed, cap))) return rc;
    size_t table = 1024;
    while (table < 4 * cap) table <<= 1;
    if ((rc = grow(e, &e->t_table_keys, table))) return rc;
    if ((rc = grow(e, &e->t_table_vals, table))) return rc;
    e->t_table_capacity = table;
    e->t_list_capacity = cap;
    return EPI_OK;
}

int ensure_scan(epi_engine* e) {
    if (e->t_block_counts) return EPI_OK;
    int rc;
    if ((rc = grow(e, &e->t_block_counts, (size_t)(e->P.n + 255) / 256 + 1))) return rc;
    return grow(e, &e->t_total, 4);
}

void note_launch(epi_engine* e, unsigned n = 1) {
    e->launches += n;
    e->kernel_launches[KK_TRAVEL] += n;
}

}  // namespace

extern "C" {

int epi_travel_pack(epi_engine* e, uint32_t hour, int kind, void* send_buf, uint64_t capacity_records, uint32_t* counts_out) {
    if (!e || !counts_out) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (!e->multi) return engine_fail(e, EPI_ERR_STATE, "not a multi-region engine (epi_create_multi)");
    const uint32_t R = (uint32_t)e->n_regions, h = hour % 24u;
    std::fill(counts_out, counts_out + R, 0u);
    TravelArgs A{};
    A.kind = kind;
    A.hour = hour;
    A.hour_of_day = h;
    if (kind == TRAVEL_MIGRATE) {
        // Citizen::can_migrate's hour window (citizen/mod.rs:460-462); EngineMigrationPlan::percent_outgoing (:44-49)
        if (!e->migration_enabled || h != 0 || !(hour > e->start_migration_hour && hour < e->end_migration_hour)) return EPI_OK;
        uint64_t planned_total = 0;
        for (uint32_t v : e->migration_row) planned_total += v;
        if (planned_total == 0 || e->population == 0) return EPI_OK;
        A.thr_outgoing = bernoulli_threshold((double)planned_total / (double)e->population);
    } else if (kind == TRAVEL_COMMUTE) {
        if (!e->commute_enabled || !(h == 7 || h == 17)) return EPI_OK;
    } else {
        return engine_fail(e, EPI_ERR_ARG, "epi_travel_pack: kind must be EPI_TRAVEL_MIGRATE or EPI_TRAVEL_COMMUTE");
    }
    CU(cudaSetDevice(e->device));
    int rc = ensure_scan(e);
    if (rc) return rc;
    // ordered compaction of the leaving agents (ascending slot = the phase-B order of this implementation)
    launch_travel_select(e->P, e->D, A, e->t_block_counts, e->t_total, nullptr, nullptr, 0, e->stream);
    note_launch(e, 2);
    CU(cudaMemcpyAsync(e->h_small, e->t_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    const uint32_t total = e->h_small[0];
    if (total == 0) return EPI_OK;
    if ((rc = ensure_lists(e, total))) return rc;
    launch_travel_select(e->P, e->D, A, e->t_block_counts, e->t_total, e->t_out_slots, e->t_out_dest, 1, e->stream);
    note_launch(e);
    std::vector<uint32_t> slots(total), dest(total);
    CU(cudaMemcpyAsync(slots.data(), e->t_out_slots, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(dest.data(), e->t_out_dest, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    std::vector<uint32_t> send;
    send.reserve(total);
    if (kind == TRAVEL_MIGRATE) {
        // alloc_outgoing_to_regions (engine_migration_plan.rs:51-77) + MigratorsByRegion::alloc_citizens
        // (migrators_by_engine.rs:27-55): regions in plan order take floor(share * total) from the front; the rest stays
        uint64_t planned_total = 0;
        for (uint32_t v : e->migration_row) planned_total += v;
        uint32_t front = 0;
        for (uint32_t to = 0; to < R; ++to) {
            if ((int)to == e->P.region || e->migration_row[to] == 0) continue;
            const double share = (double)e->migration_row[to] / (double)planned_total;
            uint32_t count = (uint32_t)(int32_t)(share * (double)(int32_t)total);
            if (count > total - front) count = total - front;
            counts_out[to] = count;
            for (uint32_t k = 0; k < count; ++k) send.push_back(slots[front + k]);
            front += count;
        }
    } else {
        // CommutersByRegion::get_commuters_by_region (commuters_by_region.rs:59-78): per region in plan order
        for (uint32_t k = 0; k < total; ++k)
            if (dest[k] >= R) return engine_fail(e, EPI_ERR_STATE, "commuter with a work / home region outside the travel plan");
        for (uint32_t to = 0; to < R; ++to)
            for (uint32_t k = 0; k < total; ++k)
                if (dest[k] == to) { send.push_back(slots[k]); counts_out[to]++; }
    }
    const uint32_t n_send = (uint32_t)send.size();
    if (n_send == 0) return EPI_OK;
    if (!send_buf || n_send > capacity_records) return engine_fail(e, EPI_ERR_ARG, "epi_travel_pack: send buffer too small for " + std::to_string(n_send) + " records");
    CU(cudaMemcpyAsync(e->t_idx, send.data(), n_send * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    launch_travel_pack(e->P, e->D, e->t_idx, n_send, (TravelRecord*)send_buf, e->stream);
    note_launch(e);
    if (kind == TRAVEL_MIGRATE) {
        // remove_migrators (allocation_map.rs:165-192): the leavers' houses / offices lose an occupant
        std::vector<TravelRecord> recs(n_send);
        CU(cudaMemcpyAsync(recs.data(), send_buf, (size_t)n_send * sizeof(TravelRecord), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        for (const TravelRecord& r : recs) {
            if (!e->houses_occupancy.remove_occupant(house_index_of(e->geo, r.home))) return engine_fail(e, EPI_ERR_STATE, "Could not find house");
            if (((r.st >> ST_WS_SHIFT) & 3u) != WS_NA && !e->offices_occupancy.remove_occupant(office_index_of(e->geo, r.work)))
                return engine_fail(e, EPI_ERR_STATE, "Could not find office");
        }
        for (uint32_t sl : send) e->free_slots.push_back(sl);
    } else {
        CU(cudaStreamSynchronize(e->stream));  // `send` is read by the H2D copy
        for (uint32_t k = 0; k < total; ++k) e->free_slots.push_back(slots[k]);  // remove_commuters walks the list in selection order
    }
    e->population -= n_send;
    e->have_last_row = false;
    CU(cudaGetLastError());
    return EPI_OK;
}

int epi_travel_unpack(epi_engine* e, uint32_t hour, int kind, const void* recv_buf, const uint32_t* counts_in) {
    if (!e || !counts_in) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (!e->multi) return engine_fail(e, EPI_ERR_STATE, "not a multi-region engine (epi_create_multi)");
    if (kind != TRAVEL_MIGRATE && kind != TRAVEL_COMMUTE) return engine_fail(e, EPI_ERR_ARG, "epi_travel_unpack: bad kind");
    uint64_t n64 = 0;
    for (int r = 0; r < e->n_regions; ++r) n64 += counts_in[r];
    if (n64 == 0) return EPI_OK;
    if (!recv_buf) return engine_fail(e, EPI_ERR_ARG, "null receive buffer");
    const uint32_t n_in = (uint32_t)n64;
    if (e->free_slots.size() < n_in) return engine_fail(e, EPI_ERR_STATE, "region is out of agent slots: raise extra_capacity (" + std::to_string(n_in) + " arrivals)");
    CU(cudaSetDevice(e->device));
    int rc = ensure_lists(e, n_in);
    if (rc) return rc;
    TravelArgs A{};
    A.kind = kind;
    A.hour = hour;
    A.hour_of_day = hour % 24u;
    std::vector<uint32_t> idx(3 * (size_t)n_in, 0u);  // in_slot | in_home | in_work
    uint32_t* in_slot = idx.data();
    uint32_t* in_home = idx.data() + n_in;
    uint32_t* in_work = idx.data() + 2 * (size_t)n_in;
    auto take_slot = [&]() { const uint32_t s = e->free_slots.back(); e->free_slots.pop_back(); return s; };
    if (kind == TRAVEL_MIGRATE) {
        // assimilate_migrators (allocation_map.rs:214-243): choose_house_with_free_space / choose_office_with_free_space
        std::vector<TravelRecord> recs(n_in);
        CU(cudaMemcpyAsync(recs.data(), recv_buf, (size_t)n_in * sizeof(TravelRecord), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        for (uint32_t k = 0; k < n_in; ++k) {
            const bool working = ((recs[k].st >> ST_WS_SHIFT) & 3u) != WS_NA;
            if (e->houses_occupancy.empty()) return engine_fail(e, EPI_ERR_STATE, "Couldn't find any house with free space!");
            const uint32_t house = e->houses_occupancy.pop_min();
            if (e->houses_occupancy.occupants(house) >= 4) return engine_fail(e, EPI_ERR_STATE, "Couldn't find any house with free space!");
            uint32_t office = 0;
            if (working) {
                office = e->offices_occupancy.pop_min();
                if (e->offices_occupancy.occupants(office) >= 100) return engine_fail(e, EPI_ERR_STATE, "Couldn't find any offices with free space!");
            }
            in_slot[k] = take_slot();
            in_home[k] = house_origin(e->geo, house);
            in_work[k] = 0;  // WorkStatus::NA after from_migrator: the office only counts in the occupancy heap
            e->houses_occupancy.add_occupant(house);
            if (working) e->offices_occupancy.add_occupant(office);
        }
    } else {
        // assimilate_commuters (allocation_map.rs:245-277): an office is assigned at the absolute hour 7 only (:260)
        for (uint32_t k = 0; k < n_in; ++k) {
            in_slot[k] = take_slot();
            if (hour == 7u) {
                const uint32_t office = e->offices_occupancy.pop_min();
                if (e->offices_occupancy.occupants(office) >= 100) return engine_fail(e, EPI_ERR_STATE, "Couldn't find any offices with free space!");
                e->offices_occupancy.add_occupant(office);
                in_work[k] = office_origin(e->geo, office);
            }
        }
    }
    CU(cudaMemcpyAsync(e->t_idx, idx.data(), idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    launch_travel_install(e->P, e->D, A, (const TravelRecord*)recv_buf, n_in, e->t_idx, e->t_idx + n_in, e->t_idx + 2 * (size_t)n_in, e->stream);
    note_launch(e);
    // select_starting_points: placement rounds until every arrival holds a distinct vacant cell
    CU(cudaMemsetAsync(e->t_placed, 0, n_in, e->stream));
    size_t table = 1024;
    while (table < 4 * (size_t)n_in) table <<= 1;
    for (uint32_t attempt = 0;; ++attempt) {
        if (attempt > 64) return engine_fail(e, EPI_ERR_STATE, "Not enough locations are available for travellers");
        launch_travel_round(e->P, e->D, A, n_in, attempt, e->t_placed, e->t_idx, e->t_table_keys, e->t_table_vals, (uint32_t)table - 1u, e->t_total, e->stream);
        note_launch(e, 2);
        CU(cudaMemcpyAsync(e->h_small, e->t_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        if (e->h_small[0] == 0) break;
    }
    e->population += n_in;
    e->have_last_row = false;
    CU(cudaGetLastError());
    return EPI_OK;
}

int epi_finish_hour(epi_engine* e, uint32_t hour, epi_counts* out) {
    if (!e || !out) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    // Counts after remove_* / assimilate_* adjusted them: the running totals on the device
    std::vector<uint32_t> tot((size_t)TOT_COPIES * 8);
    CU(cudaMemcpyAsync(tot.data(), e->D.tot, tot.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    uint32_t c[6] = {0, 0, 0, 0, 0, 0};
    for (uint32_t k = 0; k < TOT_COPIES; ++k)
        for (int j = 0; j < 6; ++j) c[j] += tot[(size_t)k * 8 + j];
    epi_counts row{hour, c[0], c[1], c[2], c[3], c[4], c[5]};
    const uint64_t total = (uint64_t)c[0] + c[1] + c[2] + c[3] + c[4] + c[5];
    if (total != e->population)
        return engine_fail(e, EPI_ERR_STATE, "counts total " + std::to_string(total) + " != population " + std::to_string(e->population) + " after the exchange of hour " + std::to_string(hour));
    e->last_counts = row;
    e->have_last_row = true;
    *out = row;
    return process_interventions(e, row, false);
}

}  // extern "C"
