// Host side of the traveller exchange: launches the device kernels of travel.cu and reads back the per-region record counts.
// The reference's sequential bookkeeping (allotment of migrators, occupancy heaps, agent slots) runs on the device.
// C ABI: epi_travel_pack / epi_travel_unpack / epi_finish_hour (include/epi.h).
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

void note_launch(epi_engine* e, unsigned n = 1) {
    e->launches += n;
    e->kernel_launches[KK_TRAVEL] += n;
}

// CUDA events around a group of travel kernels when per-kernel timing is on (epi_set_kernel_timing)
struct TimedGroup {
    epi_engine* e;
    cudaEvent_t a = nullptr, b = nullptr;
    explicit TimedGroup(epi_engine* e_) : e(e_) {
        if (e->timing) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, e->stream); }
    }
    ~TimedGroup() {
        if (e->timing) { cudaEventRecord(b, e->stream); e->pending_events.push_back({KK_TRAVEL, {a, b}}); }
    }
};

std::string travel_error(uint32_t err) {
    std::string m;
    auto add = [&](uint32_t bit, const char* text) { if (err & bit) { if (!m.empty()) m += "; "; m += text; } };
    add(TERR_NO_SLOTS, "region is out of agent slots: raise extra_capacity");
    add(TERR_LIST_OVERFLOW, "more travellers than the exchange lists hold (travel plan underestimates the traffic)");
    add(TERR_SEGMENT_OVERFLOW, "a destination's records do not fit its segment of the send buffer: raise stride_records");
    add(TERR_NO_HOUSE, "Could not find house");
    add(TERR_NO_OFFICE, "Could not find office");
    add(TERR_HOUSES_FULL, "Couldn't find any house with free space!");
    add(TERR_OFFICES_FULL, "Couldn't find any offices with free space!");
    add(TERR_BAD_REGION, "commuter with a work / home region outside the travel plan");
    add(TERR_PERCENT, "migration plan row exceeds the region's population: percent_outgoing > 1");
    add(TERR_NO_PLACE, "Not enough locations are available for travellers");
    add(TERR_PEER_TIMEOUT, "a peer region never delivered its travellers (its process died?)");
    return m;
}

// zero headers: nobody travels from this region in this exchange
int write_empty_headers(epi_engine* e, void* send_buf, uint64_t stride) {
    if (!send_buf) return EPI_OK;
    CU(cudaMemset2DAsync(send_buf, (size_t)stride * sizeof(TravelRecord), 0, sizeof(TravelRecord), (size_t)e->n_regions, e->stream));
    return EPI_OK;
}

}  // namespace

namespace epi {
// The exchange keeps its books on the device (TravelVars: counts, population, sticky error flags).  This waits for the stream and
// brings them to the host: the population mirror, and an error if any exchange of the run failed.
int sync_travel(epi_engine* e) {
    if (!e->multi) return EPI_OK;
    CU(cudaMemcpyAsync(e->h_tv, e->T.tv, sizeof(TravelVars), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaGetLastError());
    e->pack_unsettled = e->unpack_unsettled = false;
    return travel_status(e, e->h_tv->err, e->h_tv->population);
}
int travel_status(epi_engine* e, uint32_t err, uint32_t population) {
    if (err) {
        // sticky on the device: clear it so that the caller can go on after a recoverable error (a pack that overflowed removed nobody)
        cudaMemsetAsync(&e->T.tv->err, 0, sizeof(uint32_t), e->stream);
        cudaMemsetAsync(&e->T.tv->abort, 0, sizeof(uint32_t), e->stream);
        return engine_fail(e, EPI_ERR_STATE, "traveller exchange: " + travel_error(err));
    }
    e->population = population;
    return EPI_OK;
}
}  // namespace epi

extern "C" {

int epi_travel_pack(epi_engine* e, uint32_t hour, int kind, void* send_buf, uint64_t stride_records, uint32_t* counts_out) {
    return epi_travel_pack_impl(e, hour, kind, send_buf, stride_records, counts_out, nullptr, nullptr, 0);
}

}  // extern "C"

// peer_recv / peer_flags: the peer transport's destinations (multi.cpp); the leave kernel pushes the packed segments there itself
int epi_travel_pack_impl(epi_engine* e, uint32_t hour, int kind, void* send_buf, uint64_t stride_records, uint32_t* counts_out, epi::TravelRecord* const* peer_recv,
                         uint32_t* const* peer_flags, uint32_t exchange_no) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (!e->multi) return engine_fail(e, EPI_ERR_STATE, "not a multi-region engine (epi_create_multi)");
    if (kind != TRAVEL_MIGRATE && kind != TRAVEL_COMMUTE) return engine_fail(e, EPI_ERR_ARG, "epi_travel_pack: kind must be EPI_TRAVEL_MIGRATE or EPI_TRAVEL_COMMUTE");
    if (!send_buf || stride_records < 2 || stride_records > 0xFFFFFFFFull) return engine_fail(e, EPI_ERR_ARG, "epi_travel_pack: send buffer / stride_records");
    const uint32_t R = (uint32_t)e->n_regions, h = hour % 24u;
    if (counts_out) std::fill(counts_out, counts_out + R, 0u);
    CU(cudaSetDevice(e->device));
    TravelArgs A{};
    A.kind = kind;
    A.hour = hour;
    A.hour_of_day = h;
    A.row_index = 0xFFFFFFFFu;
    if (kind == TRAVEL_MIGRATE) {
        // Citizen::can_migrate's hour window (citizen/mod.rs:460-462).  EngineMigrationPlan::percent_outgoing (:44-49) is taken
        // on the device from the region's current population (k_travel_select); a row that exceeds it is an error there.
        uint64_t planned_total = 0;
        for (uint32_t v : e->migration_row) planned_total += v;
        if (!e->migration_enabled || h != 0 || !(hour > e->start_migration_hour && hour < e->end_migration_hour) || planned_total == 0) {
            if (peer_recv) return engine_fail(e, EPI_ERR_STATE, "epi_exchange at an hour without travellers");  // the hour loop never asks for it (exchange_kind_of)
            return write_empty_headers(e, send_buf, stride_records);
        }
    } else if (!e->commute_enabled || !(h == 7 || h == 17)) {
        if (peer_recv) return engine_fail(e, EPI_ERR_STATE, "epi_exchange at an hour without travellers");
        return write_empty_headers(e, send_buf, stride_records);
    }
    // every slot is visited (and drawn for) by a migration's leave kernel only: that one gets the wide grid, the rest wait less at their barriers
    if (!e->travel_blocks) { e->travel_blocks = travel_grid_blocks(e->device, 4); e->travel_blocks_small = travel_grid_blocks(e->device, 1); }
    {
        TimedGroup t(e);
        const cudaError_t cerr = launch_travel_leave(e->P, e->D, A, e->T, e->t_block_counts, (TravelRecord*)send_buf, (uint32_t)stride_records,
                                                     e->travel_blocks, peer_recv, peer_flags, exchange_no, e->stream);
        if (cerr != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("k_travel_leave (cooperative launch): ") + cudaGetErrorString(cerr));
        note_launch(e);
    }
    CU(cudaGetLastError());
    e->have_last_row = false;
    e->pack_unsettled = true;
    if (!counts_out) return EPI_OK;  // deferred: the records are in flight on the stream; errors surface when rows are collected
    const int rc = sync_travel(e);
    if (rc) return rc;
    for (uint32_t to = 0; to < R; ++to) counts_out[to] = e->h_tv->cnt[to];  // still the leavers' counts: no unpack ran in between
    return EPI_OK;
}

extern "C" {

int epi_travel_unpack(epi_engine* e, uint32_t hour, int kind, const void* recv_buf, uint64_t stride_records, uint32_t* counts_in) {
    return epi_travel_unpack_impl(e, hour, kind, recv_buf, stride_records, counts_in, nullptr, 0);
}

}  // extern "C"

// wait_flags: the peers' "segment complete" flags of the peer transport (multi.cpp), waited for on the device
int epi_travel_unpack_wait(epi_engine* e, uint32_t hour, int kind, const void* recv_buf, uint64_t stride_records, const uint32_t* wait_flags, uint32_t exchange_no) {
    return epi_travel_unpack_impl(e, hour, kind, recv_buf, stride_records, nullptr, wait_flags, exchange_no);
}

int epi_travel_unpack_impl(epi_engine* e, uint32_t hour, int kind, const void* recv_buf, uint64_t stride_records, uint32_t* counts_in, const uint32_t* wait_flags,
                           uint32_t exchange_no) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (!e->multi) return engine_fail(e, EPI_ERR_STATE, "not a multi-region engine (epi_create_multi)");
    if (kind != TRAVEL_MIGRATE && kind != TRAVEL_COMMUTE) return engine_fail(e, EPI_ERR_ARG, "epi_travel_unpack: bad kind");
    if (!recv_buf || stride_records < 2 || stride_records > 0xFFFFFFFFull) return engine_fail(e, EPI_ERR_ARG, "epi_travel_unpack: receive buffer / stride_records");
    const uint32_t R = (uint32_t)e->n_regions, stride = (uint32_t)stride_records;
    CU(cudaSetDevice(e->device));
    // The number of arrivals is only known on the device (the segment headers); the kernels read it there and the grids are
    // sized for the most this region can take.  The placement rounds loop on the device until everybody has a cell.
    (void)R;
    TravelArgs A{};
    A.kind = kind;
    A.hour = hour;
    A.hour_of_day = hour % 24u;
    // the Counts row of an exchange hour that is queued (epi_enqueue_hour) is written by the device; a caller that drives the
    // exchange by hand gets it from epi_finish_hour
    A.row_index = 0xFFFFFFFFu;
    if (!e->pend_kind.empty() && e->pend_kind.back() == 2 && e->pend_first + (uint32_t)e->pend_kind.size() - 1u == hour) {
        A.row_index = hour - e->pend_first;
        e->pend_kind.back() = 3;  // collect_hours: this row is the device's
    }
    // every slot is visited (and drawn for) by a migration's leave kernel only: that one gets the wide grid, the rest wait less at their barriers
    if (!e->travel_blocks) { e->travel_blocks = travel_grid_blocks(e->device, 4); e->travel_blocks_small = travel_grid_blocks(e->device, 1); }
    {
        TimedGroup t(e);
        const cudaError_t cerr = launch_travel_arrive(e->P, e->D, A, e->T, (const TravelRecord*)recv_buf, stride, e->geo.n_houses, e->geo.n_offices, e->travel_blocks, wait_flags, exchange_no, e->stream);
        if (cerr != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("k_travel_arrive (cooperative launch): ") + cudaGetErrorString(cerr));
        note_launch(e);
    }
    CU(cudaGetLastError());
    e->unpack_unsettled = true;
    e->have_last_row = false;
    if (!counts_in) return EPI_OK;  // deferred
    const int rc = sync_travel(e);
    if (rc) return rc;
    for (uint32_t r = 0; r < R; ++r) counts_in[r] = e->h_tv->cnt[r];
    return EPI_OK;
}

// Peer transport: leave + push + wait + arrive of one exchange hour in a single cooperative launch (k_travel_exchange).  The hour
// must be one the travel plan exchanges at (epi_exchange_kind) and must have been queued with epi_enqueue_hour.
int epi_travel_exchange_fused(epi_engine* e, uint32_t hour, int kind, void* send_buf, const void* recv_buf, uint64_t stride_records, epi::TravelRecord* const* peer_recv,
                              uint32_t* const* peer_flags, const uint32_t* wait_flags, uint32_t exchange_no) {
    CU(cudaSetDevice(e->device));
    TravelArgs A{};
    A.kind = kind;
    A.hour = hour;
    A.hour_of_day = hour % 24u;
    A.row_index = 0xFFFFFFFFu;
    if (!e->pend_kind.empty() && e->pend_kind.back() == 2 && e->pend_first + (uint32_t)e->pend_kind.size() - 1u == hour) {
        A.row_index = hour - e->pend_first;
        e->pend_kind.back() = 3;  // collect_hours: this row is the device's
    }
    if (!e->travel_blocks) { e->travel_blocks = travel_grid_blocks(e->device, 4); e->travel_blocks_small = travel_grid_blocks(e->device, 1); }
    {
        TimedGroup t(e);
        const cudaError_t cerr = launch_travel_exchange(e->P, e->D, A, e->T, e->t_block_counts, (TravelRecord*)send_buf, (uint32_t)stride_records, (const TravelRecord*)recv_buf,
                                                        e->geo.n_houses, e->geo.n_offices, e->travel_blocks, peer_recv, peer_flags, wait_flags, exchange_no, e->stream);
        if (cerr != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("k_travel_exchange (cooperative launch): ") + cudaGetErrorString(cerr));
        note_launch(e);
    }
    CU(cudaGetLastError());
    e->have_last_row = false;
    e->pack_unsettled = e->unpack_unsettled = true;
    return EPI_OK;
}

extern "C" {

int epi_finish_hour(epi_engine* e, uint32_t hour, epi_counts* out) {
    if (!e || !out) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    {
        const int rc = sync_travel(e);
        if (rc) return rc;
    }
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
