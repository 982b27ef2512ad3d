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
    add(TERR_LIST_OVERFLOW, "more travellers than the exchange lists hold (travel plan underestimates the traffic)");
    add(TERR_SEGMENT_OVERFLOW, "a destination's records do not fit its segment of the send buffer: raise stride_records");
    add(TERR_NO_HOUSE, "Could not find house");
    add(TERR_NO_OFFICE, "Could not find office");
    add(TERR_HOUSES_FULL, "Couldn't find any house with free space!");
    add(TERR_OFFICES_FULL, "Couldn't find any offices with free space!");
    add(TERR_BAD_REGION, "commuter with a work / home region outside the travel plan");
    return m;
}

// zero headers: nobody travels from this region in this exchange
int write_empty_headers(epi_engine* e, void* send_buf, uint64_t stride) {
    if (!send_buf) return EPI_OK;
    CU(cudaMemset2DAsync(send_buf, (size_t)stride * sizeof(TravelRecord), 0, sizeof(TravelRecord), (size_t)e->n_regions, e->stream));
    return EPI_OK;
}

// What a deferred pack / unpack left for the host to settle (epi_finish_hour or the next synchronous call does it)
int settle_exchange(epi_engine* e) {
    if (!e->pack_unsettled && !e->unpack_unsettled) return EPI_OK;
    const uint32_t R = (uint32_t)e->n_regions;
    for (;;) {
        CU(cudaMemcpyAsync(e->h_tv, e->T.tv, (8 + R) * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaGetLastError());
        const uint32_t err = e->h_tv->err;
        if (err) e->pack_unsettled = e->unpack_unsettled = false;  // the exchange is abandoned (a pack that overflowed removed nobody)
        if (err & TERR_NO_SLOTS) return engine_fail(e, EPI_ERR_STATE, "region is out of agent slots: raise extra_capacity");
        if (err) return engine_fail(e, EPI_ERR_STATE, "traveller exchange: " + travel_error(err));
        if (!e->unpack_unsettled || e->h_tv->pending == 0) break;
        // select_starting_points: more placement rounds until every arrival holds a distinct vacant cell
        if (e->unpack_attempt > 64) return engine_fail(e, EPI_ERR_STATE, "Not enough locations are available for travellers");
        TimedGroup t(e);
        note_launch(e, launch_travel_rounds(e->P, e->D, e->unpack_args, e->T, e->unpack_max_arrivals, e->unpack_attempt, 4, e->stream));
        e->unpack_attempt += 4;
    }
    if (e->pack_unsettled) e->population -= e->h_tv->n_send;
    if (e->unpack_unsettled) {
        e->population += e->h_tv->n_in;
        launch_travel_arrivals_done(e->T, e->stream);  // the arrivals' slots leave the free stack
        note_launch(e);
    }
    e->pack_unsettled = e->unpack_unsettled = false;
    return EPI_OK;
}

}  // namespace

extern "C" {

int epi_travel_pack(epi_engine* e, uint32_t hour, int kind, void* send_buf, uint64_t stride_records, uint32_t* counts_out) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (!e->multi) return engine_fail(e, EPI_ERR_STATE, "not a multi-region engine (epi_create_multi)");
    if (kind != TRAVEL_MIGRATE && kind != TRAVEL_COMMUTE) return engine_fail(e, EPI_ERR_ARG, "epi_travel_pack: kind must be EPI_TRAVEL_MIGRATE or EPI_TRAVEL_COMMUTE");
    if (!send_buf || stride_records < 2 || stride_records > 0xFFFFFFFFull) return engine_fail(e, EPI_ERR_ARG, "epi_travel_pack: send buffer / stride_records");
    const uint32_t R = (uint32_t)e->n_regions, h = hour % 24u;
    if (counts_out) std::fill(counts_out, counts_out + R, 0u);
    CU(cudaSetDevice(e->device));
    int rc = settle_exchange(e);  // an earlier deferred exchange (its epi_finish_hour was skipped)
    if (rc) return rc;
    TravelArgs A{};
    A.kind = kind;
    A.hour = hour;
    A.hour_of_day = h;
    if (kind == TRAVEL_MIGRATE) {
        // Citizen::can_migrate's hour window (citizen/mod.rs:460-462); EngineMigrationPlan::percent_outgoing (:44-49)
        uint64_t planned_total = 0;
        for (uint32_t v : e->migration_row) planned_total += v;
        if (!e->migration_enabled || h != 0 || !(hour > e->start_migration_hour && hour < e->end_migration_hour) || planned_total == 0 || e->population == 0)
            return write_empty_headers(e, send_buf, stride_records);
        // gen_bool(percent_outgoing) panics for p > 1 (rand 0.8 Bernoulli::new; allocation_map.rs:110): an error here
        if (planned_total > e->population)
            return engine_fail(e, EPI_ERR_STATE, "migration plan row (" + std::to_string(planned_total) + " outgoing) exceeds the region's population (" +
                                                     std::to_string(e->population) + "): percent_outgoing > 1");
        A.thr_outgoing = bernoulli_threshold((double)planned_total / (double)e->population);
    } else if (!e->commute_enabled || !(h == 7 || h == 17)) {
        return write_empty_headers(e, send_buf, stride_records);
    }
    CU(cudaMemsetAsync(&e->T.tv->err, 0, sizeof(uint32_t), e->stream));
    {
        TimedGroup t(e);
        note_launch(e, launch_travel_leave(e->P, e->D, A, e->T, e->t_block_counts, (TravelRecord*)send_buf, (uint32_t)stride_records, e->stream));
    }
    e->have_last_row = false;
    e->pack_unsettled = true;
    if (!counts_out) return EPI_OK;  // deferred: the records are in flight on the stream, the host settles later
    rc = settle_exchange(e);
    if (rc) return rc;
    for (uint32_t to = 0; to < R; ++to) counts_out[to] = e->h_tv->cnt[to];  // still the leavers' counts: no unpack ran in between
    return EPI_OK;
}

int epi_travel_unpack(epi_engine* e, uint32_t hour, int kind, const void* recv_buf, uint64_t stride_records, uint32_t* counts_in) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (!e->multi) return engine_fail(e, EPI_ERR_STATE, "not a multi-region engine (epi_create_multi)");
    if (kind != TRAVEL_MIGRATE && kind != TRAVEL_COMMUTE) return engine_fail(e, EPI_ERR_ARG, "epi_travel_unpack: bad kind");
    if (!recv_buf || stride_records < 2 || stride_records > 0xFFFFFFFFull) return engine_fail(e, EPI_ERR_ARG, "epi_travel_unpack: receive buffer / stride_records");
    const uint32_t R = (uint32_t)e->n_regions, stride = (uint32_t)stride_records;
    CU(cudaSetDevice(e->device));
    if (e->unpack_unsettled) {
        const int rc = settle_exchange(e);
        if (rc) return rc;
    }
    // The number of arrivals is only known on the device (the segment headers); the kernels read it there and the grids are
    // sized for the most this region can take, so the host need not synchronise before the placement rounds have run.
    e->unpack_max_arrivals = (uint32_t)std::min<uint64_t>(e->T.list_cap, (uint64_t)R * (stride - 1u));
    TravelArgs& A = e->unpack_args;
    A = TravelArgs{};
    A.kind = kind;
    A.hour = hour;
    A.hour_of_day = hour % 24u;
    if (!e->pack_unsettled) CU(cudaMemsetAsync(&e->T.tv->err, 0, sizeof(uint32_t), e->stream));
    {
        TimedGroup t(e);
        note_launch(e, launch_travel_arrive(e->P, e->D, A, e->T, (const TravelRecord*)recv_buf, stride, e->unpack_max_arrivals, e->geo.n_houses, e->geo.n_offices, e->stream));
        note_launch(e, launch_travel_rounds(e->P, e->D, A, e->T, e->unpack_max_arrivals, 0, 3, e->stream));
    }
    e->unpack_attempt = 3;
    e->unpack_unsettled = true;
    e->have_last_row = false;
    if (!counts_in) return EPI_OK;  // deferred
    const int rc = settle_exchange(e);
    if (rc) return rc;
    for (uint32_t r = 0; r < R; ++r) counts_in[r] = e->h_tv->cnt[r];
    return EPI_OK;
}

int epi_finish_hour(epi_engine* e, uint32_t hour, epi_counts* out) {
    if (!e || !out) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    {
        const int rc = settle_exchange(e);
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
