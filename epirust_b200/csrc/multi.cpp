// The Transport of a multi-region run (NCCL all-to-allv of packed traveller records, or device copies when every region is
// hosted by one process) and Epidemiology::run_multi_engine's hour loop on top of the region engines.
// Reference: engine/src/transport/mod.rs:34-42, transport/mpi_transport.rs:44-215, epidemiology_simulation.rs:276-547,
// orchestrator/src/ticks.rs:35-89,175-180.  C ABI: the "multi-region" block of include/epi.h.
#include "multi.h"
#include "kernels.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace epi;

#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t _err = (call);                                                                                 \
        if (_err != cudaSuccess)                                                                                   \
            return engine_fail(e, EPI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_err));              \
    } while (0)
#define NC(call)                                                                                                   \
    do {                                                                                                           \
        ncclResult_t _err = (call);                                                                                \
        if (_err != ncclSuccess)                                                                                   \
            return engine_fail(e, EPI_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(_err));              \
    } while (0)

namespace epi {

Comm::~Comm() {
    for (RegionBuffers& b : buffers) {
        cudaSetDevice(b.device);
        if (b.send) cudaFree(b.send);
        if (b.recv) cudaFree(b.recv);
        for (int k = 0; k < 2; ++k)
            if (b.d_caps[k]) cudaFree(b.d_caps[k]);
        if (b.packed) cudaEventDestroy(b.packed);
        if (b.copied) cudaEventDestroy(b.copied);
    }
    for (void* m : peer_mapped) cudaIpcCloseMemHandle(m);
    if (recv2) cudaFree(recv2);
    if (flags) cudaFree(flags);
    if (d_peer_recv) cudaFree(d_peer_recv);
    if (d_peer_flags) cudaFree(d_peer_flags);
    if (d_sum) cudaFree(d_sum);
    if (h_sum) cudaFreeHost(h_sum);
    if (nccl) ncclCommDestroy(nccl);
}

int exchange_kind_of(bool migration_enabled, bool commute_enabled, uint32_t start_migration_hour, uint32_t end_migration_hour, uint32_t hour) {
    const uint32_t h = hour % 24u;
    if (commute_enabled && (h == 7u || h == 17u)) return EPI_TRAVEL_COMMUTE;  // constants.rs ROUTINE_TRAVEL_START_TIME / END_TIME
    // a migration hour outside the window moves nobody in any region (Citizen::can_migrate), so every rank skips it alike
    if (migration_enabled && h == 0u && hour > start_migration_hour && hour < end_migration_hour) return EPI_TRAVEL_MIGRATE;
    return -1;
}

bool is_tick_hour(bool migration_enabled, bool commute_enabled, uint32_t hour) {
    const uint32_t h = hour % 24u;
    if (!commute_enabled && (h == 7u || h == 17u)) return false;
    if (!migration_enabled && h == 0u) return false;
    return hour <= 1u || h == 0u || h == 7u || h == 17u;
}

int run_multi_schedule(std::vector<RegionOps*>& regions, ExchangeOps& x, const PlanInfo& plan, uint32_t first_hour, uint32_t n_hours, bool terminate_when_clear,
                       epi_counts* rows_out, uint32_t* n_rows) {
    *n_rows = 0;
    if (n_hours == 0) return EPI_OK;
    auto kind_of = [&](uint32_t hour) { return exchange_kind_of(plan.migration_enabled, plan.commute_enabled, plan.start_migration_hour, plan.end_migration_hour, hour); };
    auto tick = [&](uint32_t hour) { return is_tick_hour(plan.migration_enabled, plan.commute_enabled, hour); };
    uint32_t hour = first_hour, last = first_hour + n_hours - 1u;
    std::vector<epi_counts> got;
    while (hour <= last) {
        // The host waits for the device only where it has to look at Counts: at a decision hour of process_interventions (start of
        // day, vaccination hour, unlock hour); with the termination rule also at a tick hour.  An exchange needs no wait: the
        // counts, the population and the errors of the exchange stay on the device, and the exchange hour's Counts row reaches
        // the counts ring like every other row.
        uint32_t decision = 0xFFFFFFFFu;
        for (RegionOps* r : regions) decision = std::min(decision, r->next_decision_hour(hour));
        if (terminate_when_clear)
            for (uint32_t h = hour; h <= last && h < hour + 24u; ++h)
                if (tick(h)) { decision = std::min(decision, h); break; }
        const uint32_t seg_end = std::min(last, decision);  // last hour queued in this round
        uint32_t h = hour;
        while (h <= seg_end) {
            uint32_t xh = h;
            while (xh <= seg_end && kind_of(xh) < 0) ++xh;  // the next exchange hour of the round, or seg_end + 1
            if (xh > h)
                for (RegionOps* r : regions) {
                    const int rc = r->enqueue_hours(h, xh - h);
                    if (rc) return rc;
                }
            if (xh <= seg_end) {
                for (RegionOps* r : regions) {
                    const int rc = r->enqueue_hour(xh);
                    if (rc) return rc;
                }
                const int rc = x.exchange(xh, kind_of(xh));
                if (rc) return rc;
            }
            h = xh + 1u;
        }
        unsigned long long active = 0;
        for (size_t i = 0; i < regions.size(); ++i) {
            const int rc = regions[i]->collect(got);
            if (rc) return rc;
            if (got.size() != seg_end - hour + 1u) return EPI_ERR_STATE;
            epi_counts* out = rows_out + i * (size_t)n_hours + (hour - first_hour);
            for (size_t k = 0; k < got.size(); ++k) out[k] = got[k];
            const epi_counts& c = got.back();
            active += (unsigned long long)c.exposed + c.infected + c.hospitalized;
        }
        if (terminate_when_clear && tick(seg_end)) {
            // TickAcks::should_terminate (ticks.rs:175-180): the NEXT tick carries terminate = true and the engines break before
            // simulating its hour (epidemiology_simulation.rs:336-349)
            unsigned long long total = 0;
            const int rc = x.all_reduce_sum(active, &total);
            if (rc) return rc;
            if (total == 0) {
                uint32_t t = seg_end + 1u;
                while (t <= last && !tick(t)) ++t;
                if (t <= last) last = t - 1u;
            }
        }
        hour = seg_end + 1u;
    }
    *n_rows = last - first_hour + 1u;
    return EPI_OK;
}

}  // namespace epi

namespace {

PlanInfo plan_of(const epi_engine* e) {
    PlanInfo p;
    p.migration_enabled = e->migration_enabled;
    p.commute_enabled = e->commute_enabled;
    p.start_migration_hour = e->start_migration_hour;
    p.end_migration_hour = e->end_migration_hour;
    return p;
}

// records (header included) region `from` can send to region `to` in one exchange of `kind`
void segment_caps(const epi_engine* e, std::vector<uint32_t> cap[2], uint32_t* stride) {
    const size_t R = (size_t)e->n_regions;
    cap[0].assign(R * R, 1u);
    cap[1].assign(R * R, 1u);
    if (e->migration_enabled)
        for (size_t from = 0; from < R; ++from) {
            // outgoing total ~ Binomial(eligible agents, sum(row) / population): mean <= sum(row), sigma <= sqrt(sum(row)); region `to`
            // takes floor(share * total) of it (engine_migration_plan.rs:51-77): matrix entry + 8 sigma bounds it
            uint64_t row = 0;
            for (size_t to = 0; to < R; ++to) row += e->migration_mat[from * R + to];
            const uint32_t slack = 8u * (uint32_t)std::ceil(std::sqrt((double)row)) + 16u;
            for (size_t to = 0; to < R; ++to) cap[EPI_TRAVEL_MIGRATE][from * R + to] = e->migration_mat[from * R + to] + slack + 1u;
        }
    if (e->commute_enabled)
        for (size_t from = 0; from < R; ++from)
            for (size_t to = 0; to < R; ++to)  // 07:00: the matrix entry leaves (citizen_factory.rs:90-110); 17:00: at most those who came return
                cap[EPI_TRAVEL_COMMUTE][from * R + to] = std::max(e->commute_mat[from * R + to], e->commute_mat[to * R + from]) + 1u;
    uint32_t s = 2;
    for (int k = 0; k < 2; ++k)
        for (uint32_t v : cap[k]) s = std::max(s, v);
    *stride = s;
}

int alloc_region_buffers(epi_engine* e, Comm& c, RegionBuffers& b) {
    const size_t R = (size_t)c.n, bytes = R * c.stride * sizeof(TravelRecord);
    b.device = e->device;
    CU(cudaSetDevice(e->device));
    CU(cudaMalloc((void**)&b.send, bytes));
    CU(cudaMalloc((void**)&b.recv, bytes));
    CU(cudaMemset(b.send, 0, bytes));
    CU(cudaMemset(b.recv, 0, bytes));
    for (int k = 0; k < 2; ++k) {
        CU(cudaMalloc((void**)&b.d_caps[k], R * sizeof(uint32_t)));
        CU(cudaMemcpy(b.d_caps[k], c.cap[k].data() + (size_t)e->P.region * R, R * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    CU(cudaEventCreateWithFlags(&b.packed, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b.copied, cudaEventDisableTiming));
    return EPI_OK;
}

// Peer transport: exchange CUDA IPC handles of the receive areas and flag words through the communicator, map every peer's.
// When the ranks cannot reach each other's memory (no peer access between the devices) peer_ok stays false on every rank and the
// exchange uses ncclSend / ncclRecv.
int setup_peer_transport(epi_engine* e, Comm& c) {
    const size_t R = (size_t)c.n, me = (size_t)c.rank;
    CU(cudaSetDevice(e->device));
    CU(cudaMalloc((void**)&c.recv2, 2 * R * c.stride * sizeof(TravelRecord)));
    CU(cudaMalloc((void**)&c.flags, R * sizeof(uint32_t)));
    CU(cudaMemset(c.recv2, 0, 2 * R * c.stride * sizeof(TravelRecord)));
    CU(cudaMemset(c.flags, 0, R * sizeof(uint32_t)));
    struct Card { cudaIpcMemHandle_t recv, flags; int device; int ok; };
    static_assert(sizeof(Card) % 4 == 0, "Card is sent as words");
    Card mine{};
    mine.device = e->device;
    mine.ok = cudaIpcGetMemHandle(&mine.recv, c.recv2) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags, c.flags) == cudaSuccess;
    cudaGetLastError();
    Card* d_cards = nullptr;
    CU(cudaMalloc((void**)&d_cards, R * sizeof(Card)));
    CU(cudaMemcpyAsync(d_cards + me, &mine, sizeof(Card), cudaMemcpyHostToDevice, e->stream));
    NC(ncclAllGather(d_cards + me, d_cards, sizeof(Card), ncclChar, c.nccl, e->stream));
    std::vector<Card> cards(R);
    CU(cudaMemcpyAsync(cards.data(), d_cards, R * sizeof(Card), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(d_cards);
    std::vector<TravelRecord*> peer_recv(R, nullptr);
    std::vector<uint32_t*> peer_flags(R, nullptr);
    int ok = 1;
    for (size_t p = 0; p < R; ++p) ok = ok && cards[p].ok;
    for (size_t p = 0; p < R && ok; ++p) {
        if (p == me) { peer_recv[p] = c.recv2; peer_flags[p] = c.flags; continue; }
        void *r = nullptr, *f = nullptr;
        if (cudaIpcOpenMemHandle(&r, cards[p].recv, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        c.peer_mapped.push_back(r);
        if (cudaIpcOpenMemHandle(&f, cards[p].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
        c.peer_mapped.push_back(f);
        peer_recv[p] = (TravelRecord*)r;
        peer_flags[p] = (uint32_t*)f;
    }
    cudaGetLastError();
    // all ranks take the same path: one that could not map a peer sends everybody back to NCCL
    int* d_ok = nullptr;
    CU(cudaMalloc((void**)&d_ok, sizeof(int)));
    CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    NC(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c.nccl, e->stream));
    CU(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(d_ok);
    if (!ok) return EPI_OK;
    CU(cudaMalloc((void**)&c.d_peer_recv, R * sizeof(TravelRecord*)));
    CU(cudaMalloc((void**)&c.d_peer_flags, R * sizeof(uint32_t*)));
    CU(cudaMemcpy(c.d_peer_recv, peer_recv.data(), R * sizeof(TravelRecord*), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c.d_peer_flags, peer_flags.data(), R * sizeof(uint32_t*), cudaMemcpyHostToDevice));
    c.peer_ok = true;
    return EPI_OK;
}

int pack_deferred(epi_engine* e, Comm& c, RegionBuffers& b, uint32_t hour, int kind) {
    e->T.seg_cap = b.d_caps[kind];
    const int rc = epi_travel_pack(e, hour, kind, b.send, c.stride, nullptr);
    e->T.seg_cap = nullptr;
    return rc;
}

// TravelCounter (listeners/travel_counter.rs:27-92): the migrators that left, by destination and state -- the used part of the
// send segments is copied to the host behind the pack kernels; tallied after the hour's wait (note_outgoing)
int stage_outgoing(epi_engine* e, Comm& c, RegionBuffers& b, uint32_t hour, int kind) {
    if (!e->count_outgoing || kind != EPI_TRAVEL_MIGRATE) return EPI_OK;
    const size_t R = (size_t)c.n, bytes = R * c.stride * sizeof(TravelRecord);
    if (e->h_outgoing_bytes < bytes) {
        if (e->h_outgoing) cudaFreeHost(e->h_outgoing);
        e->h_outgoing = nullptr;
        CU(cudaMallocHost((void**)&e->h_outgoing, bytes));
        e->h_outgoing_bytes = bytes;
    }
    const uint32_t* cap = c.cap[kind].data() + (size_t)e->P.region * R;
    for (size_t to = 0; to < R; ++to)
        CU(cudaMemcpyAsync(e->h_outgoing + to * c.stride * sizeof(TravelRecord), b.send + to * c.stride, (size_t)cap[to] * sizeof(TravelRecord),
                           cudaMemcpyDeviceToHost, e->stream));
    e->outgoing_staged_hour = hour;
    e->outgoing_staged_stride = c.stride;
    return EPI_OK;
}

// Listener::outgoing_migrators_added (epidemiology_simulation.rs:413-415): at every hour % 24 == 0 of a migration-enabled run one
// CountsByRegion per region the plan sends migrators to (engine_migration_plan.rs:57-60), all zero outside the migration window
void note_outgoing(epi_engine* e, uint32_t hour) {
    if (!e->count_outgoing || !e->migration_enabled || hour % 24u != 0) return;
    const size_t R = (size_t)e->n_regions;
    const bool staged = e->outgoing_staged_hour == hour;
    const TravelRecord* seg = reinterpret_cast<const TravelRecord*>(e->h_outgoing);
    for (size_t to = 0; to < R; ++to) {
        if ((int)to == e->P.region || e->migration_row[to] == 0) continue;
        uint32_t by_state[4] = {0, 0, 0, 0};
        if (staged) {
            const TravelRecord* s = seg + to * e->outgoing_staged_stride;
            for (uint32_t k = 0; k < s[0].st; ++k) {
                const uint32_t state = s[1 + k].st & ST_STATE_MASK;
                if (state < 4) by_state[state]++;
            }
        }
        e->outgoing_travels.push_back({hour, (uint32_t)to, by_state[0], by_state[1], by_state[2], by_state[3]});
    }
    e->outgoing_staged_hour = 0;
}

int nccl_exchange(epi_engine* e, uint32_t hour, int kind) {
    Comm& c = *e->comm;
    RegionBuffers& b = c.buffers[0];
    const size_t R = (size_t)c.n, me = (size_t)c.rank;
    CU(cudaSetDevice(e->device));
    if (c.peer_ok) {
        // every destination's segment goes straight into that rank's receive area over NVLink (the tail of the leave kernel); the
        // receiver's arrive kernel waits for our flag
        const uint32_t no = ++c.exchange_no;
        e->T.seg_cap = b.d_caps[kind];
        const int rc = epi_travel_exchange_fused(e, hour, kind, b.send, c.recv2 + (size_t)(no & 1u) * R * c.stride, c.stride, c.d_peer_recv, c.d_peer_flags, c.flags, no);
        e->T.seg_cap = nullptr;
        if (rc) return rc;
        return stage_outgoing(e, c, b, hour, kind);
    }
    int rc = pack_deferred(e, c, b, hour, kind);
    if (rc) return rc;
    rc = stage_outgoing(e, c, b, hour, kind);
    if (rc) return rc;
    // all-to-allv: segment p of the send buffer -> rank p, segment p of the receive buffer <- rank p; only the records the plan
    // can produce for the pair travel (the header carries the actual count)
    static const bool skip_nccl = std::getenv("EPI_DEBUG_NO_NCCL") != nullptr;  // timing experiments only: nobody arrives
    if (skip_nccl) {
        CU(cudaMemsetAsync(b.recv, 0, R * c.stride * sizeof(TravelRecord), e->stream));
        return epi_travel_unpack(e, hour, kind, b.recv, c.stride, nullptr);
    }
    NC(ncclGroupStart());
    for (size_t p = 0; p < R; ++p) {
        if (p == me) continue;
        NC(ncclSend(b.send + p * c.stride, (size_t)c.cap[kind][me * R + p] * sizeof(TravelRecord), ncclChar, (int)p, c.nccl, e->stream));
        NC(ncclRecv(b.recv + p * c.stride, (size_t)c.cap[kind][p * R + me] * sizeof(TravelRecord), ncclChar, (int)p, c.nccl, e->stream));
    }
    NC(ncclGroupEnd());
    CU(cudaMemcpyAsync(b.recv + me * c.stride, b.send + me * c.stride, (size_t)c.cap[kind][me * R + me] * sizeof(TravelRecord), cudaMemcpyDeviceToDevice, e->stream));
    return epi_travel_unpack(e, hour, kind, b.recv, c.stride, nullptr);
}

int local_exchange(Comm& c, uint32_t hour, int kind) {
    const size_t R = (size_t)c.n;
    for (size_t s = 0; s < R; ++s) {
        epi_engine* e = c.engines[s];
        CU(cudaSetDevice(e->device));
        int rc = pack_deferred(e, c, c.buffers[s], hour, kind);
        if (rc) return rc;
        rc = stage_outgoing(e, c, c.buffers[s], hour, kind);
        if (rc) return rc;
        CU(cudaEventRecord(c.buffers[s].packed, e->stream));
    }
    for (size_t r = 0; r < R; ++r) {  // region r receives segment r of every source, in source order
        epi_engine* e = c.engines[r];
        CU(cudaSetDevice(e->device));
        for (size_t s = 0; s < R; ++s) {
            if (s != r) CU(cudaStreamWaitEvent(e->stream, c.buffers[s].packed, 0));
            CU(cudaMemcpyAsync(c.buffers[r].recv + s * c.stride, c.buffers[s].send + r * c.stride, (size_t)c.cap[kind][s * R + r] * sizeof(TravelRecord),
                               cudaMemcpyDefault, e->stream));
        }
        CU(cudaEventRecord(c.buffers[r].copied, e->stream));
        const int rc = epi_travel_unpack(e, hour, kind, c.buffers[r].recv, c.stride, nullptr);
        if (rc) return rc;
    }
    for (size_t s = 0; s < R; ++s) {  // a send buffer may be rewritten only after every reader has copied its segment
        epi_engine* e = c.engines[s];
        CU(cudaSetDevice(e->device));
        for (size_t r = 0; r < R; ++r)
            if (r != s) CU(cudaStreamWaitEvent(e->stream, c.buffers[r].copied, 0));
    }
    return EPI_OK;
}

struct EngineOps : RegionOps {
    epi_engine* e;
    std::vector<epi_counts> buf;
    explicit EngineOps(epi_engine* e_) : e(e_), buf(RING_ROWS) {}
    uint32_t next_decision_hour(uint32_t hour) override { return epi_next_decision_hour(e, hour); }
    int enqueue_hours(uint32_t first_hour, uint32_t n) override { return epi_enqueue_hours(e, first_hour, n); }
    int enqueue_hour(uint32_t hour) override { return epi_enqueue_hour(e, hour); }
    int collect(std::vector<epi_counts>& rows) override {
        uint32_t n = 0;
        const int rc = epi_collect_hours(e, buf.data(), (uint32_t)buf.size(), &n);
        rows.assign(buf.begin(), buf.begin() + n);
        if (!rc)
            for (const epi_counts& c : rows) note_outgoing(e, c.hour);
        return rc;
    }
};

struct CommExchange : ExchangeOps {
    Comm& c;
    explicit CommExchange(Comm& c_) : c(c_) {}
    int exchange(uint32_t hour, int kind) override { return c.nccl ? nccl_exchange(c.engines[0], hour, kind) : local_exchange(c, hour, kind); }
    int all_reduce_sum(unsigned long long local, unsigned long long* total) override {
        *total = local;
        if (!c.nccl) return EPI_OK;
        epi_engine* e = c.engines[0];
        CU(cudaSetDevice(e->device));
        *c.h_sum = local;
        CU(cudaMemcpyAsync(c.d_sum, c.h_sum, sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
        NC(ncclAllReduce(c.d_sum, c.d_sum, 1, ncclUint64, ncclSum, c.nccl, e->stream));
        CU(cudaMemcpyAsync(c.h_sum, c.d_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        *total = *c.h_sum;
        return EPI_OK;
    }
};

// recording stand-ins for epi_multi_schedule_trace
struct TraceRegion : RegionOps {
    std::string& out;
    std::vector<uint32_t> vaccinate;
    uint32_t unlock_hour;
    std::vector<uint32_t> queued;
    TraceRegion(std::string& o, std::vector<uint32_t> v, uint32_t u) : out(o), vaccinate(std::move(v)), unlock_hour(u) {}
    uint32_t next_decision_hour(uint32_t hour) override {
        uint32_t d = (hour + 23u) / 24u * 24u;
        for (uint32_t v : vaccinate)
            if (v >= hour) d = std::min(d, v);
        if (unlock_hour && unlock_hour >= hour) d = std::min(d, unlock_hour);
        return d;
    }
    int enqueue_hours(uint32_t first_hour, uint32_t n) override {
        out += "hours " + std::to_string(first_hour) + " " + std::to_string(n) + "\n";
        for (uint32_t k = 0; k < n; ++k) queued.push_back(first_hour + k);
        return EPI_OK;
    }
    int enqueue_hour(uint32_t hour) override {
        out += "exchange_hour " + std::to_string(hour) + "\n";
        queued.push_back(hour);
        return EPI_OK;
    }
    int collect(std::vector<epi_counts>& rows) override {
        out += "collect";
        rows.clear();
        for (uint32_t h : queued) {
            out += " " + std::to_string(h);
            rows.push_back(epi_counts{h, 1, 0, 0, 0, 0, 0});
        }
        out += "\n";
        queued.clear();
        return EPI_OK;
    }
};
struct TraceExchange : ExchangeOps {
    std::string& out;
    explicit TraceExchange(std::string& o) : out(o) {}
    int exchange(uint32_t hour, int kind) override {
        out += "exchange " + std::to_string(hour) + " " + std::to_string(kind) + "\n";
        return EPI_OK;
    }
    int all_reduce_sum(unsigned long long local, unsigned long long* total) override {
        *total = local;
        return EPI_OK;
    }
};

}  // namespace

extern "C" {

int epi_comm_unique_id(void* id_out) {
    epi_engine* e = nullptr;
    if (!id_out) return engine_fail(e, EPI_ERR_ARG, "null argument");
    static_assert(sizeof(ncclUniqueId) == EPI_COMM_ID_BYTES, "EPI_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
    ncclUniqueId id;
    NC(ncclGetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof(id));
    return EPI_OK;
}

int epi_comm_init(epi_engine* e, int n_ranks, int rank, const void* unique_id) {
    if (!e || !unique_id) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (!e->multi) return engine_fail(e, EPI_ERR_STATE, "not a multi-region engine (epi_create_multi)");
    if (n_ranks != e->n_regions || rank != e->P.region)
        return engine_fail(e, EPI_ERR_ARG, "epi_comm_init: n_ranks / rank must be the travel plan's region count / this engine's region index");
    if (e->comm) return engine_fail(e, EPI_ERR_STATE, "the engine already has a communicator");
    auto c = std::make_shared<Comm>();
    c->n = n_ranks;
    c->rank = rank;
    segment_caps(e, c->cap, &c->stride);
    c->engines.push_back(e);
    c->buffers.resize(1);
    int rc = alloc_region_buffers(e, *c, c->buffers[0]);
    if (rc) return rc;
    CU(cudaMalloc((void**)&c->d_sum, sizeof(unsigned long long)));
    CU(cudaMallocHost((void**)&c->h_sum, sizeof(unsigned long long)));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    NC(ncclCommInitRank(&c->nccl, n_ranks, id, rank));
    e->comm = c;
    if (!std::getenv("EPI_NO_PEER")) {
        rc = setup_peer_transport(e, *c);
        if (rc) { e->comm.reset(); return rc; }
    }
    return EPI_OK;
}

int epi_comm_init_local(epi_engine* const* engines, int n_engines) {
    epi_engine* e = (engines && n_engines > 0) ? engines[0] : nullptr;
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null argument");
    auto c = std::make_shared<Comm>();
    c->n = n_engines;
    for (int r = 0; r < n_engines; ++r) {
        epi_engine* er = engines[r];
        if (!er || !er->multi || er->n_regions != n_engines || er->P.region != r)
            return engine_fail(e, EPI_ERR_ARG, "epi_comm_init_local: engines[r] must be region r of a plan with n_engines regions");
        if (er->comm) return engine_fail(e, EPI_ERR_STATE, "an engine already has a communicator");
        c->engines.push_back(er);
    }
    segment_caps(e, c->cap, &c->stride);
    c->buffers.resize((size_t)n_engines);
    for (int r = 0; r < n_engines; ++r) {
        const int rc = alloc_region_buffers(engines[r], *c, c->buffers[(size_t)r]);
        if (rc) return rc == EPI_OK ? rc : engine_fail(e, rc, engines[r]->err);
    }
    for (int r = 0; r < n_engines; ++r) engines[r]->comm = c;
    return EPI_OK;
}

int epi_comm_destroy(epi_engine* e) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    if (e->comm) {
        CU(cudaSetDevice(e->device));
        CU(cudaStreamSynchronize(e->stream));
        if (!e->comm->nccl)  // a local transport is shared: every region lets go of it
            for (epi_engine* other : std::vector<epi_engine*>(e->comm->engines))
                if (other != e) other->comm.reset();
        e->comm.reset();
    }
    return EPI_OK;
}

int epi_exchange_kind(const epi_engine* e, uint32_t hour) {
    if (!e || !e->multi) return -1;
    return exchange_kind_of(e->migration_enabled, e->commute_enabled, e->start_migration_hour, e->end_migration_hour, hour);
}

int epi_exchange(epi_engine* e, uint32_t hour, int kind) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    if (!e->comm || !e->comm->nccl) return engine_fail(e, EPI_ERR_STATE, "epi_exchange needs an NCCL communicator (epi_comm_init)");
    if (kind != EPI_TRAVEL_MIGRATE && kind != EPI_TRAVEL_COMMUTE) return engine_fail(e, EPI_ERR_ARG, "epi_exchange: bad kind");
    return nccl_exchange(e, hour, kind);
}

int epi_run_multi_hours(epi_engine* const* engines, int n_local, uint32_t first_hour, uint32_t n_hours, int terminate_when_clear, epi_counts* rows_out,
                        uint32_t* n_rows) {
    epi_engine* e = (engines && n_local > 0) ? engines[0] : nullptr;
    if (!e || !n_rows || (!rows_out && n_hours)) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (!e->comm) return engine_fail(e, EPI_ERR_STATE, "epi_run_multi_hours needs a communicator (epi_comm_init / epi_comm_init_local)");
    Comm& c = *e->comm;
    if (c.nccl ? n_local != 1 : n_local != c.n) return engine_fail(e, EPI_ERR_ARG, "epi_run_multi_hours: pass the one engine of this rank, or every region of a local communicator");
    for (int r = 0; r < n_local; ++r)
        if (!engines[r] || engines[r]->comm != e->comm || (!c.nccl && engines[r] != c.engines[(size_t)r]))
            return engine_fail(e, EPI_ERR_ARG, "epi_run_multi_hours: engines do not belong to one communicator, in region order");
    std::vector<EngineOps> ops;
    ops.reserve((size_t)n_local);
    std::vector<RegionOps*> regions;
    for (int r = 0; r < n_local; ++r) { ops.emplace_back(engines[r]); regions.push_back(&ops.back()); }
    CommExchange x(c);
    uint32_t done = 0;
    *n_rows = 0;
    // chunks of at most 240 hours keep the Counts ring (RING_ROWS) far from full
    while (done < n_hours) {
        const uint32_t n = std::min(n_hours - done, 240u);
        uint32_t got = 0;
        std::vector<epi_counts> rows((size_t)n_local * n);
        const int rc = run_multi_schedule(regions, x, plan_of(e), first_hour + done, n, terminate_when_clear != 0, rows.data(), &got);
        if (rc) {
            for (int r = 1; r < n_local; ++r)
                if (!engines[r]->err.empty() && e->err.empty()) e->err = engines[r]->err;
            return rc;
        }
        for (int r = 0; r < n_local; ++r)
            std::copy(rows.begin() + (size_t)r * n, rows.begin() + (size_t)r * n + got, rows_out + (size_t)r * n_hours + done);
        done += got;
        if (got < n) break;  // the termination rule fired
    }
    *n_rows = done;
    return EPI_OK;
}

int epi_count_outgoing(epi_engine* e, int on) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    e->count_outgoing = on != 0;
    return EPI_OK;
}

int epi_outgoing_travels(const epi_engine* e, epi_outgoing_travel* out, uint32_t max_rows, uint32_t* n) {
    if (!e || !n) return engine_fail(e, EPI_ERR_ARG, "null argument");
    *n = (uint32_t)e->outgoing_travels.size();
    if (out)
        for (uint32_t i = 0; i < std::min(max_rows, *n); ++i) out[i] = e->outgoing_travels[i];
    return EPI_OK;
}

int epi_should_terminate(const epi_counts* acks, int n_acks) {
    if (!acks) return 0;
    uint64_t exposed = 0, infected = 0, hospitalized = 0;
    for (int i = 0; i < n_acks; ++i) { exposed += acks[i].exposed; infected += acks[i].infected; hospitalized += acks[i].hospitalized; }
    return exposed == 0 && infected == 0 && hospitalized == 0;
}

int epi_multi_schedule_trace(const epi_travel_plan* plan, const uint32_t* vaccinate_hours, int n_vaccinate, uint32_t unlock_hour, uint32_t first_hour,
                             uint32_t n_hours, char* out, uint64_t out_bytes) {
    epi_engine* e = nullptr;
    if (!plan || !out || out_bytes == 0 || (n_vaccinate > 0 && !vaccinate_hours)) return engine_fail(e, EPI_ERR_ARG, "null argument");
    std::string text;
    TraceRegion region(text, std::vector<uint32_t>(vaccinate_hours, vaccinate_hours + std::max(n_vaccinate, 0)), unlock_hour);
    TraceExchange x(text);
    std::vector<RegionOps*> regions{&region};
    PlanInfo p;
    p.migration_enabled = plan->migration_enabled != 0;
    p.commute_enabled = plan->commute_enabled != 0;
    p.start_migration_hour = plan->start_migration_hour;
    p.end_migration_hour = plan->end_migration_hour;
    std::vector<epi_counts> rows(n_hours);
    uint32_t n_rows = 0;
    const int rc = run_multi_schedule(regions, x, p, first_hour, n_hours, false, rows.data(), &n_rows);
    if (rc) return rc;
    for (uint32_t k = 0; k < n_rows; ++k)
        if (rows[k].hour != first_hour + k) return engine_fail(e, EPI_ERR_STATE, "schedule: the row of hour " + std::to_string(first_hour + k) + " is missing or misplaced");
    if (text.size() + 1 > out_bytes) return engine_fail(e, EPI_ERR_ARG, "epi_multi_schedule_trace: out too small");
    std::memcpy(out, text.c_str(), text.size() + 1);
    return EPI_OK;
}

}  // extern "C"
