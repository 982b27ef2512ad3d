// The Transport of a multi-region run and the multi-engine hour loop (engine/src/transport/mod.rs:34-42,
// transport/mpi_transport.rs:44-215, epidemiology_simulation.rs:276-547) -- see the "multi-region" block of include/epi.h.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <string>
#include <vector>

#include "engine.h"

namespace epi {

// The segments of one region: what it sends to / receives from every region per exchange, in device memory.
struct RegionBuffers {
    TravelRecord *send = nullptr, *recv = nullptr;  // [n][stride]
    uint32_t* d_caps[2] = {nullptr, nullptr};        // per kind [n]: records (header included) segment `to` may hold
    cudaEvent_t packed = nullptr, copied = nullptr;  // local transport: ordering between the regions' streams
    int device = 0;
};

struct Comm {
    int n = 0;                 // regions == ranks
    ncclComm_t nccl = nullptr; // one process per region; nullptr: every region is hosted by this process
    int rank = 0;              // nccl: this process' region
    uint32_t stride = 0;       // records per segment in the buffers (the largest cap)
    // cap[kind][from * n + to]: records (header included) that cross the wire from `from` to `to` in an exchange of `kind`
    std::vector<uint32_t> cap[2];
    std::vector<epi_engine*> engines;      // local transport: all regions, by region index; nccl: the one engine
    std::vector<RegionBuffers> buffers;    // parallel to `engines`
    unsigned long long *d_sum = nullptr, *h_sum = nullptr;  // termination rule: all-reduce of exposed + infected + hospitalized
    // Peer transport (one process per region, NVLink / NVSwitch peer memory): every rank exposes a double-buffered receive area and
    // a flag word per source through CUDA IPC; a sender writes its segment straight into the destination's memory and then raises
    // its flag there; the receiver's arrive kernel waits for the flags.  NCCL only carried the IPC handles (and the termination
    // rule's all-reduce).  peer_ok == false: the ncclSend / ncclRecv path.
    bool peer_ok = false;
    TravelRecord* recv2 = nullptr;             // [2][n][stride] this rank's receive area (buffer = exchange number & 1)
    uint32_t* flags = nullptr;                 // [n] flags[q] = number of the last exchange whose segment from rank q is complete
    std::vector<void*> peer_mapped;            // what cudaIpcOpenMemHandle returned (closed by ~Comm)
    TravelRecord** d_peer_recv = nullptr;      // device array [n]: rank p's recv2 as seen from here
    uint32_t** d_peer_flags = nullptr;         // device array [n]: rank p's flags as seen from here
    uint32_t exchange_no = 0;                  // exchanges done
    ~Comm();
};

// exchange kind of `hour` under a travel plan, or -1 (mpi_transport.rs:60-76 + the migration window of citizen/mod.rs:460-462)
int exchange_kind_of(bool migration_enabled, bool commute_enabled, uint32_t start_migration_hour, uint32_t end_migration_hour, uint32_t hour);
// is `hour` one of the orchestrator's tick hours (orchestrator/src/ticks.rs:44-54)
bool is_tick_hour(bool migration_enabled, bool commute_enabled, uint32_t hour);

// What the hour loop needs from a region and from the transport: the real engines, or recording stand-ins (epi_multi_schedule_trace).
struct RegionOps {
    virtual ~RegionOps() {}
    virtual uint32_t next_decision_hour(uint32_t hour) = 0;
    virtual int enqueue_hours(uint32_t first_hour, uint32_t n) = 0;
    virtual int enqueue_hour(uint32_t hour) = 0;
    virtual int collect(std::vector<epi_counts>& rows) = 0;  // waits; the rows of every queued hour, exchange hours included
};
struct ExchangeOps {
    virtual ~ExchangeOps() {}
    virtual int exchange(uint32_t hour, int kind) = 0;
    // sum of `local` over all ranks (termination rule)
    virtual int all_reduce_sum(unsigned long long local, unsigned long long* total) = 0;
};
struct PlanInfo {
    bool migration_enabled = false, commute_enabled = false;
    uint32_t start_migration_hour = 0, end_migration_hour = 0;
};
// rows_out[region][n_hours]; *n_rows = hours executed
int run_multi_schedule(std::vector<RegionOps*>& regions, ExchangeOps& x, const PlanInfo& plan, uint32_t first_hour, uint32_t n_hours, bool terminate_when_clear,
                       epi_counts* rows_out, uint32_t* n_rows);

}  // namespace epi
