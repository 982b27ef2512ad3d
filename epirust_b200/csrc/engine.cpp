// C ABI of include/epi.h: engine lifecycle, the per-hour step, interventions sweeps, state import/export, measurement.
#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>

#include "kernels.h"
#include "philox.cuh"

using namespace epi;

namespace epi {
static std::mutex g_err_mutex;
static std::string g_err;
void set_global_error(const std::string& msg) {
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_err = msg;
}
int engine_fail(const epi_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg;
    else set_global_error(msg);
    return code;
}
}  // namespace epi

#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t _err = (call);                                                                                 \
        if (_err != cudaSuccess)                                                                                   \
            return engine_fail(e, EPI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_err));              \
    } while (0)

enum KernelKind { KK_HOUR = 0, KK_COMMIT = 1, KK_SCAN = 2, KK_SLEEP = 3, KK_SWEEP = 4, KK_SPARE = 5, KK_TRAVEL = 6, KK_MISC = 7 };

namespace {

template <class T>
cudaError_t dev_alloc(epi_engine* e, T** p, size_t count) {
    cudaError_t r = cudaMalloc((void**)p, count * sizeof(T));
    if (r == cudaSuccess) e->device_bytes += count * sizeof(T);
    return r;
}

// bracket one launch with events when per-kernel timing is on
struct Timed {
    epi_engine* e;
    int kind;
    cudaEvent_t a = nullptr, b = nullptr;
    // hod: the hour of day of an hour / commit launch (per-hour-of-day sums, epi_get_hour_times), or -1
    Timed(epi_engine* e_, int kind_, int hod = -1) : e(e_), kind(kind_ | ((hod + 1) << 8)) {
        e->launches++;
        e->kernel_launches[kind_]++;
        if (e->timing) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, e->stream);
        }
    }
    ~Timed() {
        if (e->timing) {
            cudaEventRecord(b, e->stream);
            e->pending_events.push_back({kind, {a, b}});
        }
    }
};

void drain_events(epi_engine* e) {
    for (auto& pe : e->pending_events) {
        float ms = 0.f;
        cudaEventSynchronize(pe.second.second);
        cudaEventElapsedTime(&ms, pe.second.first, pe.second.second);
        e->kernel_ms[pe.first & 0xFF] += ms;
        const int hod = (pe.first >> 8) - 1, kk = pe.first & 0xFF;
        if (hod >= 0 && (kk == KK_HOUR || kk == KK_COMMIT)) {
            e->hour_ms[hod * 2 + (kk == KK_COMMIT)] += ms;
            e->hour_launches[hod * 2 + (kk == KK_COMMIT)]++;
        }
        cudaEventDestroy(pe.second.first);
        cudaEventDestroy(pe.second.second);
    }
    e->pending_events.clear();
}

size_t n_cells(const epi_engine* e) { return (size_t)e->geo.pitch * e->geo.rows; }  // logical cells (host API)
// grid allocation with its zero padding (layout.h); claim[] has one word per byte of it
size_t grid_alloc_bytes(const epi_engine* e) { return e->P.grid_bytes(); }

int rebuild_grid(epi_engine* e) {
    CU(cudaMemsetAsync(e->grid_alloc, 0, grid_alloc_bytes(e), e->stream));
    CU(cudaMemsetAsync(e->d_misc, 0, 2 * sizeof(uint32_t), e->stream));
    {
        Timed t(e, KK_MISC);
        launch_build_grid(e->P, e->D, e->d_misc + 1, e->stream);
    }
    uint32_t collisions = 0;
    CU(cudaMemcpyAsync(&collisions, e->d_misc + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (collisions) return engine_fail(e, EPI_ERR_STATE, "two agents on one cell (" + std::to_string(collisions) + " collisions)");
    e->claim_dirty = true;
    // running Counts totals: absolute recount into copy 0
    CU(cudaMemsetAsync(e->D.tot, 0, (size_t)TOT_COPIES * 8 * sizeof(uint32_t), e->stream));
    {
        Timed t(e, KK_MISC);
        launch_recount(e->P, e->D, e->stream);
    }
    e->have_last_row = false;
    return EPI_OK;
}

void drop_graph(epi_engine* e) {
    if (e->day_graph) {
        cudaGraphExecDestroy(e->day_graph);
        e->day_graph = nullptr;
    }
    for (auto& g : e->segment_graphs)
        if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    e->segment_graphs.clear();
}

// which tile order serves hour-of-day h: 0 = office tiles, 1 = house tiles, -1 = none (tiles.cu)
int tile_order_of(uint32_t h) {
    if ((h >= 9u && h <= 11u) || (h >= 13u && h <= 15u)) return 0;
    if (h >= 18u && h <= 22u) return 1;
    return -1;
}

TileGeom make_tile_geom(const epi_engine* e, int order) {
    auto env_int = [](const char* name, int dflt) { const char* v = std::getenv(name); return v && *v ? std::max(1, std::atoi(v)) : dflt; };
    TileGeom g{};
    if (order == 0) {
        g.ox = e->P.work.sx; g.oy = e->P.work.sy; g.unit = 10; g.units_x = e->P.office_nx; g.units_y = e->P.office_ny;
        g.tile_units = env_int("EPI_TILE_OFFICES", 8);
        g.cls = 1;  // RC_OFFICE
        g.cap = (uint32_t)g.tile_units * 100u;
        g.threads = (uint32_t)env_int("EPI_TILE_OFFICE_THREADS", 256);
    } else {
        g.ox = e->P.housing().sx; g.oy = e->P.housing().sy; g.unit = 2; g.units_x = e->P.house_nx; g.units_y = e->P.house_ny;
        g.tile_units = env_int("EPI_TILE_HOUSES", 64);
        g.cls = 0;  // RC_HOME
        g.cap = (uint32_t)g.tile_units * 4u;
        g.threads = (uint32_t)env_int("EPI_TILE_HOUSE_THREADS", 192);
    }
    g.threads = std::min(256u, (g.threads + 31u) / 32u * 32u);
    g.chunks = (g.units_x + g.tile_units - 1) / g.tile_units;
    g.n_tiles = (uint32_t)g.chunks * (uint32_t)g.units_y;
    g.sp = 16 + ((g.unit * g.tile_units + 15 + 15) & ~15) + 16;
    return g;
}

// (re)build the two tile orders from the agents' current home / work / work-status words
int rebuild_tiles(epi_engine* e) {
    e->tiles_ready = false;
    if (!e->tiles_enabled) return EPI_OK;
    drop_graph(e);
    for (int o = 0; o < 2; ++o) {
        cudaError_t r = build_tile_order(e->P, e->D, e->tile_geom[o], e->tile_ptrs[o], e->tile_keys_a, e->tile_keys_b, e->tile_ids, e->tile_temp, e->tile_temp_bytes, e->stream);
        if (r != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("tile order: ") + cudaGetErrorString(r));
        CU(cudaMemsetAsync(e->tile_ptrs[o].dirty, 0, (size_t)e->tile_geom[o].n_tiles * sizeof(uint32_t), e->stream));
        e->launches += 2;
    }
    CU(cudaMemsetAsync(e->d_tile_misc, 0, 4 * sizeof(uint32_t), e->stream));
    uint32_t first_generic[2] = {0, 0};
    for (int o = 0; o < 2; ++o) CU(cudaMemcpyAsync(&first_generic[o], e->tile_ptrs[o].start + 2 * (size_t)e->tile_geom[o].n_tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (int o = 0; o < 2; ++o) e->tile_generic_bound[o] = e->P.n - first_generic[o];
    e->tiles_ready = true;
    return EPI_OK;
}

// Off unless asked for (EPI_TILES=1 / epi_set_tiles): bit-exact, but measured slower than the id-order kernels on B200 -- the
// tile kernel saves ~45 % of the DRAM bytes of a plain hour and the commit pass of its members, and pays ~40 % more issue slots
// per agent-hour (two window sources, warp-level settlement) in a kernel that was issue-bound to begin with (profiles/r02_tile_*).
int setup_tiles(epi_engine* e, bool force) {
    const char* v = std::getenv("EPI_TILES");
    e->tiles_enabled = !e->multi && (force || (v && v[0] == '1'));
    if (!e->tiles_enabled) return EPI_OK;
    if (e->tile_temp) return EPI_OK;  // already allocated
    const size_t n = e->P.n;
    bool ok = true;
    for (int o = 0; o < 2; ++o) {
        e->tile_geom[o] = make_tile_geom(e, o);
        if (e->tile_geom[o].n_tiles == 0 || tile_shared_bytes(e->tile_geom[o]) > 200u * 1024u) { e->tiles_enabled = false; return EPI_OK; }
        ok &= dev_alloc(e, &e->tile_ptrs[o].perm, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->tile_ptrs[o].start, 2 * (size_t)e->tile_geom[o].n_tiles + 2) == cudaSuccess;
        ok &= dev_alloc(e, &e->tile_ptrs[o].dirty, (size_t)e->tile_geom[o].n_tiles) == cudaSuccess;
    }
    ok &= dev_alloc(e, &e->tile_keys_a, n) == cudaSuccess;
    ok &= dev_alloc(e, &e->tile_keys_b, n) == cudaSuccess;
    ok &= dev_alloc(e, &e->tile_ids, n) == cudaSuccess;
    ok &= dev_alloc(e, &e->d_tile_misc, 4) == cudaSuccess;
    e->tile_temp_bytes = tile_sort_temp_bytes(e->P.n);
    uint8_t* temp = nullptr;
    ok &= dev_alloc(e, &temp, e->tile_temp_bytes) == cudaSuccess;
    e->tile_temp = temp;
    if (!ok) return engine_fail(e, EPI_ERR_CUDA, std::string("device allocation failed (tile orders): ") + cudaGetErrorString(cudaGetLastError()));
    for (int o = 0; o < 2; ++o) { e->tile_ptrs[o].n_housing = e->d_tile_misc; e->tile_ptrs[o].misc = e->d_tile_misc; }
    cudaError_t r = tiles_configure(e->tile_geom[0], e->tile_geom[1]);
    if (r != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("tile kernels: ") + cudaGetErrorString(r));
    return EPI_OK;
}

// enqueue one simulated hour (kernels only).  `inject`: draws table already on device.
// prev_was_sleep: the previous hour of this same call was a sleep hour that already ran k_sleep -> nothing to do.
int enqueue_hour(epi_engine* e, uint32_t hour, uint32_t hour_offset, bool inject, bool skip_sleep) {
    const uint32_t h = hour % 24u;
    if (h >= 1 && h <= 6) {
        if (skip_sleep) return EPI_OK;
        Timed t(e, KK_SLEEP);
        launch_sleep(e->P, e->D, hour_offset, e->stream);
        return EPI_OK;
    }
    if (h == 0) {
        cudaMemsetAsync(e->D.hosp_first, 0xFF, sizeof(uint32_t), e->stream);
        Timed t(e, KK_SCAN);
        launch_hospital_scan(e->P, e->D, e->stream);
    }
    const int order = (e->tiles_ready && !inject) ? tile_order_of(h) : -1;
    if (order >= 0) {
        // a plain movement hour: generic segment, then one CTA per office / house tile (tiles.cu); the commit pass only sees the
        // proposals of the agents that took the global path
        if (order == 1 && (h == 18u || hour_offset == 0)) {  // evening phase starts (or this batch starts inside it): who may step into any house?
            Timed t(e, KK_MISC);
            launch_count_housing(e->P, e->D, e->d_tile_misc, e->stream);
        }
        {
            Timed t(e, KK_HOUR, (int)h);
            e->launches += launch_hour_tiles(e->P, e->D, e->tile_geom[order], e->tile_ptrs[order], e->tile_generic_bound[order], hour_offset, e->stream) - 1u;
        }
        {
            Timed t(e, KK_COMMIT, (int)h);
            launch_commit(e->P, e->D, hour_offset, true, true, e->stream);
        }
        e->tile_hours++;
        return EPI_OK;
    }
    {
        Timed t(e, KK_HOUR, (int)h);
        launch_hour(e->P, e->D, h, hour_offset, inject, e->stream);
    }
    {
        Timed t(e, KK_COMMIT, (int)h);
        // the tile kernels rely on prop[] being all zero when their hour starts
        launch_commit(e->P, e->D, hour_offset, false, e->tiles_ready && tile_order_of((h + 1u) % 24u) >= 0, e->stream);
    }
    return EPI_OK;
}

// claim stamps: stamp = hour - epoch_base + 1 must stay below 2^(32 - id_bits) (k_hour / k_commit shift it left by id_bits).
// Called for every piece of work that is launched under one Clock value (a single hour, a graph of <= 24 hours): when the
// piece does not fit the current epoch the claim array is zeroed and the epoch restarts at the piece's first hour, so a
// chunk of any length (epi_run_hours: up to RING_ROWS hours, 256 stamps at 10 M agents) is split at the stamp limit.
int ensure_epoch(epi_engine* e, uint32_t first_hour, uint32_t last_hour) {
    const uint64_t limit = 1ull << (32 - e->P.id_bits);
    if ((uint64_t)(last_hour - first_hour) + 1ull >= limit)
        return engine_fail(e, EPI_ERR_STATE, "claim stamps: " + std::to_string(last_hour - first_hour + 1) + " hours under one clock do not fit " + std::to_string(32 - e->P.id_bits) + " stamp bits");
    const bool fits = !e->claim_dirty && first_hour >= e->epoch_base && (uint64_t)(last_hour - e->epoch_base) + 1ull < limit;
    if (!fits) {
        CU(cudaMemsetAsync(e->D.claim, 0, grid_alloc_bytes(e) * sizeof(uint32_t), e->stream));
        e->epoch_base = first_hour;
        e->claim_dirty = false;
        e->epoch_resets++;
    }
    return EPI_OK;
}

int build_day_graph(epi_engine* e) {
    // capture hours with h%24 = 1..23,0 (offsets 0..23) once; replayed for every aligned day
    cudaGraph_t graph = nullptr;
    const bool timing = e->timing;
    e->timing = false;
    const uint64_t launches0 = e->launches, tile_hours0 = e->tile_hours;
    uint64_t kl0[EPI_N_KERNEL_KINDS];
    memcpy(kl0, e->kernel_launches, sizeof(kl0));
    CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    bool sleep_done = false;
    for (uint32_t off = 0; off < 24; ++off) {
        const uint32_t hour = 1 + off;  // representative: only hour % 24 matters for the structure
        const uint32_t h = hour % 24u;
        const bool is_sleep = h >= 1 && h <= 6;
        enqueue_hour(e, hour, off, false, is_sleep && sleep_done);
        if (is_sleep) sleep_done = true;
    }
    cudaError_t r = cudaStreamEndCapture(e->stream, &graph);
    e->timing = timing;
    e->day_graph_launches = (uint32_t)(e->launches - launches0);
    e->launches = launches0;
    e->tile_hours = tile_hours0;
    memcpy(e->kernel_launches, kl0, sizeof(kl0));
    if (r != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(r));
    r = cudaGraphInstantiate(&e->day_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (r != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(r));
    return EPI_OK;
}

void row_to_counts(const uint32_t* row, uint32_t hour, epi_counts* out) {
    out->hour = hour;
    out->susceptible = row[0]; out->exposed = row[1]; out->infected = row[2];
    out->hospitalized = row[3]; out->recovered = row[4]; out->deceased = row[5];
}

// ---- queued hours (multi-region days without intermediate host waits) ---------------------------------------------------
// Capture hours [first, first + n) (n <= 24, no exchange inside) once per (hour of day, n); the device-resident Clock carries
// the absolute hour and the ring base, so the same graph serves every day.
int segment_graph(epi_engine* e, uint32_t first_hour, uint32_t n, epi_engine::SegmentGraph** out) {
    const uint32_t key = (first_hour % 24u) * 32u + n;
    for (auto& g : e->segment_graphs)
        if (g.first == key) { *out = &g.second; return EPI_OK; }
    epi_engine::SegmentGraph sg;
    cudaGraph_t graph = nullptr;
    const uint64_t launches0 = e->launches, tile_hours0 = e->tile_hours;
    uint64_t kl0[EPI_N_KERNEL_KINDS];
    memcpy(kl0, e->kernel_launches, sizeof(kl0));
    CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    bool sleep_done = false;
    for (uint32_t off = 0; off < n; ++off) {
        const uint32_t h = (first_hour + off) % 24u;
        const bool is_sleep = h >= 1 && h <= 6;
        enqueue_hour(e, first_hour + off, off, false, is_sleep && sleep_done);
        sleep_done = is_sleep;
    }
    cudaError_t r = cudaStreamEndCapture(e->stream, &graph);
    sg.launches = (uint32_t)(e->launches - launches0);
    sg.n_sleep = (uint32_t)(e->kernel_launches[KK_SLEEP] - kl0[KK_SLEEP]);
    sg.n_active = (uint32_t)(e->kernel_launches[KK_HOUR] - kl0[KK_HOUR]);
    sg.n_scan = (uint32_t)(e->kernel_launches[KK_SCAN] - kl0[KK_SCAN]);
    sg.n_tile = (uint32_t)(e->tile_hours - tile_hours0);
    e->launches = launches0;
    e->tile_hours = tile_hours0;
    memcpy(e->kernel_launches, kl0, sizeof(kl0));
    if (r != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(r));
    r = cudaGraphInstantiate(&sg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (r != cudaSuccess) return engine_fail(e, EPI_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(r));
    e->segment_graphs.emplace_back(key, sg);
    *out = &e->segment_graphs.back().second;
    return EPI_OK;
}

// append hours [first, first + n) to the queue; exchange_hour: the single hour of an exchange (its kernels only)
int queue_hours(epi_engine* e, uint32_t first_hour, uint32_t n, bool exchange_hour) {
    if (n == 0) return EPI_OK;
    if (!e->pend_kind.empty() && first_hour != e->pend_first + (uint32_t)e->pend_kind.size())
        return engine_fail(e, EPI_ERR_STATE, "queued hours must be consecutive");
    if (e->pend_kind.size() + n > RING_ROWS) return engine_fail(e, EPI_ERR_STATE, "too many queued hours: call epi_collect_hours");
    if (e->pend_kind.empty()) e->pend_first = first_hour;
    const uint32_t row0 = first_hour - e->pend_first;
    CU(cudaMemsetAsync(e->D.counts + (size_t)row0 * 8, 0, (size_t)n * 8 * sizeof(uint32_t), e->stream));
    int rc = EPI_OK;
    uint32_t off = 0;
    while (off < n) {
        const uint32_t hour = first_hour + off;
        const uint32_t len = std::min(n - off, 24u);
        rc = ensure_epoch(e, hour, hour + len - 1);
        if (rc) return rc;
        {
            Timed t(e, KK_MISC);
            launch_set_clock(e->d_clock, Clock{hour, e->epoch_base, e->pend_first, 0}, e->stream);
        }
        if (!exchange_hour && !e->timing && e->graphs_enabled) {
            epi_engine::SegmentGraph* sg = nullptr;
            rc = segment_graph(e, hour, len, &sg);
            if (rc) return rc;
            CU(cudaGraphLaunch(sg->exec, e->stream));
            e->launches += sg->launches;
            e->kernel_launches[KK_SLEEP] += sg->n_sleep; e->kernel_launches[KK_HOUR] += sg->n_active; e->kernel_launches[KK_COMMIT] += sg->n_active;
            e->kernel_launches[KK_SCAN] += sg->n_scan;
            e->tile_hours += sg->n_tile;
        }
        bool sleep_done = false;
        for (uint32_t k = 0; k < len; ++k) {
            const uint32_t h = (hour + k) % 24u;
            const bool is_sleep = h >= 1 && h <= 6;
            if (exchange_hour || e->timing || !e->graphs_enabled) {
                rc = enqueue_hour(e, hour + k, k, false, is_sleep && sleep_done);
                if (rc) return rc;
            }
            e->pend_kind.push_back(exchange_hour ? 2 : (is_sleep && sleep_done ? 0 : 1));
            e->pend_population.push_back(e->population);
            sleep_done = is_sleep;
        }
        off += len;
    }
    CU(cudaGetLastError());
    return EPI_OK;
}

// wait for the queued hours and turn their ring rows into Counts (the per-row checks of run_chunk); exchange rows are left out
int collect_hours(epi_engine* e, std::vector<epi_counts>& rows) {
    rows.clear();
    const uint32_t n = (uint32_t)e->pend_kind.size();
    if (n == 0) return EPI_OK;
    CU(cudaMemcpyAsync(e->h_counts, e->D.counts, (size_t)n * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    e->h_small[63] = 0;
    if (e->tiles_ready) CU(cudaMemcpyAsync(e->h_small + 63, e->d_tile_misc + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (e->timing) drain_events(e);
    if (e->h_small[63]) return engine_fail(e, EPI_ERR_STATE, "tile kernels: an agent's home / work word does not match its tile (stale tile order)");
    const std::vector<uint8_t> kind = e->pend_kind;
    const std::vector<uint32_t> population = e->pend_population;
    const uint32_t first_hour = e->pend_first;
    e->pend_kind.clear();
    e->pend_population.clear();
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t* row = e->h_counts + (size_t)k * 8;
        uint32_t pop = population[k];
        if (kind[k] == 2) continue;  // an exchange hour driven by hand: its row comes from epi_finish_hour
        if (kind[k] == 3) {
            // an exchange hour: the device wrote its row after the arrivals were placed (k_travel_place), with the region's new
            // population and the exchange's error flags next to it
            const int rc = travel_status(e, row[7], row[6]);
            if (rc) return rc;
            e->pack_unsettled = e->unpack_unsettled = false;
            pop = e->population;
            row_to_counts(row, first_hour + k, &e->last_counts);
            e->have_last_row = true;
        } else if (kind[k] == 1) {
            const epi_counts prev = e->last_counts;
            row_to_counts(row, first_hour + k, &e->last_counts);
            const uint32_t h = (first_hour + k) % 24u;
            if (h >= 1 && h <= 6 && e->have_last_row) {
                const epi_counts& c = e->last_counts;
                if (c.susceptible != prev.susceptible || c.exposed != prev.exposed || c.infected != prev.infected ||
                    c.hospitalized != prev.hospitalized || c.recovered != prev.recovered || c.deceased != prev.deceased)
                    return engine_fail(e, EPI_ERR_STATE, "incrementally tracked Counts diverged from the recount at hour " + std::to_string(first_hour + k));
            }
            e->have_last_row = true;
        } else e->last_counts.hour = first_hour + k;
        if (e->multi) pop = e->population;  // as of the last exchange row seen
        const epi_counts& c = e->last_counts;
        const uint64_t total = (uint64_t)c.susceptible + c.exposed + c.infected + c.hospitalized + c.recovered + c.deceased;
        if (total != pop)
            return engine_fail(e, EPI_ERR_STATE, "counts total " + std::to_string(total) + " != population " + std::to_string(pop) + " at hour " + std::to_string(first_hour + k));
        rows.push_back(c);
    }
    return EPI_OK;
}

// run hours [first, first+n) (n <= RING_ROWS), rows to out
int run_chunk(epi_engine* e, uint32_t first_hour, uint32_t n, bool inject, epi_counts* out) {
    if (!e->pend_kind.empty()) {
        if (e->pend_kind.size() == 1 && e->pend_kind[0] == 2) { e->pend_kind.clear(); e->pend_population.clear(); }  // an exchange hour settled by epi_finish_hour
        else return engine_fail(e, EPI_ERR_STATE, "hours are queued: call epi_collect_hours first");
    }
    CU(cudaMemsetAsync(e->D.counts, 0, (size_t)n * 8 * sizeof(uint32_t), e->stream));
    int rc = EPI_OK;
    std::vector<uint8_t> ran(n, 0);  // hour produced its own counts row
    uint32_t off = 0;
    bool sleep_done = false;  // a k_sleep already ran in the current run of consecutive sleep hours
    while (off < n) {
        const uint32_t hour = first_hour + off;
        const bool aligned_day = !inject && !e->timing && e->graphs_enabled && hour % 24u == 1u && n - off >= 24u;
        rc = ensure_epoch(e, hour, aligned_day ? hour + 23u : hour);
        if (rc) return rc;
        {
            Timed t(e, KK_MISC);
            launch_set_clock(e->d_clock, Clock{hour, e->epoch_base, first_hour, 0}, e->stream);
        }
        if (aligned_day) {
            if (!e->day_graph) {
                rc = build_day_graph(e);
                if (rc) return rc;
            }
            CU(cudaGraphLaunch(e->day_graph, e->stream));
            e->launches += e->day_graph_launches;
            e->kernel_launches[KK_SLEEP] += 1; e->kernel_launches[KK_HOUR] += 18; e->kernel_launches[KK_COMMIT] += 18; e->kernel_launches[KK_SCAN] += 1;
            if (e->tiles_ready) e->tile_hours += 11;
            for (uint32_t k = 0; k < 24; ++k) { const uint32_t h = (hour + k) % 24u; ran[off + k] = !(h >= 2 && h <= 6); }
            off += 24;
            sleep_done = false;
            continue;
        }
        const uint32_t h = hour % 24u;
        const bool is_sleep = h >= 1 && h <= 6;
        rc = enqueue_hour(e, hour, 0, inject, is_sleep && sleep_done);
        if (rc) return rc;
        ran[off] = !(is_sleep && sleep_done);
        sleep_done = is_sleep;
        ++off;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(e->h_counts, e->D.counts, (size_t)n * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    e->h_small[63] = 0;
    if (e->tiles_ready) CU(cudaMemcpyAsync(e->h_small + 63, e->d_tile_misc + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (e->timing) drain_events(e);
    if (e->h_small[63]) return engine_fail(e, EPI_ERR_STATE, "tile kernels: an agent's home / work word does not match its tile (stale tile order)");
    for (uint32_t k = 0; k < n; ++k) {
        if (ran[k]) {
            const epi_counts prev = e->last_counts;
            row_to_counts(e->h_counts + (size_t)k * 8, first_hour + k, &e->last_counts);
            const uint32_t h = (first_hour + k) % 24u;
            if (h >= 1 && h <= 6 && e->have_last_row) {
                // k_sleep recounts every agent; the previous row came from the incrementally tracked totals: they must agree
                const epi_counts& c = e->last_counts;
                if (c.susceptible != prev.susceptible || c.exposed != prev.exposed || c.infected != prev.infected ||
                    c.hospitalized != prev.hospitalized || c.recovered != prev.recovered || c.deceased != prev.deceased)
                    return engine_fail(e, EPI_ERR_STATE, "incrementally tracked Counts diverged from the recount at hour " + std::to_string(first_hour + k));
            }
            e->have_last_row = true;
        } else e->last_counts.hour = first_hour + k;  // sleep hour: nothing observable changed (citizen/mod.rs:244-248)
        out[k] = e->last_counts;
        const uint64_t total = (uint64_t)out[k].susceptible + out[k].exposed + out[k].infected + out[k].hospitalized + out[k].recovered + out[k].deceased;
        if (total != e->population)  // allocation_map.rs:128 assert_eq!(csv_record.total(), current_population)
            return engine_fail(e, EPI_ERR_STATE, "counts total " + std::to_string(total) + " != population " + std::to_string(e->population) + " at hour " + std::to_string(first_hour + k));
    }
    return EPI_OK;
}

int upload_agents(epi_engine* e, const HostAgents& a) {
    const size_t nb = (size_t)e->P.n * sizeof(uint32_t);
    CU(cudaMemcpyAsync(e->D.cell, a.cell.data(), nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.st, a.st.data(), nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.t0, a.t0.data(), nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.home, a.home.data(), nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.work, a.work.data(), nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.wsa, a.wsa.data(), nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.reg, a.reg.data(), nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(e->D.prop, 0, nb, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return EPI_OK;
}

int snapshot_initial(epi_engine* e) {
    const size_t nb = (size_t)e->P.n * sizeof(uint32_t);
    CU(cudaMemcpyAsync(e->i_cell, e->D.cell, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->i_st, e->D.st, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->i_t0, e->D.t0, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->i_home, e->D.home, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->i_work, e->D.work, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->i_wsa, e->D.wsa, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->i_reg, e->D.reg, nb, cudaMemcpyDeviceToDevice, e->stream));
    return EPI_OK;
}

// tie order of the occupancy heaps (grid.rs:67-73): the greatest (start.x, start.y) first -- see house_rank() in travel.cu
uint32_t house_rank_of_index(const Geometry& g, uint32_t idx) {
    const int hx = (int)(idx % (uint32_t)g.house_nx), hy = (int)(idx / (uint32_t)g.house_nx);
    return (uint32_t)((g.house_nx - 1 - hx) * g.house_ny + (g.house_ny - 1 - hy));
}
uint32_t office_rank_of_index(const Geometry& g, uint32_t idx) {
    const int ox = (int)(idx % (uint32_t)g.office_nx), oy = (int)(idx / (uint32_t)g.office_nx);
    return (uint32_t)((g.office_nx - 1 - ox) * g.office_ny + (g.office_ny - 1 - oy));
}

template <class T>
int travel_alloc(epi_engine* e, T** p, size_t count) {
    CU(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    e->device_bytes += count * sizeof(T);
    e->travel_allocs.push_back((void*)*p);
    return EPI_OK;
}

// device-side bookkeeping of the traveller exchange (multi-region engines)
int alloc_travel(epi_engine* e, const epi_travel_plan* plan) {
    const uint32_t R = (uint32_t)plan->n_regions, self = (uint32_t)e->P.region;
    uint64_t most = 0;
    for (int m = 0; m < 2; ++m) {
        const uint32_t* mat = m == 0 ? (plan->migration_enabled ? plan->migration : nullptr) : (plan->commute_enabled ? plan->commute : nullptr);
        if (!mat) continue;
        uint64_t row = 0, col = 0;
        for (uint32_t k = 0; k < R; ++k) { row += mat[(size_t)self * R + k]; col += mat[(size_t)k * R + self]; }
        most = std::max(most, std::max(row, col));
    }
    epi::TravelPtrs& T = e->T;
    T.n_regions = (int)R;
    T.list_cap = (uint32_t)std::min<uint64_t>(2 * most + 65536, e->P.n);
    uint32_t table = 1024;
    while (table < 4u * T.list_cap) table <<= 1;
    T.table_mask = table - 1u;
    const size_t nbh = (e->geo.n_houses + 255) / 256, nbo = (e->geo.n_offices + 255) / 256;
    int rc;
    uint32_t* row = nullptr;
    if ((rc = travel_alloc(e, &T.occ_house, e->geo.n_houses))) return rc;
    if ((rc = travel_alloc(e, &T.occ_office, e->geo.n_offices))) return rc;
    if ((rc = travel_alloc(e, &T.free_stack, e->P.n))) return rc;
    if ((rc = travel_alloc(e, &T.tv, 1))) return rc;
    if ((rc = travel_alloc(e, &row, R))) return rc;
    T.plan_row = row;
    if ((rc = travel_alloc(e, &T.list_slot, T.list_cap))) return rc;
    if ((rc = travel_alloc(e, &T.list_dest, T.list_cap))) return rc;
    if ((rc = travel_alloc(e, &T.list_pos, T.list_cap))) return rc;
    if ((rc = travel_alloc(e, &T.arrivals, T.list_cap))) return rc;
    if ((rc = travel_alloc(e, &T.arr_widx, T.list_cap))) return rc;
    if ((rc = travel_alloc(e, &T.arr_house, T.list_cap))) return rc;
    if ((rc = travel_alloc(e, &T.arr_office, T.list_cap))) return rc;
    if ((rc = travel_alloc(e, &T.placed, T.list_cap))) return rc;
    if ((rc = travel_alloc(e, &T.table_keys, table))) return rc;
    if ((rc = travel_alloc(e, &T.table_vals, table))) return rc;
    if ((rc = travel_alloc(e, &T.bh_house, nbh * HOUSE_CAP))) return rc;
    if ((rc = travel_alloc(e, &T.pref_house, nbh * HOUSE_CAP))) return rc;
    if ((rc = travel_alloc(e, &T.bh_office, nbo * OFFICE_CAP))) return rc;
    if ((rc = travel_alloc(e, &T.pref_office, nbo * OFFICE_CAP))) return rc;
    if ((rc = travel_alloc(e, &T.plan_house, 1))) return rc;
    if ((rc = travel_alloc(e, &T.plan_office, 1))) return rc;
    if ((rc = travel_alloc(e, &e->t_block_counts, (size_t)(e->P.n + 1023) / 1024 + 1 + ((size_t)(e->P.n + 1023) / 1024) * 32))) return rc;  // block counts | warp ballots
    if ((rc = travel_alloc(e, &T.chunk_dest, ((size_t)T.list_cap / 256 + 2) * R))) return rc;
    {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
        const size_t grid_most = (size_t)std::max(1, sms) * 8;  // k_travel_arrive never runs more blocks than this
        if ((rc = travel_alloc(e, &T.hslice, 2 * (size_t)OFFICE_CAP * grid_most))) return rc;
        if ((rc = travel_alloc(e, &T.tot_house, (size_t)TOT_COPIES * HOUSE_CAP))) return rc;
        if ((rc = travel_alloc(e, &T.tot_office, (size_t)TOT_COPIES * OFFICE_CAP))) return rc;
        if ((rc = travel_alloc(e, &T.foreign, ((size_t)(e->P.n + 1023) / 1024) * 32 + 32))) return rc;
    }
    CU(cudaMallocHost((void**)&e->h_tv, sizeof(epi::TravelVars)));
    CU(cudaMemsetAsync(T.tv, 0, sizeof(epi::TravelVars), e->stream));
    // the placement rounds' hash table starts empty and every round leaves it empty (k_travel_place)
    CU(cudaMemsetAsync(T.table_keys, 0, (size_t)table * sizeof(uint32_t), e->stream));
    CU(cudaMemsetAsync(T.table_vals, 0xFF, (size_t)table * sizeof(uint32_t), e->stream));
    CU(cudaMemcpyAsync(row, e->migration_row.data(), R * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return EPI_OK;
}

// bookkeeping of the exchange back to its initial state (set_start_locations_and_occupancies, grid.rs:125-155)
int reset_travel_state(epi_engine* e) {
    e->population = e->cfg.number_of_agents;
    if (!e->multi) return EPI_OK;
    e->pack_unsettled = e->unpack_unsettled = false;
    const uint32_t top = (uint32_t)e->free_stack0.size();
    CU(cudaMemsetAsync(e->T.tv, 0, sizeof(epi::TravelVars), e->stream));
    CU(cudaMemcpyAsync(&e->T.tv->free_top, &top, sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(&e->T.tv->population, &e->population, sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->T.free_stack, e->free_stack0.data(), e->free_stack0.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->T.occ_house, e->occ_house0.data(), e->occ_house0.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->T.occ_office, e->occ_office0.data(), e->occ_office0.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return EPI_OK;
}

// the running level histograms of the occupancy heaps and the foreign-slot marks, from the agents and occupancies now on the device
int recount_travel(epi_engine* e) {
    if (!e->multi) return EPI_OK;
    CU(cudaMemsetAsync(e->T.tot_house, 0, (size_t)TOT_COPIES * HOUSE_CAP * sizeof(uint32_t), e->stream));
    CU(cudaMemsetAsync(e->T.tot_office, 0, (size_t)TOT_COPIES * OFFICE_CAP * sizeof(uint32_t), e->stream));
    CU(cudaMemsetAsync(e->T.foreign, 0, (((size_t)(e->P.n + 1023) / 1024) * 32 + 32) * sizeof(uint32_t), e->stream));
    launch_travel_recount(e->P, e->D, e->T, e->geo.n_houses, e->geo.n_offices, e->stream);
    e->launches += 3;
    CU(cudaGetLastError());
    return EPI_OK;
}

void initial_counts(epi_engine* e) {
    const epi_config& c = e->cfg;
    const uint32_t total = c.exposed + c.infected_mild_asymptomatic + c.infected_mild_symptomatic + c.infected_severe;
    e->last_counts = epi_counts{0, c.number_of_agents - total, c.exposed, total - c.exposed, 0, 0, 0};
}

}  // namespace

extern "C" {

const char* epi_version(void) { return "epirust_b200 0.2 (sm_100a)"; }

int epi_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

const char* epi_last_error(const epi_engine* e) {
    if (e) return e->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mutex);
    static thread_local std::string copy;
    copy = g_err;
    return copy.c_str();
}

int epi_create(const epi_config* cfg, uint64_t seed, int device, epi_engine** out) { return epi_create_multi(cfg, seed, device, 0, nullptr, 0, out); }

int epi_create_region(const epi_config* cfg, uint64_t seed, int device, int region, epi_engine** out) {
    return epi_create_multi(cfg, seed, device, region, nullptr, 0, out);
}

int epi_create_multi(const epi_config* cfg_in, uint64_t seed, int device, int region, const epi_travel_plan* plan, uint32_t extra_capacity, epi_engine** out) {
    if (!cfg_in || !out) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    *out = nullptr;
    // Population::Csv: the population file decides the number of agents (Grid::read_population, grid.rs:194-231)
    PopulationRecords records;
    epi_config resolved;
    try {
        resolved = resolve_population(*cfg_in, records);
    } catch (const std::exception& ex) {
        return engine_fail(nullptr, EPI_ERR_IO, ex.what());
    }
    const epi_config* cfg = &resolved;
    const bool from_csv = resolved.population_csv_file[0] != 0;
    const std::string bad = validate_config(*cfg);
    if (!bad.empty()) return engine_fail(nullptr, EPI_ERR_CONFIG, bad);
    if (region < 0 || region > 254) return engine_fail(nullptr, EPI_ERR_ARG, "region must be in 0..254");
    if (plan) {
        if (plan->n_regions < 1 || plan->n_regions > 255 || region >= plan->n_regions) return engine_fail(nullptr, EPI_ERR_ARG, "travel plan: bad n_regions / region");
        if ((plan->migration_enabled && !plan->migration) || (plan->commute_enabled && !plan->commute)) return engine_fail(nullptr, EPI_ERR_ARG, "travel plan: enabled matrix is null");
    }
    if ((uint64_t)cfg->number_of_agents + extra_capacity > (1u << 27)) return engine_fail(nullptr, EPI_ERR_CONFIG, "agent slots must be <= 2^27");
    int n_dev = 0;
    cudaError_t cr = cudaGetDeviceCount(&n_dev);
    if (cr != cudaSuccess || n_dev == 0)
        return engine_fail(nullptr, EPI_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(cr));
    if (device < 0 || device >= n_dev) return engine_fail(nullptr, EPI_ERR_ARG, "device index out of range");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10)
        return engine_fail(nullptr, EPI_ERR_CUDA, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                                      "; this library is built for sm_100a only");
    epi_engine* e = new epi_engine(*cfg);
    e->seed = seed;
    e->device = device;
    auto fail = [&](int rc) {
        set_global_error(e->err);
        epi_destroy(e);
        return rc;
    };
    try {
        e->geo = make_geometry(cfg->grid_size, cfg->number_of_agents, cfg->hospital_beds_percentage);
        e->P = make_params(*cfg, e->geo, seed, region);
        HostAgents agents;
        build_population(*cfg, e->geo, seed, region, agents, from_csv ? &records : nullptr);
        const uint32_t n_agents = cfg->number_of_agents, capacity = n_agents + extra_capacity;
        if (plan) {
            e->multi = true;
            e->n_regions = plan->n_regions;
            e->migration_enabled = plan->migration_enabled != 0;
            e->commute_enabled = plan->commute_enabled != 0;
            e->start_migration_hour = plan->start_migration_hour;
            e->end_migration_hour = plan->end_migration_hour;
            const size_t R = (size_t)plan->n_regions;
            e->migration_row.assign(R, 0);
            e->commute_row.assign(R, 0);
            if (e->migration_enabled) {
                e->migration_row.assign(plan->migration + (size_t)region * R, plan->migration + (size_t)(region + 1) * R);
                e->migration_mat.assign(plan->migration, plan->migration + R * R);
            }
            if (e->commute_enabled) {
                e->commute_mat.assign(plan->commute, plan->commute + R * R);
                e->commute_row.assign(plan->commute + (size_t)region * R, plan->commute + (size_t)(region + 1) * R);
                apply_commute_plan(agents, n_agents, region, e->commute_row);
            }
            // houses that have residents enter the heap with their resident count; every office enters with its number of
            // workers whose work region is this one (grid.rs:125-155, 262-277).  Stored in tie order.
            e->occ_house0.assign(e->geo.n_houses, 0);
            e->occ_office0.assign(e->geo.n_offices, 0);
            for (uint32_t i = 0; i < n_agents; ++i) {
                e->occ_house0[house_rank_of_index(e->geo, house_index_of(e->geo, agents.home[i]))]++;
                const bool working = ((agents.st[i] >> ST_WS_SHIFT) & 3u) != WS_NA;
                if (working && ((agents.reg[i] >> 8) & 0xFFu) == (uint32_t)region) e->occ_office0[office_rank_of_index(e->geo, office_index_of(e->geo, agents.work[i]))]++;
            }
            for (uint32_t& v : e->occ_house0)
                if (v == 0) v = OCC_ABSENT;
        }
        // empty slots for arrivals
        agents.resize(capacity);
        for (uint32_t i = n_agents; i < capacity; ++i) { agents.st[i] = ST_ABSENT; agents.cell[i] = agents.t0[i] = agents.home[i] = agents.work[i] = agents.wsa[i] = agents.reg[i] = 0; }
        e->free_stack0.clear();
        for (uint32_t sl = capacity; sl-- > n_agents;) e->free_stack0.push_back(sl);  // pop order: n, n+1, ...
        e->P.n = capacity;
        uint32_t bits = 1;
        while ((1ull << bits) < (uint64_t)capacity) ++bits;
        e->P.id_bits = bits;
        e->population = n_agents;
        if (cudaSetDevice(device) != cudaSuccess) { e->err = "cudaSetDevice failed"; return fail(EPI_ERR_CUDA); }
        if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) { e->err = "cudaStreamCreate failed"; return fail(EPI_ERR_CUDA); }
        e->stream = e->own_stream;
        const size_t n = e->P.n, cells = grid_alloc_bytes(e);
        bool ok = true;
        ok &= dev_alloc(e, &e->D.cell, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.st, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.t0, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.home, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.work, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.wsa, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.prop, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.reg, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->i_reg, n) == cudaSuccess;
        ok &= cudaMallocHost((void**)&e->h_small, 64 * sizeof(uint32_t)) == cudaSuccess;
        ok &= dev_alloc(e, &e->grid_alloc, grid_alloc_bytes(e)) == cudaSuccess;
        e->D.grid = e->grid_alloc;
        ok &= dev_alloc(e, &e->D.claim, cells) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.counts, (size_t)RING_ROWS * 8) == cudaSuccess;
        ok &= dev_alloc(e, &e->D.tot, (size_t)TOT_COPIES * 8) == cudaSuccess;
        ok &= dev_alloc(e, &e->d_clock, 1) == cudaSuccess;
        ok &= dev_alloc(e, &e->d_misc, 4) == cudaSuccess;
        ok &= dev_alloc(e, &e->i_cell, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->i_st, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->i_t0, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->i_home, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->i_work, n) == cudaSuccess;
        ok &= dev_alloc(e, &e->i_wsa, n) == cudaSuccess;
        ok &= cudaMallocHost((void**)&e->h_counts, (size_t)RING_ROWS * 8 * sizeof(uint32_t)) == cudaSuccess;
        if (!ok) { e->err = std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError()); return fail(EPI_ERR_CUDA); }
        e->D.hosp_first = e->d_misc;
        e->D.clock = e->d_clock;
        e->D.draws = nullptr;
        e->D.trace = nullptr;
        if (std::getenv("EPI_TRACE")) {  // debugging timeline, read back with epi_debug_trace
            ok = dev_alloc(e, &e->D.trace, (size_t)1 << 16) == cudaSuccess;
            if (ok) cudaMemset(e->D.trace, 0, sizeof(unsigned long long) << 16);
        }
        int rc = EPI_OK;
        if (plan) {
            rc = alloc_travel(e, plan);
            if (rc) return fail(rc);
            rc = reset_travel_state(e);
            if (rc) return fail(rc);
        }
        rc = upload_agents(e, agents);
        if (rc) return fail(rc);
        rc = snapshot_initial(e);
        if (rc) return fail(rc);
        rc = rebuild_grid(e);
        if (rc) return fail(rc);
        rc = recount_travel(e);
        if (rc) return fail(rc);
        initial_counts(e);
        rc = setup_tiles(e, false);
        if (rc) return fail(rc);
        rc = rebuild_tiles(e);
        if (rc) return fail(rc);
    } catch (const std::exception& ex) {
        e->err = ex.what();
        return fail(EPI_ERR_CONFIG);
    }
    *out = e;
    return EPI_OK;
}

void epi_destroy(epi_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->comm.reset();
    drop_graph(e);
    for (auto& pe : e->pending_events) { cudaEventDestroy(pe.second.first); cudaEventDestroy(pe.second.second); }
    void* ptrs[] = {e->D.cell, e->D.st, e->D.t0, e->D.home, e->D.work, e->D.wsa, e->D.prop, e->grid_alloc, e->D.claim, e->D.counts, e->D.tot,
                    e->d_clock, e->d_misc, e->d_draws, e->i_cell, e->i_st, e->i_t0, e->i_home, e->i_work, e->i_wsa, e->D.reg, e->i_reg, e->D.trace};
    for (void* p : ptrs) if (p) cudaFree(p);
    void* tile_ptrs[] = {e->tile_ptrs[0].perm, e->tile_ptrs[0].start, e->tile_ptrs[0].dirty, e->tile_ptrs[1].perm, e->tile_ptrs[1].start, e->tile_ptrs[1].dirty,
                         e->tile_keys_a, e->tile_keys_b, e->tile_ids, e->d_tile_misc, e->tile_temp};
    for (void* p : tile_ptrs) if (p) cudaFree(p);
    for (void* p : e->travel_allocs) cudaFree(p);
    if (e->h_tv) cudaFreeHost(e->h_tv);
    if (e->h_outgoing) cudaFreeHost(e->h_outgoing);
    if (e->h_counts) cudaFreeHost(e->h_counts);
    if (e->h_small) cudaFreeHost(e->h_small);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

uint32_t epi_population(const epi_engine* e) { return e ? e->population : 0; }
uint32_t epi_capacity(const epi_engine* e) { return e ? e->P.n : 0; }

int epi_counts_at_start(const epi_engine* e, epi_counts* out) {
    if (!e || !out) return engine_fail(e, EPI_ERR_ARG, "null argument");
    const epi_config& c = e->cfg;
    const uint32_t total = c.exposed + c.infected_mild_asymptomatic + c.infected_mild_symptomatic + c.infected_severe;
    *out = epi_counts{0, c.number_of_agents - total, c.exposed, total - c.exposed, 0, 0, 0};
    return EPI_OK;
}

int epi_set_stream(epi_engine* e, void* cuda_stream) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    drop_graph(e);
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return EPI_OK;
}

int epi_sync(epi_engine* e) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    return EPI_OK;
}

int epi_reset(epi_engine* e) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    const size_t nb = (size_t)e->P.n * sizeof(uint32_t);
    e->pend_kind.clear();
    e->pend_population.clear();
    CU(cudaMemcpyAsync(e->D.cell, e->i_cell, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.st, e->i_st, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.t0, e->i_t0, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.home, e->i_home, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.work, e->i_work, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.wsa, e->i_wsa, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.reg, e->i_reg, nb, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemsetAsync(e->D.prop, 0, nb, e->stream));
    {
        const int rc = reset_travel_state(e);
        if (rc) return rc;
    }
    if (e->P.hospital_gen != 0) { e->P.hospital_gen = 0; drop_graph(e); }
    initial_counts(e);
    e->interventions = epi::Interventions(e->cfg);
    e->events.clear();
    e->outgoing_travels.clear();
    e->outgoing_staged_hour = 0;
    {
        const int rc = recount_travel(e);
        if (rc) return rc;
    }
    return rebuild_grid(e);
}

int epi_step(epi_engine* e, uint32_t hour, epi_counts* out) {
    if (!e || !out) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    return run_chunk(e, hour, 1, false, out);
}

int epi_enqueue_hour(epi_engine* e, uint32_t hour) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    if (e->pend_kind.size() == 1 && e->pend_kind[0] == 2) { e->pend_kind.clear(); e->pend_population.clear(); }  // the previous exchange hour was settled by epi_finish_hour
    return queue_hours(e, hour, 1, true);
}

int epi_enqueue_hours(epi_engine* e, uint32_t first_hour, uint32_t n_hours) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    if (e->pend_kind.size() == 1 && e->pend_kind[0] == 2) { e->pend_kind.clear(); e->pend_population.clear(); }
    return queue_hours(e, first_hour, n_hours, false);
}

int epi_collect_hours(epi_engine* e, epi_counts* rows_out, uint32_t max_rows, uint32_t* n_rows) {
    if (!e || !n_rows || (!rows_out && max_rows)) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    *n_rows = 0;
    std::vector<epi_counts> rows;
    int rc = collect_hours(e, rows);
    if (rc) return rc;
    if (rows.size() > max_rows) return engine_fail(e, EPI_ERR_ARG, "epi_collect_hours: rows_out too small");
    for (const epi_counts& c : rows) {
        rows_out[(*n_rows)++] = c;
        rc = process_interventions(e, c, false);  // allocation_map.rs:306-337, as in epi_simulate_hours
        if (rc) return rc;
    }
    return EPI_OK;
}

uint32_t epi_next_decision_hour(const epi_engine* e, uint32_t hour) {
    if (!e) return hour;
    // the next hour whose Counts the host must see before the following hour may run: start of day (lockdown.rs:55,
    // hospital.rs:55,70), a configured vaccination hour (vaccination.rs:52), the unlock hour (lockdown.rs:69-73)
    const epi::Interventions& iv = e->interventions;
    uint32_t d = (hour + 23u) / 24u * 24u;
    d = std::min(d, iv.vaccinate.next_hour(hour));
    if (iv.lockdown.is_locked_down() && iv.lockdown.unlock_hour() >= hour) d = std::min(d, iv.lockdown.unlock_hour());
    return d;
}

int epi_step_with_draws(epi_engine* e, uint32_t hour, const uint64_t* draws, epi_counts* out) {
    if (!e || !out || !draws) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    const size_t count = (size_t)e->P.n * EPI_DRAWS_PER_AGENT;
    if (e->draws_capacity < count) {
        if (e->d_draws) cudaFree(e->d_draws);
        e->d_draws = nullptr;
        CU(dev_alloc(e, &e->d_draws, count));
        e->draws_capacity = count;
    }
    CU(cudaMemcpyAsync(e->d_draws, draws, count * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream));
    e->D.draws = e->d_draws;
    const int rc = run_chunk(e, hour, 1, true, out);
    e->D.draws = nullptr;
    return rc;
}

int epi_run_hours(epi_engine* e, uint32_t first_hour, uint32_t n_hours, epi_counts* rows_out) {
    if (!e || (!rows_out && n_hours)) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    uint32_t done = 0;
    while (done < n_hours) {
        const uint32_t n = std::min(n_hours - done, RING_ROWS);
        const int rc = run_chunk(e, first_hour + done, n, false, rows_out + done);
        if (rc) return rc;
        done += n;
    }
    return EPI_OK;
}

int epi_lock_city(epi_engine* e) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    { Timed t(e, KK_SWEEP); launch_lock(e->P, e->D, e->stream); }
    CU(cudaGetLastError());
    return EPI_OK;
}
int epi_unlock_city(epi_engine* e) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    { Timed t(e, KK_SWEEP); launch_unlock(e->P, e->D, e->stream); }
    CU(cudaGetLastError());
    return EPI_OK;
}
int epi_vaccinate(epi_engine* e, double p, uint32_t hour) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    if (!(p >= 0.0 && p <= 1.0)) return engine_fail(e, EPI_ERR_ARG, "vaccination percentage out of [0,1]");  // gen_bool panics
    CU(cudaSetDevice(e->device));
    { Timed t(e, KK_SWEEP); launch_vaccinate(e->P, e->D, bernoulli_threshold(p), hour, e->stream); }
    CU(cudaGetLastError());
    return EPI_OK;
}
int epi_expand_hospital(epi_engine* e) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    if (e->P.hospital_gen != 1) { e->P.hospital_gen = 1; drop_graph(e); }
    return EPI_OK;
}

int epi_get_state(epi_engine* e, int32_t* cx, int32_t* cy, uint32_t* st, uint32_t* t0, uint32_t* home, uint32_t* work, uint32_t* wsa) {
    if (!e || !cx || !cy || !st || !t0 || !home || !work || !wsa) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    const size_t n = e->P.n, nb = n * sizeof(uint32_t);
    std::vector<uint32_t> cell(n);
    CU(cudaMemcpyAsync(cell.data(), e->D.cell, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(st, e->D.st, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(t0, e->D.t0, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(home, e->D.home, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(work, e->D.work, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(wsa, e->D.wsa, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < n; ++i) {
        cx[i] = (int32_t)(cell[i] & CELL_XMASK);
        cy[i] = (int32_t)(cell[i] >> CELL_BITS);
        // canonical form: fields that the reference's enum variant does not carry read as 0
        const uint32_t state = st[i] & ST_STATE_MASK, sev = (st[i] >> ST_SEV_SHIFT) & 3u, ws = (st[i] >> ST_WS_SHIFT) & 3u;
        if (state == ST_ABSENT) { cx[i] = cy[i] = 0; st[i] = ST_ABSENT; t0[i] = home[i] = work[i] = wsa[i] = 0; continue; }
        if (!(state == ST_E || (state == ST_I && sev == SEV_PRE))) t0[i] = 0;
        home[i] = house_index_of(e->geo, home[i]);
        work[i] = ws == WS_NA ? 0u : office_index_of(e->geo, work[i]);
        if (ws != WS_STAFF) wsa[i] = 0;
    }
    return EPI_OK;
}

int epi_citizen_states(epi_engine* e, char* state_out, int32_t* x_out, int32_t* y_out, uint32_t* slot_out, uint32_t capacity, uint32_t* n_out) {
    if (!e || !n_out || (capacity && (!state_out || !x_out || !y_out))) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    const size_t n = e->P.n, nb = n * sizeof(uint32_t);
    std::vector<uint32_t> cell(n), st(n);
    CU(cudaMemcpyAsync(cell.data(), e->D.cell, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(st.data(), e->D.st, nb, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    static const char letters[5] = {'s', 'e', 'i', 'r', 'd'};  // CitizenState::state_str (citizen_state.rs:34-42)
    uint32_t live = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t state = st[i] & ST_STATE_MASK;
        if (state > ST_D) continue;  // empty slot
        if (live < capacity) {
            state_out[live] = letters[state];
            x_out[live] = (int32_t)(cell[i] & CELL_XMASK);
            y_out[live] = (int32_t)(cell[i] >> CELL_BITS);
            if (slot_out) slot_out[live] = (uint32_t)i;
        }
        ++live;
    }
    *n_out = live;
    return EPI_OK;
}

int epi_build_population(const epi_config* cfg, uint64_t seed, int32_t* cx, int32_t* cy, uint32_t* st, uint32_t* t0, uint32_t* home, uint32_t* work,
                         uint32_t* wsa) {
    if (!cfg || !cx || !cy || !st || !t0 || !home || !work || !wsa) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    try {
        PopulationRecords records;
        const epi_config c = resolve_population(*cfg, records);
        const std::string bad = validate_config(c);
        if (!bad.empty()) return engine_fail(nullptr, EPI_ERR_CONFIG, bad);
        const Geometry geo = make_geometry(c.grid_size, c.number_of_agents, c.hospital_beds_percentage);
        HostAgents a;
        build_population(c, geo, seed, 0, a, c.population_csv_file[0] ? &records : nullptr);
        for (size_t i = 0; i < a.size(); ++i) {
            const uint32_t ws = (a.st[i] >> ST_WS_SHIFT) & 3u;
            cx[i] = (int32_t)(a.cell[i] & CELL_XMASK);
            cy[i] = (int32_t)(a.cell[i] >> CELL_BITS);
            st[i] = a.st[i];
            t0[i] = 0;
            home[i] = house_index_of(geo, a.home[i]);
            work[i] = ws == WS_NA ? 0u : office_index_of(geo, a.work[i]);
            wsa[i] = ws == WS_STAFF ? a.wsa[i] : 0u;
        }
    } catch (const std::exception& ex) {
        return engine_fail(nullptr, EPI_ERR_CONFIG, ex.what());
    }
    return EPI_OK;
}

int epi_population_size(const epi_config* cfg, uint32_t* n) {
    if (!cfg || !n) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    try {
        PopulationRecords records;
        *n = resolve_population(*cfg, records).number_of_agents;
    } catch (const std::exception& ex) {
        return engine_fail(nullptr, EPI_ERR_IO, ex.what());
    }
    return EPI_OK;
}

int epi_set_state(epi_engine* e, uint32_t n, const int32_t* cx, const int32_t* cy, const uint32_t* st, const uint32_t* t0, const uint32_t* home,
                  const uint32_t* work, const uint32_t* wsa) {
    if (!e || !cx || !cy || !st || !t0 || !home || !work || !wsa) return engine_fail(e, EPI_ERR_ARG, "null argument");
    if (n != e->P.n) return engine_fail(e, EPI_ERR_ARG, "epi_set_state: n must equal epi_capacity()");
    if (e->multi) return engine_fail(e, EPI_ERR_STATE, "epi_set_state is not available on a multi-region engine");
    CU(cudaSetDevice(e->device));
    // The caller's arrays go to the device as they are (no host-side pass over the population); a kernel packs the cells and
    // turns the house / office indices into origins.  cell_x / cell_y are staged in the claim array, which is zeroed again
    // before its next use (claim_dirty).
    const size_t nb = (size_t)n * sizeof(uint32_t);
    if (grid_alloc_bytes(e) * sizeof(uint32_t) < 2 * nb) return engine_fail(e, EPI_ERR_STATE, "epi_set_state: grid too small to stage the cell arrays");
    int32_t* d_cx = (int32_t*)e->D.claim;
    int32_t* d_cy = d_cx + n;
    e->claim_dirty = true;
    CU(cudaMemcpyAsync(d_cx, cx, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_cy, cy, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.st, st, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.t0, t0, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.home, home, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.work, work, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->D.wsa, wsa, nb, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(e->d_misc + 2, 0, sizeof(uint32_t), e->stream));
    {
        Timed t(e, KK_MISC);
        launch_import_state(e->P, e->D, d_cx, d_cy, e->geo.n_houses, e->geo.n_offices, e->d_misc + 2, e->stream);
    }
    uint32_t bad = 0;
    CU(cudaMemcpyAsync(&bad, e->d_misc + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (bad) return engine_fail(e, EPI_ERR_ARG, "epi_set_state: " + std::to_string(bad) + " agents with a cell outside the grid or a house / office index out of range");
    const int rc = rebuild_grid(e);
    if (rc) return rc;
    return rebuild_tiles(e);  // home / work / work status may have changed
}

int epi_get_regions(epi_engine* e, uint32_t* reg) {
    if (!e || !reg) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    std::vector<uint32_t> st(e->P.n);
    CU(cudaMemcpyAsync(reg, e->D.reg, (size_t)e->P.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaMemcpyAsync(st.data(), e->D.st, (size_t)e->P.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (uint32_t i = 0; i < e->P.n; ++i)
        if ((st[i] & ST_STATE_MASK) == ST_ABSENT) reg[i] = 0;
    return EPI_OK;
}

int epi_geometry(const epi_engine* e, int32_t* out) {
    if (!e || !out) return engine_fail(e, EPI_ERR_ARG, "null argument");
    const Rect rs[4] = {e->geo.housing, e->geo.transport, e->geo.work, e->P.hospital()};
    for (int i = 0; i < 4; ++i) { out[4 * i] = rs[i].sx; out[4 * i + 1] = rs[i].sy; out[4 * i + 2] = rs[i].ex; out[4 * i + 3] = rs[i].ey; }
    out[16] = (int32_t)e->geo.n_houses; out[17] = (int32_t)e->geo.n_offices; out[18] = e->geo.grid_size;
    return EPI_OK;
}

int epi_get_grid(epi_engine* e, uint8_t* out, uint64_t capacity, uint32_t* pitch, uint32_t* rows) {
    if (!e || !pitch || !rows) return engine_fail(e, EPI_ERR_ARG, "null argument");
    *pitch = e->geo.pitch; *rows = e->geo.rows;
    if (!out) return EPI_OK;
    if (capacity < n_cells(e)) return engine_fail(e, EPI_ERR_ARG, "epi_get_grid: buffer too small");
    CU(cudaSetDevice(e->device));
    std::vector<uint8_t> raw(grid_alloc_bytes(e));
    CU(cudaMemcpyAsync(raw.data(), e->D.grid, raw.size(), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    for (uint32_t y = 0; y < e->geo.rows; ++y)  // device layout -> row-major [rows][pitch]
        for (uint32_t x = 0; x < e->geo.pitch; ++x) out[(size_t)y * e->geo.pitch + x] = raw[e->P.cell_offset((int)x, (int)y)];
    return EPI_OK;
}

int epi_set_kernel_timing(epi_engine* e, int on) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    drain_events(e);
    e->timing = on != 0;
    for (int k = 0; k < EPI_N_KERNEL_KINDS; ++k) { e->kernel_ms[k] = 0; e->kernel_launches[k] = 0; }
    for (int k = 0; k < 48; ++k) { e->hour_ms[k] = 0; e->hour_launches[k] = 0; }
    return EPI_OK;
}
int epi_get_kernel_times(epi_engine* e, double* ms_total, uint64_t* launches) {
    if (!e || !ms_total || !launches) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    drain_events(e);
    for (int k = 0; k < EPI_N_KERNEL_KINDS; ++k) { ms_total[k] = e->kernel_ms[k]; launches[k] = e->kernel_launches[k]; }
    return EPI_OK;
}
int epi_get_hour_times(epi_engine* e, double* ms_total, uint64_t* launches) {
    if (!e || !ms_total || !launches) return engine_fail(e, EPI_ERR_ARG, "null argument");
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    drain_events(e);
    for (int k = 0; k < 48; ++k) { ms_total[k] = e->hour_ms[k]; launches[k] = e->hour_launches[k]; }
    return EPI_OK;
}
uint64_t epi_launch_count(const epi_engine* e, int reset) {
    if (!e) return 0;
    const uint64_t v = e->launches;
    if (reset) const_cast<epi_engine*>(e)->launches = 0;
    return v;
}
uint64_t epi_device_bytes(const epi_engine* e) { return e ? e->device_bytes : 0; }
uint64_t epi_epoch_resets(const epi_engine* e) { return e ? e->epoch_resets : 0; }

uint64_t epi_tile_hours(const epi_engine* e) { return e ? e->tile_hours : 0; }

int epi_debug_trace(epi_engine* e, uint64_t* out, uint64_t max_words, uint64_t* n_words) {
    if (!e || !out || !n_words) return engine_fail(e, EPI_ERR_ARG, "null argument");
    *n_words = 0;
    if (!e->D.trace) return EPI_OK;
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    unsigned long long used = 0;
    CU(cudaMemcpy(&used, e->D.trace, sizeof(used), cudaMemcpyDeviceToHost));
    used = std::min<unsigned long long>(used, (1ull << 16) - 2ull);
    used = std::min<unsigned long long>(used, max_words);
    CU(cudaMemcpy(out, e->D.trace + 1, used * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU(cudaMemset(e->D.trace, 0, sizeof(unsigned long long)));
    *n_words = used;
    return EPI_OK;
}

int epi_set_tiles(epi_engine* e, int on) {
    if (!e) return engine_fail(e, EPI_ERR_ARG, "null engine");
    CU(cudaSetDevice(e->device));
    if (!on) {
        if (e->tiles_ready) drop_graph(e);
        e->tiles_ready = false;
        e->tiles_enabled = false;
        return EPI_OK;
    }
    if (e->multi) return engine_fail(e, EPI_ERR_STATE, "the tile kernels serve standalone engines");
    if (e->tiles_ready) return EPI_OK;
    {
        const int rc = setup_tiles(e, true);
        if (rc) return rc;
    }
    if (!e->tiles_enabled) return engine_fail(e, EPI_ERR_STATE, "the tile kernels cannot serve this geometry");
    CU(cudaMemsetAsync(e->D.prop, 0, (size_t)e->P.n * sizeof(uint32_t), e->stream));  // the tile kernels start from all-zero proposals
    return rebuild_tiles(e);
}

}  // extern "C"
