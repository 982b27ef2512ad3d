// The multi-region `Configuration` (common/src/config/configuration.rs:28-117, travel_plan_config.rs:22-71) and one rank of
// `engine-app -m mpi` (engine-app/src/main.rs:131-166 + Epidemiology::run_multi_engine + the listeners' output files).
// C ABI: epi_configuration_* and epi_run_region (include/epi.h).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>

#include "engine.h"
#include "json.h"
#include "simulation.h"

using namespace epi;

struct epi_configuration {
    std::vector<std::string> engine_ids;  // engine_configs order
    std::vector<epi_config> configs;
    std::vector<std::string> regions;     // travel_plan.regions: index == region == rank (mpi_transport.rs:44-52)
    bool migration_enabled = false, commute_enabled = false;
    std::vector<uint32_t> migration, commute;  // [from][to] over `regions`
    uint32_t start_migration_hour = 0, end_migration_hour = 0;
    std::vector<int> engine_of_region;  // region index -> index in engine_configs
};

namespace {

constexpr double TRANSPORT_AREA_RELATIVE_SIZE = 0.2;  // engine/src/models/constants.rs:23

std::vector<uint32_t> read_matrix(const JsonValue& block, const char* name, size_t R) {
    const JsonValue* m = block.find("matrix");
    if (!m || m->kind != JsonValue::Array) throw std::runtime_error(std::string("travel_plan.") + name + ".matrix is required when enabled");
    if (m->arr.size() != R) throw std::runtime_error(std::string("travel_plan.") + name + ".matrix must have one row per region");
    std::vector<uint32_t> out(R * R);
    for (size_t i = 0; i < R; ++i) {
        const JsonValue& row = m->arr[i];
        if (row.kind != JsonValue::Array || row.arr.size() != R) throw std::runtime_error(std::string("travel_plan.") + name + ".matrix must be square");
        for (size_t j = 0; j < R; ++j) out[i * R + j] = row.arr[j].as_u32("matrix entry");
    }
    return out;
}

// Configuration::read (configuration.rs:49-57) + TravelPlanConfig::validate_regions (travel_plan_config.rs:60-62)
void read_configuration(const std::string& path, epi_configuration& c) {
    const JsonValue doc = json_parse(json_read_file(path));
    const JsonValue& engines = doc.at("engine_configs");
    const JsonValue& tp = doc.at("travel_plan");
    if (engines.kind != JsonValue::Array) throw std::runtime_error("`engine_configs` must be an array");
    const JsonValue& regions = tp.at("regions");
    if (regions.kind != JsonValue::Array) throw std::runtime_error("`travel_plan.regions` must be an array");
    for (const JsonValue& r : regions.arr) c.regions.push_back(r.as_string("regions entry"));
    for (const JsonValue& ec : engines.arr) {
        c.engine_ids.push_back(ec.at("engine_id").as_string("engine_id"));
        epi_config cfg;
        config_from_value(ec.at("config"), cfg);
        c.configs.push_back(cfg);
    }
    const size_t R = c.regions.size();
    bool match = c.engine_ids.size() == R;
    c.engine_of_region.assign(R, -1);
    for (size_t i = 0; i < c.engine_ids.size() && match; ++i) {
        const auto it = std::find(c.regions.begin(), c.regions.end(), c.engine_ids[i]);
        if (it == c.regions.end()) match = false;
        else c.engine_of_region[(size_t)(it - c.regions.begin())] = (int)i;
    }
    for (int v : c.engine_of_region) match = match && v >= 0;
    if (!match) throw std::runtime_error("Engine names should match regions in travel plan");
    const JsonValue& mig = tp.at("migration");
    const JsonValue& com = tp.at("commute");
    c.migration_enabled = mig.at("enabled").as_bool("migration.enabled");
    c.commute_enabled = com.at("enabled").as_bool("commute.enabled");
    if (c.migration_enabled) c.migration = read_matrix(mig, "migration", R);
    if (c.commute_enabled) c.commute = read_matrix(com, "commute", R);
    if (const JsonValue* v = mig.find("start_migration_hour")) c.start_migration_hour = v->as_u32("start_migration_hour");
    if (const JsonValue* v = mig.find("end_migration_hour")) c.end_migration_hour = v->as_u32("end_migration_hour");
}

uint64_t row_sum(const std::vector<uint32_t>& m, size_t R, size_t r) {
    uint64_t s = 0;
    for (size_t k = 0; k < R; ++k) s += m[r * R + k];
    return s;
}
uint64_t col_sum(const std::vector<uint32_t>& m, size_t R, size_t r) {
    uint64_t s = 0;
    for (size_t k = 0; k < R; ++k) s += m[k * R + r];
    return s;
}

// Configuration::validate (configuration.rs:59-117): transport capacity and the grid / population ratio of every engine
void validate_configuration(const epi_configuration& c) {
    const size_t R = c.regions.size();
    for (size_t r = 0; r < R; ++r) {
        const size_t i = (size_t)c.engine_of_region[r];
        const epi_config& cfg = c.configs[i];
        const bool is_auto = cfg.population_csv_file[0] == 0;
        const int64_t grid_size = cfg.grid_size;
        const int64_t n_agents = is_auto ? cfg.number_of_agents : 0;
        const double pt = is_auto ? cfg.public_transport_percentage : 0.0;
        int64_t total_population = n_agents;
        const int64_t transport_cells = ((int64_t)std::ceil((double)grid_size * TRANSPORT_AREA_RELATIVE_SIZE) - 1) * grid_size;
        if (c.commute_enabled) {
            const int64_t incoming = (int64_t)col_sum(c.commute, R, r), outgoing = (int64_t)row_sum(c.commute, R, r);
            if ((int64_t)std::ceil((double)n_agents * pt) - outgoing + incoming > transport_cells)
                throw std::runtime_error("For engine id - " + c.engine_ids[i] + ", Incoming commuters are more than engine transport capacity");
            total_population += incoming - outgoing;
        }
        if (c.migration_enabled) total_population += (int64_t)col_sum(c.migration, R, r) - (int64_t)row_sum(c.migration, R, r);
        if (total_population <= 0 || (grid_size * grid_size) / total_population < 3)
            throw std::runtime_error(c.engine_ids[i] + ": Not enough space to accumulate the migrators/commuters");
    }
}

// agent slots to reserve for the arrivals of region r: every commuter of a day plus the migrators of the whole window
uint32_t arrival_capacity(const epi_configuration& c, size_t r, uint32_t hours) {
    const size_t R = c.regions.size();
    uint64_t extra = 0;
    if (c.commute_enabled) extra += col_sum(c.commute, R, r);
    if (c.migration_enabled) {
        const uint32_t last = std::min(c.end_migration_hour, hours);
        const uint64_t days = last >= c.start_migration_hour ? (last - c.start_migration_hour) / 24u + 1u : 0u;
        extra += col_sum(c.migration, R, r) * days;
    }
    return (uint32_t)std::min<uint64_t>(extra + extra / 8 + 1024, 1u << 26);
}

// CsvListener, InterventionReporter and TravelCounter at simulation_ended (listeners/csv_service.rs:44-71,
// intervention_reporter.rs:28-63, travel_counter.rs:69-79)
void write_outputs(const std::string& base, const epi_counts* rows, uint32_t n_rows, const epi_intervention_event* ev, uint32_t n_events,
                   const epi_outgoing_travel* tr, uint32_t n_travels, const std::vector<std::string>* region_names) {
    Listeners l;
    l.counts.assign(rows, rows + n_rows);
    static const char* names[3] = {"lockdown", "vaccination", "build_new_hospital"};
    for (uint32_t i = 0; i < n_events; ++i) {
        const epi_intervention_event& x = ev[i];
        if (x.kind < 0 || x.kind > 2) throw std::runtime_error("intervention event of unknown kind");
        l.interventions.push_back({x.hour, names[x.kind], x.kind == 0 ? (x.status ? "{\"status\":\"locked_down\"}" : "{\"status\":\"lockdown_revoked\"}") : "{}"});
    }
    l.simulation_ended(base);
    if (!region_names) return;
    std::ofstream f(base + "_outgoing_travels.csv");
    if (!f) throw std::runtime_error("Failed to write to file " + base + "_outgoing_travels.csv");
    if (n_travels) f << "hr,destination,susceptible,exposed,infected,recovered\n";  // csv::Writer::serialize writes the header with the first record
    for (uint32_t i = 0; i < n_travels; ++i) {
        const epi_outgoing_travel& t = tr[i];
        if (t.destination >= region_names->size()) throw std::runtime_error("outgoing travel to an unknown region");
        f << t.hr << ',' << (*region_names)[t.destination] << ',' << t.susceptible << ',' << t.exposed << ',' << t.infected << ',' << t.recovered << '\n';
    }
}

void write_region_outputs(const std::string& base, const std::vector<epi_counts>& rows, epi_engine* e, const epi_configuration& c) {
    uint32_t n_events = 0, n_travels = 0;
    epi_intervention_events(e, nullptr, 0, &n_events);
    std::vector<epi_intervention_event> ev(n_events);
    if (n_events) epi_intervention_events(e, ev.data(), n_events, &n_events);
    epi_outgoing_travels(e, nullptr, 0, &n_travels);
    std::vector<epi_outgoing_travel> tr(n_travels);
    if (n_travels) epi_outgoing_travels(e, tr.data(), n_travels, &n_travels);
    write_outputs(base, rows.data(), (uint32_t)rows.size(), ev.data(), n_events, tr.data(), n_travels, &c.regions);
}

}  // namespace

extern "C" {

int epi_configuration_read(const char* json_path, epi_configuration** out) {
    if (!json_path || !out) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    *out = nullptr;
    auto c = std::make_unique<epi_configuration>();
    try {
        read_configuration(json_path, *c);
    } catch (const std::exception& ex) {
        const std::string msg = ex.what();
        return engine_fail(nullptr, msg.rfind("cannot open", 0) == 0 ? EPI_ERR_IO : EPI_ERR_CONFIG, msg);
    }
    try {
        validate_configuration(*c);
    } catch (const std::exception& ex) {
        return engine_fail(nullptr, EPI_ERR_CONFIG, ex.what());
    }
    *out = c.release();
    return EPI_OK;
}

void epi_configuration_free(epi_configuration* c) { delete c; }

int epi_configuration_regions(const epi_configuration* c) { return c ? (int)c->regions.size() : 0; }

const char* epi_configuration_region_name(const epi_configuration* c, int region) {
    return (c && region >= 0 && (size_t)region < c->regions.size()) ? c->regions[(size_t)region].c_str() : nullptr;
}

int epi_configuration_engine_config(const epi_configuration* c, int region, epi_config* out) {
    if (!c || !out || region < 0 || (size_t)region >= c->regions.size()) return engine_fail(nullptr, EPI_ERR_ARG, "epi_configuration_engine_config: bad argument");
    *out = c->configs[(size_t)c->engine_of_region[(size_t)region]];
    return EPI_OK;
}

int epi_configuration_travel_plan(const epi_configuration* c, int n_regions, epi_travel_plan* out, uint32_t* migration_out, uint32_t* commute_out) {
    if (!c || !out || !migration_out || !commute_out) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    const size_t R = c->regions.size();
    if (n_regions < 1 || (size_t)n_regions > R) return engine_fail(nullptr, EPI_ERR_ARG, "epi_configuration_travel_plan: n_regions out of range");
    const size_t n = (size_t)n_regions;
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < n; ++j) {
            migration_out[i * n + j] = c->migration_enabled ? c->migration[i * R + j] : 0u;
            commute_out[i * n + j] = c->commute_enabled ? c->commute[i * R + j] : 0u;
        }
    out->n_regions = n_regions;
    out->migration_enabled = c->migration_enabled;
    out->commute_enabled = c->commute_enabled;
    out->migration = migration_out;
    out->commute = commute_out;
    out->start_migration_hour = c->start_migration_hour;
    out->end_migration_hour = c->end_migration_hour;
    return EPI_OK;
}

int epi_write_outputs(const char* output_dir, const char* engine_id, const epi_counts* rows, uint32_t n_rows, const epi_intervention_event* events,
                      uint32_t n_events, const epi_outgoing_travel* travels, uint32_t n_travels, const char* const* region_names, char* base_out,
                      uint64_t base_bytes) {
    if (!output_dir || !engine_id || (!rows && n_rows) || (!events && n_events) || (travels && !region_names))
        return engine_fail(nullptr, EPI_ERR_ARG, "epi_write_outputs: null argument");
    try {
        const std::string base = output_file_format(output_dir, engine_id);
        std::vector<std::string> names;
        if (travels) {
            uint32_t most = 0;
            for (uint32_t i = 0; i < n_travels; ++i) most = std::max(most, travels[i].destination + 1u);
            for (uint32_t r = 0; r < most; ++r) names.push_back(region_names[r]);
        }
        write_outputs(base, rows, n_rows, events, n_events, travels, n_travels, travels ? &names : nullptr);
        if (base_out) {
            if (base.size() + 1 > base_bytes) return engine_fail(nullptr, EPI_ERR_ARG, "epi_write_outputs: base_out too small");
            std::memcpy(base_out, base.c_str(), base.size() + 1);
        }
    } catch (const std::exception& ex) {
        return engine_fail(nullptr, EPI_ERR_IO, ex.what());
    }
    return EPI_OK;
}

uint32_t epi_configuration_arrival_capacity(const epi_configuration* c, int region) {
    if (!c || region < 0 || (size_t)region >= c->regions.size()) return 0;
    return arrival_capacity(*c, (size_t)region, c->configs[(size_t)c->engine_of_region[(size_t)region]].hours);
}

int epi_run_region(const epi_configuration* c, int region, int n_ranks, const void* unique_id, uint64_t seed, int device, const char* output_dir,
                   int terminate_when_clear, epi_counts* rows_out, uint32_t max_rows, uint32_t* n_rows, double* loop_seconds) {
    if (!c || !unique_id) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    const size_t R_all = c->regions.size();
    if (n_ranks < 1 || (size_t)n_ranks > R_all || region < 0 || region >= n_ranks) return engine_fail(nullptr, EPI_ERR_ARG, "epi_run_region: region / n_ranks out of range");
    // fewer ranks than regions: the first n_ranks regions of the travel plan take part
    const size_t R = (size_t)n_ranks;
    std::vector<uint32_t> mig(R * R), com(R * R);
    epi_travel_plan plan;
    int rc = epi_configuration_travel_plan(c, n_ranks, &plan, mig.data(), com.data());
    if (rc) return rc;
    const epi_config& cfg = c->configs[(size_t)c->engine_of_region[(size_t)region]];
    const std::string engine_id = c->regions[(size_t)region];
    std::printf("in multi-engine mode\n");
    std::fflush(stdout);
    epi_engine* e = nullptr;
    // the engine's Philox key: seed + region (callers pass one seed for the run)
    rc = epi_create_multi(&cfg, seed + (uint64_t)region, device, region, &plan, arrival_capacity(*c, (size_t)region, cfg.hours), &e);
    if (rc) return rc;
    auto fail = [&](int code) {
        set_global_error("[" + engine_id + "] " + e->err);
        epi_destroy(e);
        return code;
    };
    if ((rc = epi_comm_init(e, n_ranks, region, unique_id))) return fail(rc);
    if ((rc = epi_count_outgoing(e, output_dir != nullptr))) return fail(rc);
    const uint32_t hours = cfg.hours > 0 ? cfg.hours - 1u : 0u;  // for simulation_hour in 1..config.get_hours()
    std::vector<epi_counts> rows(hours);
    const auto start = std::chrono::steady_clock::now();
    uint32_t done = 0;
    while (done < hours) {
        const uint32_t n = std::min(hours - done, 240u);
        uint32_t got = 0;
        epi_engine* one[1] = {e};
        if ((rc = epi_run_multi_hours(one, 1, 1u + done, n, terminate_when_clear, rows.data() + done, &got))) return fail(rc);
        done += got;
        if (got) {
            const epi_counts& r = rows[done - 1];
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
            std::printf("INFO - [%s] hour %u: S: %u, E:%u, I: %u, H: %u, R: %u, D: %u; Throughput: %.2f iterations/sec\n", engine_id.c_str(), r.hour, r.susceptible,
                        r.exposed, r.infected, r.hospitalized, r.recovered, r.deceased, (double)done / el);
            std::fflush(stdout);
        }
        if (got < n) break;
    }
    if ((rc = epi_sync(e))) return fail(rc);
    rows.resize(done);
    const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    std::printf("INFO - [%s] Number of iterations: %u, Total Time taken %.3f seconds; Iterations/sec: %.2f\n", engine_id.c_str(), done, elapsed,
                elapsed > 0 ? (double)done / elapsed : 0.0);
    std::fflush(stdout);
    if (output_dir) {
        try {
            write_region_outputs(output_file_format(output_dir, engine_id), rows, e, *c);
        } catch (const std::exception& ex) {
            e->err = ex.what();
            return fail(EPI_ERR_IO);
        }
    }
    if (n_rows) *n_rows = done;
    if (loop_seconds) *loop_seconds = elapsed;
    if (rows_out)
        for (uint32_t i = 0; i < std::min(max_rows, done); ++i) rows_out[i] = rows[i];
    epi_comm_destroy(e);
    epi_destroy(e);
    return EPI_OK;
}

}  // extern "C"
