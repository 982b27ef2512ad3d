#include "simulation.h"

#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>

#include "engine.h"
#include "json.h"

namespace epi {

void Listeners::simulation_ended(const std::string& base) const {
    {
        // csv crate `Writer::serialize(Counts)`: header row from the field names, '\n' terminated (counts.rs:25-34)
        std::ofstream f(base + ".csv");
        if (!f) throw std::runtime_error("Failed to write to file " + base + ".csv");
        f << "hour,susceptible,exposed,infected,hospitalized,recovered,deceased\n";
        for (const epi_counts& c : counts)
            f << c.hour << ',' << c.susceptible << ',' << c.exposed << ',' << c.infected << ',' << c.hospitalized << ',' << c.recovered << ',' << c.deceased
              << '\n';
    }
    {
        // serde_json::to_writer(Vec<InterventionReport>): compact, field order hour, intervention, data
        std::ofstream f(base + "_interventions.json");
        if (!f) throw std::runtime_error("Failed to create intervention report file");
        f << '[';
        for (size_t i = 0; i < interventions.size(); ++i) {
            if (i) f << ',';
            f << "{\"hour\":" << interventions[i].hour << ",\"intervention\":\"" << interventions[i].intervention << "\",\"data\":" << interventions[i].data << '}';
        }
        f << ']';
    }
}

std::string output_file_format(const std::string& output_dir, const std::string& engine_id) {
    const std::string dir = output_dir + "/output";
    mkdir(output_dir.c_str(), 0755);
    mkdir(dir.c_str(), 0755);
    char stamp[32];
    const std::time_t now = std::time(nullptr);
    std::tm tm_utc;
    gmtime_r(&now, &tm_utc);
    std::strftime(stamp, sizeof(stamp), "%Y-%m-%dT%H:%M:%S", &tm_utc);
    return dir + "/simulation_" + engine_id + "_" + stamp;
}

static void log_counts(const epi_counts& c) {
    std::printf("INFO - S: %u, E:%u, I: %u, H: %u, R: %u, D: %u\n", c.susceptible, c.exposed, c.infected, c.hospitalized, c.recovered, c.deceased);
}

// CitizenLocationMap::process_interventions (allocation_map.rs:306-337) on the Counts row of one hour; in a multi-region
// engine also stop_simulation's MultiEngine arm (epidemiology_simulation.rs:564-571)
int process_interventions(epi_engine* e, const epi_counts& c, bool log) {
    Interventions& iv = e->interventions;
    int rc = EPI_OK;
    if (const double* pct = iv.vaccinate.get_vaccination_percentage(c)) {
        if (log) std::printf("INFO - Vaccination\n");
        rc = epi_vaccinate(e, *pct, c.hour);
        if (rc) return rc;
        e->events.push_back({c.hour, 1, 0});
    }
    if (iv.lockdown.should_apply(c)) {
        iv.lockdown.apply();
        if (log) std::printf("INFO - Locking the city. Hour: %u\n", c.hour);
        rc = epi_lock_city(e);
        if (rc) return rc;
        e->events.push_back({c.hour, 0, 1});
    }
    if (iv.lockdown.should_unlock(c)) {
        if (log) std::printf("INFO - Unlocking city. Hour: %u\n", c.hour);
        rc = epi_unlock_city(e);
        if (rc) return rc;
        iv.lockdown.unapply();
        e->events.push_back({c.hour, 0, 0});
    }
    iv.build_new_hospital.counts_updated(c);
    if (iv.build_new_hospital.should_apply(c)) {
        if (log) std::printf("INFO - Increasing the hospital size\n");
        rc = epi_expand_hospital(e);
        if (rc) return rc;
        iv.build_new_hospital.apply();
        e->events.push_back({c.hour, 2, 0});
    }
    if (e->multi && iv.lockdown.is_locked_down() && c.exposed == 0 && c.infected == 0 && c.hospitalized == 0) iv.lockdown.set_zero_infection_hour(c.hour);
    return EPI_OK;
}

// hours [first_hour, first_hour + n_hours): simulate + process_interventions (+ stop rule)
static int simulate_hours(epi_engine* e, uint32_t first_hour, uint32_t n_hours, bool stop_rule, epi_counts* rows_out, uint32_t* n_rows, int* stopped,
                          bool log, const std::chrono::steady_clock::time_point* start_time) {
    Interventions& iv = e->interventions;
    const uint32_t end_hour = first_hour + n_hours;  // exclusive
    std::vector<epi_counts> seg;
    bool stop = false;
    uint32_t simulation_hour = first_hour, written = 0;
    while (simulation_hour < end_hour && !stop) {
        // Run up to (and including) the next hour at which a host decision can change device state:
        // start of day (lockdown.rs:55, hospital.rs:55,70), a configured vaccination hour (vaccination.rs:52),
        // the unlock hour (lockdown.rs:69-73).  The per-hour stop rule only needs the Counts rows.
        uint32_t seg_end = (simulation_hour + 23u) / 24u * 24u;
        seg_end = std::min(seg_end, iv.vaccinate.next_hour(simulation_hour));
        if (iv.lockdown.is_locked_down() && iv.lockdown.unlock_hour() >= simulation_hour) seg_end = std::min(seg_end, iv.lockdown.unlock_hour());
        seg_end = std::min(seg_end, end_hour - 1u);
        const uint32_t n = seg_end - simulation_hour + 1u;
        seg.resize(n);
        int rc = epi_run_hours(e, simulation_hour, n, seg.data());
        if (rc) return rc;
        for (uint32_t k = 0; k < n && !stop; ++k) {
            const epi_counts& c = seg[k];  // counts_at_hr.increment_hour() + simulate()
            rows_out[written++] = c;       // listeners.counts_updated
            rc = process_interventions(e, c, log);
            if (rc) return rc;
            // Epidemiology::stop_simulation, Standalone arm (epidemiology_simulation.rs:564-575)
            if (stop_rule && !e->multi && c.exposed == 0 && c.infected == 0 && c.hospitalized == 0) stop = true;
            if (!stop && c.hour % 100u == 0 && log && start_time) {
                const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - *start_time).count();
                std::printf("INFO - Throughput: %f iterations/sec; simulation hour %u\n", (double)c.hour / el, c.hour);
                log_counts(c);
            }
        }
        simulation_hour = seg_end + 1u;
    }
    *n_rows = written;
    *stopped = stop ? 1 : 0;
    return EPI_OK;
}

// EventsKafkaProducer::publish_citizen_states_buffer (events_kafka_producer.rs:62-67): serde_json of CitizenStatesAtHr, one line
static int append_citizen_states(epi_engine* e, uint32_t hour, std::FILE* f) {
    const uint32_t cap = epi_capacity(e);
    std::vector<char> state(cap);
    std::vector<int32_t> x(cap), y(cap);
    std::vector<uint32_t> slot(cap);
    uint32_t n = 0;
    const int rc = epi_citizen_states(e, state.data(), x.data(), y.data(), slot.data(), cap, &n);
    if (rc) return rc;
    std::fprintf(f, "{\"hr\":%u,\"citizen_states\":[", hour);
    for (uint32_t k = 0; k < n; ++k)
        std::fprintf(f, "%s{\"citizen_id\":\"00000000-0000-4000-8000-%012x\",\"state\":\"%c\",\"location\":{\"x\":%d,\"y\":%d}}", k ? "," : "", slot[k], state[k], x[k], y[k]);
    std::fprintf(f, "]}\n");
    return std::ferror(f) ? engine_fail(e, EPI_ERR_IO, "cannot write the citizen states file") : EPI_OK;
}

int run_single_engine(epi_engine* e, const epi_config& cfg, RunResult& result, bool log, std::FILE* citizen_states) {
    epi_counts counts_at_hr;
    int rc = epi_counts_at_start(e, &counts_at_hr);
    if (rc) return rc;
    if (log) log_counts(counts_at_hr);
    const auto start_time = std::chrono::steady_clock::now();
    result.rows.assign(cfg.hours > 0 ? cfg.hours : 1u, epi_counts{});
    uint32_t n_rows = 0;
    int stopped = 0;
    if (cfg.hours > 1 && !citizen_states) {  // for simulation_hour in 1..config.get_hours()
        rc = simulate_hours(e, 1, cfg.hours - 1u, true, result.rows.data(), &n_rows, &stopped, log, &start_time);
        if (rc) return rc;
    }
    // publish_citizen_state (allocation_map.rs:122-124): every agent's state after every hour -> the hours are run one at a time
    for (uint32_t hour = 1; citizen_states && hour < cfg.hours && !stopped; ++hour) {
        uint32_t got = 0;
        rc = simulate_hours(e, hour, 1, true, result.rows.data() + n_rows, &got, &stopped, log, &start_time);
        if (rc) return rc;
        n_rows += got;
        rc = append_citizen_states(e, hour, citizen_states);
        if (rc) return rc;
    }
    result.rows.resize(n_rows);
    result.loop_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start_time).count();
    if (log && n_rows) {
        const uint32_t last = result.rows.back().hour;
        std::printf("INFO - Number of iterations: %u, Total Time taken %f seconds\n", last, result.loop_seconds);
        std::printf("INFO - Iterations/sec: %f\n", (double)last / result.loop_seconds);
        std::printf("INFO - Agent-steps/sec: %e\n", (double)last * (double)epi_population(e) / result.loop_seconds);
    }
    static const char* names[3] = {"lockdown", "vaccination", "build_new_hospital"};
    for (const epi_intervention_event& ev : e->events) {
        const char* data = ev.kind == 0 ? (ev.status ? "{\"status\":\"locked_down\"}" : "{\"status\":\"lockdown_revoked\"}") : "{}";
        result.interventions.push_back({ev.hour, names[ev.kind], data});
    }
    return EPI_OK;
}

// ---- config JSON (common::config::Config, common/src/config/mod.rs:44-58) ------------------------------------------
void config_from_value(const JsonValue& root, epi_config& c) {
    std::memset(&c, 0, sizeof(c));
    const JsonValue& pop = root.at("population");
    if (const JsonValue* a = pop.find("Auto")) {
        c.number_of_agents = a->at("number_of_agents").as_u32("number_of_agents");
        c.public_transport_percentage = a->at("public_transport_percentage").as_number("public_transport_percentage");
        c.working_percentage = a->at("working_percentage").as_number("working_percentage");
    } else if (const JsonValue* v = pop.find("Csv")) {  // CsvPopulation { file, cols } (population.rs:30-34)
        const std::string file = v->at("file").as_string("file");
        v->at("cols");  // required by serde; not read by the engine
        if (file.empty() || file.size() >= EPI_PATH_MAX) throw std::runtime_error("population.Csv.file: empty or longer than " + std::to_string(EPI_PATH_MAX - 1) + " bytes");
        std::memcpy(c.population_csv_file, file.c_str(), file.size() + 1);
    } else {
        throw std::runtime_error("unknown variant for `population`, expected `Csv` or `Auto`");
    }
    const JsonValue* dis = root.find("disease");
    if (!dis || dis->kind == JsonValue::Null) throw std::runtime_error("`disease` is required (Config::get_disease unwraps it, common/src/config/mod.rs:83-85)");
    c.regular_transmission_start_day = dis->at("regular_transmission_start_day").as_u32("regular_transmission_start_day");
    c.high_transmission_start_day = dis->at("high_transmission_start_day").as_u32("high_transmission_start_day");
    c.last_day = dis->at("last_day").as_u32("last_day");
    c.asymptomatic_last_day = dis->at("asymptomatic_last_day").as_u32("asymptomatic_last_day");
    c.mild_infected_last_day = dis->at("mild_infected_last_day").as_u32("mild_infected_last_day");
    c.regular_transmission_rate = dis->at("regular_transmission_rate").as_number("regular_transmission_rate");
    c.high_transmission_rate = dis->at("high_transmission_rate").as_number("high_transmission_rate");
    c.death_rate = dis->at("death_rate").as_number("death_rate");
    c.percentage_asymptomatic_population = dis->at("percentage_asymptomatic_population").as_number("percentage_asymptomatic_population");
    c.percentage_severe_infected_population = dis->at("percentage_severe_infected_population").as_number("percentage_severe_infected_population");
    c.exposed_duration = dis->at("exposed_duration").as_u32("exposed_duration");
    c.pre_symptomatic_duration = dis->at("pre_symptomatic_duration").as_u32("pre_symptomatic_duration");
    const JsonValue& geo = root.at("geography_parameters");
    c.grid_size = geo.at("grid_size").as_u32("grid_size");
    c.hospital_beds_percentage = geo.at("hospital_beds_percentage").as_number("hospital_beds_percentage");
    c.hours = root.at("hours").as_u32("hours");
    const JsonValue& ivs = root.at("interventions");
    if (ivs.kind != JsonValue::Array) throw std::runtime_error("`interventions` must be an array");
    for (const JsonValue& iv : ivs.arr) {
        if (iv.kind != JsonValue::Object || iv.obj.size() != 1) throw std::runtime_error("each intervention must be an object with one variant key");
        const std::string& variant = iv.obj[0].first;
        const JsonValue& body = iv.obj[0].second;
        if (variant == "Vaccinate") {
            if (c.n_vaccinations >= EPI_MAX_VACCINATIONS) throw std::runtime_error("too many Vaccinate interventions (max 8)");
            // later entries with the same hour overwrite (vaccination.rs:43-45)
            const uint32_t at_hour = body.at("at_hour").as_u32("at_hour");
            const double percent = body.at("percent").as_number("percent");
            int slot = c.n_vaccinations;
            for (int i = 0; i < c.n_vaccinations; ++i)
                if (c.vaccinate_at_hour[i] == at_hour) slot = i;
            c.vaccinate_at_hour[slot] = at_hour;
            c.vaccinate_percent[slot] = percent;
            if (slot == c.n_vaccinations) c.n_vaccinations++;
        } else if (variant == "Lockdown") {
            if (!c.has_lockdown) {  // .next(): the first one wins (lockdown.rs:34-44)
                c.has_lockdown = 1;
                c.lockdown_at_number_of_infections = body.at("at_number_of_infections").as_u32("at_number_of_infections");
                c.essential_workers_population = body.at("essential_workers_population").as_number("essential_workers_population");
            }
        } else if (variant == "BuildNewHospital") {
            if (!c.has_build_new_hospital) {
                c.has_build_new_hospital = 1;
                c.spread_rate_threshold = body.at("spread_rate_threshold").as_u32("spread_rate_threshold");
            }
        } else {
            throw std::runtime_error("unknown variant `" + variant + "`, expected one of `Vaccinate`, `Lockdown`, `BuildNewHospital`");
        }
    }
    if (const JsonValue* si = root.find("starting_infections")) {
        c.infected_mild_asymptomatic = si->at("infected_mild_asymptomatic").as_u32("infected_mild_asymptomatic");
        c.infected_mild_symptomatic = si->at("infected_mild_symptomatic").as_u32("infected_mild_symptomatic");
        c.infected_severe = si->at("infected_severe").as_u32("infected_severe");
        c.exposed = si->at("exposed").as_u32("exposed");
    } else {
        c.exposed = 1;  // StartingInfections::default (starting_infections.rs:65-69)
    }
}

}  // namespace epi

using namespace epi;

extern "C" {

int epi_config_from_json_string(const char* json_text, epi_config* out) {
    if (!json_text || !out) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    try {
        config_from_value(json_parse(json_text), *out);
    } catch (const std::exception& ex) {
        return engine_fail(nullptr, EPI_ERR_CONFIG, ex.what());
    }
    return EPI_OK;
}

int epi_config_from_json(const char* json_path, epi_config* out) {
    if (!json_path || !out) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    std::string text;
    try {
        text = json_read_file(json_path);
    } catch (const std::exception& ex) {
        return engine_fail(nullptr, EPI_ERR_IO, ex.what());
    }
    return epi_config_from_json_string(text.c_str(), out);
}

int epi_simulate_hours(epi_engine* e, uint32_t first_hour, uint32_t n_hours, int stop_rule, epi_counts* rows_out, uint32_t* n_rows, int* stopped) {
    if (!e || (!rows_out && n_hours) || !n_rows || !stopped) return engine_fail(e, EPI_ERR_ARG, "null argument");
    *n_rows = 0;
    *stopped = 0;
    if (n_hours == 0) return EPI_OK;
    return simulate_hours(e, first_hour, n_hours, stop_rule != 0, rows_out, n_rows, stopped, false, nullptr);
}

int epi_intervention_events(const epi_engine* e, epi_intervention_event* out, uint32_t max_events, uint32_t* n) {
    if (!e || !n) return engine_fail(e, EPI_ERR_ARG, "null argument");
    *n = (uint32_t)e->events.size();
    if (out)
        for (uint32_t i = 0; i < std::min<uint32_t>(max_events, *n); ++i) out[i] = e->events[i];
    return EPI_OK;
}

int epi_run_standalone(const epi_config* cfg, uint64_t seed, int device, const char* output_dir, const char* engine_id, epi_counts* rows_out,
                       uint32_t max_rows, uint32_t* n_rows, double* loop_seconds) {
    return epi_run_standalone_ex(cfg, seed, device, output_dir, engine_id, 0, rows_out, max_rows, n_rows, loop_seconds);
}

int epi_config_citizen_state_messages(const char* json_path, int* on) {
    if (!json_path || !on) return engine_fail(nullptr, EPI_ERR_ARG, "null argument");
    *on = 0;
    try {
        std::ifstream in(json_path);
        if (!in) return engine_fail(nullptr, EPI_ERR_IO, std::string("cannot open ") + json_path);
        std::stringstream ss;
        ss << in.rdbuf();
        const JsonValue root = json_parse(ss.str());
        if (const JsonValue* v = root.find("enable_citizen_state_messages")) *on = v->kind == JsonValue::Bool && v->b ? 1 : 0;
    } catch (const std::exception& ex) {
        return engine_fail(nullptr, EPI_ERR_CONFIG, ex.what());
    }
    return EPI_OK;
}

int epi_run_standalone_ex(const epi_config* cfg, uint64_t seed, int device, const char* output_dir, const char* engine_id, int citizen_state_messages,
                          epi_counts* rows_out, uint32_t max_rows, uint32_t* n_rows, double* loop_seconds) {
    if (!cfg) return engine_fail(nullptr, EPI_ERR_ARG, "null config");
    epi_engine* e = nullptr;
    int rc = epi_create(cfg, seed, device, &e);
    if (rc) return rc;
    RunResult res;
    const bool log = std::getenv("EPI_LOG") != nullptr;
    // the listeners take their file name when they are created (epidemiology_simulation.rs:141-150), i.e. before the hour loop
    std::string base;
    std::FILE* states = nullptr;
    if (output_dir) {
        try {
            base = output_file_format(output_dir, engine_id ? engine_id : "0");
        } catch (const std::exception& ex) {
            epi_destroy(e);
            return engine_fail(nullptr, EPI_ERR_IO, ex.what());
        }
        if (citizen_state_messages) {
            states = std::fopen((base + "_citizen_states.jsonl").c_str(), "w");
            if (!states) {
                epi_destroy(e);
                return engine_fail(nullptr, EPI_ERR_IO, "cannot create " + base + "_citizen_states.jsonl");
            }
        }
    }
    rc = run_single_engine(e, *cfg, res, log, states);
    if (states) {
        if (!rc) std::fputs("{\"simulation_ended\": true}\n", states);  // events_kafka_producer.rs:77-88
        std::fclose(states);
    }
    if (rc) {
        set_global_error(e->err);
        epi_destroy(e);
        return rc;
    }
    epi_destroy(e);
    if (output_dir) {
        try {
            Listeners l;
            l.counts = res.rows;
            l.interventions = res.interventions;
            l.simulation_ended(base);
        } catch (const std::exception& ex) {
            return engine_fail(nullptr, EPI_ERR_IO, ex.what());
        }
    }
    if (n_rows) *n_rows = (uint32_t)res.rows.size();
    if (loop_seconds) *loop_seconds = res.loop_seconds;
    if (rows_out)
        for (uint32_t i = 0; i < std::min<uint32_t>(max_rows, (uint32_t)res.rows.size()); ++i) rows_out[i] = res.rows[i];
    return EPI_OK;
}

}  // extern "C"
