// Host side of the hour loop: the mirror of the reference's interventions, listeners and Epidemiology::run_single_engine,
// driving the HBM-resident engine through the C ABI.
#pragma once
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/epi.h"

namespace epi {

// engine/src/interventions/lockdown.rs
class LockdownIntervention {
  public:
    explicit LockdownIntervention(const epi_config& cfg)
        : has_config_(cfg.has_lockdown != 0), at_number_of_infections_(cfg.lockdown_at_number_of_infections),
          essential_workers_population_(cfg.essential_workers_population) {}
    bool should_apply(const epi_counts& c) const { return !is_locked_down_ && c.hour % 24u == 0 && has_config_ && c.infected > at_number_of_infections_; }
    bool should_unlock(const epi_counts& c) const { return is_locked_down_ && c.hour == unlock_hour(); }
    uint32_t unlock_hour() const { return zero_infection_hour + 21u * 24u; }  // round(QUARANTINE_DAYS * 1.5) * HOURS_IN_A_DAY
    void set_zero_infection_hour(uint32_t h) { if (zero_infection_hour == 0) zero_infection_hour = h; }
    bool apply() { if (!has_config_) return false; is_locked_down_ = true; return true; }
    void unapply() { is_locked_down_ = false; zero_infection_hour = 0; }
    bool is_locked_down() const { return is_locked_down_; }
    const char* name() const { return "lockdown"; }
    const char* json_data() const { return is_locked_down_ ? "{\"status\":\"locked_down\"}" : "{\"status\":\"lockdown_revoked\"}"; }
    uint32_t zero_infection_hour = 0;

  private:
    bool is_locked_down_ = false, has_config_;
    uint32_t at_number_of_infections_;
    double essential_workers_population_;
};

// engine/src/interventions/hospital.rs
class BuildNewHospital {
  public:
    explicit BuildNewHospital(const epi_config& cfg) : has_config_(cfg.has_build_new_hospital != 0), spread_rate_threshold_(cfg.spread_rate_threshold) {}
    bool should_apply(const epi_counts& c) const { return !has_applied_ && c.hour % 24u == 0 && has_config_ && new_infections_in_a_day_ >= spread_rate_threshold_; }
    void apply() { has_applied_ = true; }
    bool has_applied() const { return has_applied_; }
    void counts_updated(const epi_counts& c) {
        if (c.hour % 24u == 0) new_infections_in_a_day_ = c.infected > new_infections_in_a_day_ ? c.infected - new_infections_in_a_day_ : 0;  // saturating_sub (sic)
    }
    const char* name() const { return "build_new_hospital"; }
    const char* json_data() const { return "{}"; }

  private:
    bool has_config_, has_applied_ = false;
    uint32_t spread_rate_threshold_, new_infections_in_a_day_ = 0;
};

// engine/src/interventions/vaccination.rs
class VaccinateIntervention {
  public:
    explicit VaccinateIntervention(const epi_config& cfg) {
        for (int i = 0; i < cfg.n_vaccinations; ++i) by_hour_[cfg.vaccinate_at_hour[i]] = cfg.vaccinate_percent[i];
    }
    const double* get_vaccination_percentage(const epi_counts& c) const {
        auto it = by_hour_.find(c.hour);
        return it == by_hour_.end() ? nullptr : &it->second;
    }
    // first configured hour >= h, or UINT32_MAX
    uint32_t next_hour(uint32_t h) const {
        auto it = by_hour_.lower_bound(h);
        return it == by_hour_.end() ? 0xFFFFFFFFu : it->first;
    }
    const char* name() const { return "vaccination"; }
    const char* json_data() const { return "{}"; }

  private:
    std::map<uint32_t, double> by_hour_;
};

// engine/src/interventions/interventions.rs: the three state machines of one engine
struct Interventions {
    explicit Interventions(const epi_config& cfg) : vaccinate(cfg), lockdown(cfg), build_new_hospital(cfg) {}
    VaccinateIntervention vaccinate;
    LockdownIntervention lockdown;
    BuildNewHospital build_new_hospital;
};

struct InterventionReport {  // listeners/intervention_reporter.rs:28-33
    uint32_t hour;
    std::string intervention, data;
};

// CsvListener + InterventionReporter (listeners/csv_service.rs, listeners/intervention_reporter.rs)
struct Listeners {
    std::vector<epi_counts> counts;
    std::vector<InterventionReport> interventions;
    // writes <base>.csv and <base>_interventions.json
    void simulation_ended(const std::string& base) const;
};

// utils/util.rs:31-43: <output_dir>/output/simulation_<engine_id>_<UTC yyyy-mm-ddThh:mm:ss>
std::string output_file_format(const std::string& output_dir, const std::string& engine_id);

struct RunResult {
    std::vector<epi_counts> rows;
    std::vector<InterventionReport> interventions;
    double loop_seconds = 0.0;
    std::string csv_path, interventions_path;
};

struct JsonValue;
// common::config::Config from its JSON object (common/src/config/mod.rs:44-58); throws std::runtime_error
void config_from_value(const JsonValue& root, epi_config& c);

// process_interventions on one Counts row (host decisions + the sweep kernels)
int process_interventions(epi_engine* e, const epi_counts& c, bool log);

// Epidemiology::run_single_engine on an existing engine (epidemiology_simulation.rs:211-274).  Returns EPI_* code.
// citizen_states: when not null, the run goes hour by hour and appends one CitizenStatesAtHr JSON line per hour to it
// (Config.enable_citizen_state_messages; listeners/events_kafka_producer.rs:62-100).
int run_single_engine(epi_engine* e, const epi_config& cfg, RunResult& result, bool log, std::FILE* citizen_states = nullptr);

}  // namespace epi
