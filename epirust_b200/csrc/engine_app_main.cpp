// engine-app: the command line of the reference's engine binary (engine-app/src/main.rs:60-177) on top of the C ABI.
//
//   engine-app [-c FILE] [-m standalone|mpi|kafka] [-i ID] [-t THREADS] [-o OUTPUT_DIR] [--seed S] [--device D]
//
// Same flags and defaults as the reference (clap derive, main.rs:60-87): --config, --mode (default standalone), --id,
// --threads (default 4; accepted and ignored: the agent step runs on the GPU), --output-dir (default /tmp).  New:
// --seed (Philox key; the reference is unseeded) and --device.
//   standalone  Config::read(FILE or config/default.json) -> EngineApp::start_standalone (engine_app.rs:89-106); writes
//               <OUTPUT_DIR>/output/simulation_0_<UTC>.csv and ..._interventions.json.
//   mpi         one region engine per GPU, one process per region (the reference is started with `mpirun -n <regions>`,
//               main.rs:131-166): forks the region processes itself, or is one rank when RANK / WORLD_SIZE are set; the ranks
//               exchange travellers over NCCL (epi_run_region).  No Python anywhere on this path.
//   kafka       not available (needs a Kafka broker; the orchestrator's tick barrier is replaced by the collective).
#include <signal.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "epi.h"

namespace {

struct Args {
    std::string config, mode = "standalone", id, output_dir = "/tmp";
    bool has_config = false, has_id = false;
    unsigned threads = 4;
    unsigned long long seed = 1;
    int device = 0;
    int nproc = 0;  // mpi mode: processes to launch (0 = number of regions in the config)
    bool terminate_when_clear = false;  // mpi mode: the orchestrator's global termination rule (ticks.rs:175-180)
};

void usage(FILE* f) {
    std::fprintf(f,
                 "Usage: engine-app [OPTIONS]\n\n"
                 "Options:\n"
                 "  -c, --config <FILE>            Use a config file to run the simulation\n"
                 "  -m, --mode <MODE>              start the engine with a particular implementation- Kafka or MPI [default: standalone]\n"
                 "                                 [possible values: kafka, mpi, standalone]\n"
                 "  -i, --id <ID>                  An identifier for the engine\n"
                 "  -t, --threads <THREADS>        Number of parallel threads for data parallelization [default: 4] (ignored: GPU engine)\n"
                 "  -o, --output-dir <OUTPUT_DIR>  Output directory [default: /tmp]\n"
                 "      --seed <SEED>              Philox key of the run [default: 1]\n"
                 "      --device <DEVICE>          CUDA device of a standalone run [default: 0]\n"
                 "      --nproc <N>                mpi mode: regions (= GPUs = processes) to run [default: all regions of the config]\n"
                 "      --terminate-when-clear     mpi mode: stop when no region has exposed / infected / hospitalized agents (the\n"
                 "                                 orchestrator's rule in Kafka mode; the reference's MPI mode runs to `hours`)\n"
                 "  -h, --help                     Print help\n"
                 "  -V, --version                  Print version\n");
}

// returns 0 ok, 1 exit success (help / version), 2 usage error
int parse(int argc, char** argv, Args& a) {
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i], v;
        bool has_v = false;
        const size_t eq = k.find('=');
        if (k.rfind("--", 0) == 0 && eq != std::string::npos) { v = k.substr(eq + 1); k = k.substr(0, eq); has_v = true; }
        auto value = [&](std::string& out) -> bool {
            if (has_v) { out = v; return true; }
            if (i + 1 >= argc) { std::fprintf(stderr, "error: a value is required for '%s' but none was supplied\n", k.c_str()); return false; }
            out = argv[++i];
            return true;
        };
        std::string s;
        if (k == "-h" || k == "--help") { usage(stdout); return 1; }
        if (k == "-V" || k == "--version") { std::printf("engine-app (%s)\n", epi_version()); return 1; }
        if (k == "-c" || k == "--config") { if (!value(a.config)) return 2; a.has_config = true; }
        else if (k == "-m" || k == "--mode") {
            if (!value(a.mode)) return 2;
            if (a.mode != "kafka" && a.mode != "mpi" && a.mode != "standalone") {
                std::fprintf(stderr, "error: invalid value '%s' for '--mode <MODE>'\n  [possible values: kafka, mpi, standalone]\n", a.mode.c_str());
                return 2;
            }
        } else if (k == "-i" || k == "--id") { if (!value(a.id)) return 2; a.has_id = true; }
        else if (k == "-t" || k == "--threads") { if (!value(s)) return 2; a.threads = (unsigned)std::strtoul(s.c_str(), nullptr, 10); }
        else if (k == "-o" || k == "--output-dir") { if (!value(a.output_dir)) return 2; }
        else if (k == "--seed") { if (!value(s)) return 2; a.seed = std::strtoull(s.c_str(), nullptr, 10); }
        else if (k == "--device") { if (!value(s)) return 2; a.device = std::atoi(s.c_str()); }
        else if (k == "--nproc") { if (!value(s)) return 2; a.nproc = std::atoi(s.c_str()); }
        else if (k == "--terminate-when-clear") a.terminate_when_clear = true;
        else { std::fprintf(stderr, "error: unexpected argument '%s' found\n\n", k.c_str()); usage(stderr); return 2; }
    }
    return 0;
}

int run_standalone(const Args& a) {
    const std::string config_file = a.has_config ? a.config : "config/default.json";  // main.rs:168-169
    epi_config cfg;
    if (epi_config_from_json(config_file.c_str(), &cfg) != EPI_OK) {
        std::fprintf(stderr, "Failed to read config file: %s\n", epi_last_error(nullptr));
        return 1;
    }
    setenv("EPI_LOG", "1", 0);  // the reference logs to stdout (log4rs console appender)
    uint32_t n_rows = 0;
    double secs = 0.0;
    // EngineApp::start_standalone always names the engine "0" (engine_app.rs:30,101)
    int citizen_states = 0;  // Config.enable_citizen_state_messages (common/src/config/mod.rs:54-55)
    if (epi_config_citizen_state_messages(config_file.c_str(), &citizen_states) != EPI_OK) {
        std::fprintf(stderr, "engine-app: %s\n", epi_last_error(nullptr));
        return 1;
    }
    const int rc = epi_run_standalone_ex(&cfg, a.seed, a.device, a.output_dir.c_str(), "0", citizen_states, nullptr, 0, &n_rows, &secs);
    if (rc != EPI_OK) {
        std::fprintf(stderr, "engine-app: error %d: %s\n", rc, epi_last_error(nullptr));
        return 1;
    }
    return 0;
}

// ---- -m mpi: one process per region, one region per GPU ---------------------------------------------------------------
// The reference is started as `mpirun -n <regions> engine-app -m mpi` and its ranks meet through MPI (main.rs:131-147).  Here
// the ranks meet through an NCCL communicator whose unique id travels in a file:
//   * started under a launcher that sets RANK and WORLD_SIZE (torchrun, mpirun wrappers, SLURM scripts): this process is that
//     rank; the id file is $EPI_COMM_ID_FILE (default /tmp/epi_comm_<MASTER_PORT>.id), written by rank 0;
//   * started plainly: engine-app forks one child per region (before any CUDA call) and waits for them.
// Rank r runs region r = travel_plan.regions[r] on GPU r % device_count.  `-i/--id` is accepted and unused like in the
// reference's MPI arm (the engine id comes from the config, main.rs:146-151; only Kafka mode reads --id).
bool write_id_file(const std::string& path) {
    unsigned char id[EPI_COMM_ID_BYTES];
    if (epi_comm_unique_id(id) != EPI_OK) { std::fprintf(stderr, "engine-app: %s\n", epi_last_error(nullptr)); return false; }
    const std::string tmp = path + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) { std::perror("engine-app: cannot write the communicator id file"); if (f) std::fclose(f); return false; }
    std::fclose(f);
    return std::rename(tmp.c_str(), path.c_str()) == 0;
}

bool read_id_file(const std::string& path, unsigned char* id) {
    for (int tries = 0; tries < 6000; ++tries) {  // up to 10 minutes: rank 0 may still be paging the CUDA libraries in
        FILE* f = std::fopen(path.c_str(), "rb");
        if (f) {
            const size_t n = std::fread(id, 1, EPI_COMM_ID_BYTES, f);
            std::fclose(f);
            if (n == EPI_COMM_ID_BYTES) return true;
        }
        usleep(100 * 1000);
    }
    std::fprintf(stderr, "engine-app: timed out waiting for the communicator id file %s\n", path.c_str());
    return false;
}

int run_rank(const Args& a, const epi_configuration* cfg, int rank, int world, const std::string& id_file) {
    std::printf("MPI\n");  // println!("{:?}", args.mode), main.rs:104 -- every rank prints it
    std::fflush(stdout);
    unsigned char id[EPI_COMM_ID_BYTES];
    if (rank == 0 && !write_id_file(id_file)) return 1;
    if (!read_id_file(id_file, id)) return 1;
    const int n_dev = epi_device_count();
    if (n_dev <= 0) { std::fprintf(stderr, "engine-app: no CUDA device (there is no CPU fallback)\n"); return 1; }
    const int rc = epi_run_region(cfg, rank, world, id, a.seed, rank % n_dev, a.output_dir.c_str(), a.terminate_when_clear ? 1 : 0, nullptr, 0, nullptr, nullptr);
    if (rc != EPI_OK) {
        std::fprintf(stderr, "engine-app: rank %d: error %d: %s\n", rank, rc, epi_last_error(nullptr));
        return 1;
    }
    return 0;
}

int run_mpi(const Args& a) {
    const std::string config_file = a.has_config ? a.config : "engine/config/simulation.json";  // main.rs:143-144
    epi_configuration* cfg = nullptr;
    if (epi_configuration_read(config_file.c_str(), &cfg) != EPI_OK) {  // Configuration::read(..).expect + config.validate()
        std::fprintf(stderr, "Error while reading config: %s\n", epi_last_error(nullptr));
        return 1;
    }
    const int regions = epi_configuration_regions(cfg);
    const char* env_rank = std::getenv("RANK");
    const char* env_world = std::getenv("WORLD_SIZE");
    if (env_rank && env_world) {
        const int rank = std::atoi(env_rank), world = std::atoi(env_world);
        if (world < 1 || world > regions || rank < 0 || rank >= world) {
            std::fprintf(stderr, "engine-app: RANK=%d WORLD_SIZE=%d do not fit the %d regions of %s\n", rank, world, regions, config_file.c_str());
            return 1;
        }
        const char* f = std::getenv("EPI_COMM_ID_FILE");
        const char* port = std::getenv("MASTER_PORT");
        const std::string id_file = f ? f : std::string("/tmp/epi_comm_") + (port ? port : "0") + ".id";
        const int rc = run_rank(a, cfg, rank, world, id_file);
        if (rank == 0) std::remove(id_file.c_str());
        return rc;
    }
    const int world = a.nproc > 0 ? std::min(a.nproc, regions) : regions;
    const std::string id_file = "/tmp/epi_comm_" + std::to_string((long)getpid()) + ".id";
    std::remove(id_file.c_str());
    std::vector<pid_t> kids;
    for (int rank = 0; rank < world; ++rank) {
        std::fflush(stdout);
        std::fflush(stderr);
        const pid_t pid = fork();  // before any CUDA call in this process: each child initialises CUDA for itself
        if (pid < 0) { std::perror("engine-app: fork"); break; }
        if (pid == 0) _exit(run_rank(a, cfg, rank, world, id_file));
        kids.push_back(pid);
    }
    int failed = (int)kids.size() != world;
    size_t left = kids.size();
    while (left) {
        int status = 0;
        const pid_t pid = wait(&status);
        if (pid < 0) break;
        --left;
        if (!(WIFEXITED(status) && WEXITSTATUS(status) == 0) && !failed) {
            failed = 1;  // a rank died: the others would wait for it in the next collective
            for (pid_t k : kids)
                if (k != pid) kill(k, SIGTERM);
        }
    }
    std::remove(id_file.c_str());
    return failed;
}

}  // namespace

int main(int argc, char** argv) {
    Args a;
    const int pr = parse(argc, argv, a);
    if (pr) return pr == 1 ? 0 : 2;
    if (a.mode == "mpi") return run_mpi(a);  // every region process prints its own mode line
    // println!("{:?}", args.mode) (main.rs:104)
    std::printf("%s\n", a.mode == "kafka" ? "Kafka" : "Standalone");
    std::fflush(stdout);
    if (a.mode == "kafka") {
        std::fprintf(stderr,
                     "engine-app: kafka mode is not available in the B200 build: it needs a Kafka broker and the orchestrator process.\n"
                     "Use -m mpi for multi-region runs (one region per GPU; the tick barrier is the traveller-exchange collective).\n");
        return 2;
    }
    return run_standalone(a);
}
