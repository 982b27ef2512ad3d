// engine-app: the command line of the reference's engine binary (engine-app/src/main.rs:60-177) on top of the C ABI.
//
//   engine-app [-c FILE] [-m standalone|mpi|kafka] [-i ID] [-t THREADS] [-o OUTPUT_DIR] [--seed S] [--device D]
//
// Same flags and defaults as the reference (clap derive, main.rs:60-87): --config, --mode (default standalone), --id,
// --threads (default 4; accepted and ignored: the agent step runs on the GPU), --output-dir (default /tmp).  New:
// --seed (Philox key; the reference is unseeded) and --device.
//   standalone  Config::read(FILE or config/default.json) -> EngineApp::start_standalone (engine_app.rs:89-106); writes
//               <OUTPUT_DIR>/output/simulation_0_<UTC>.csv and ..._interventions.json.
//   mpi         one region engine per GPU: re-launches `python -m epirust_b200.engine_app` under torch.distributed.run with
//               one process per region (the reference is started with `mpirun -n <regions>`, main.rs:131-166).
//   kafka       not available (needs a Kafka broker; the orchestrator's tick barrier is replaced by the collective).
#include <unistd.h>

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
    const int rc = epi_run_standalone(&cfg, a.seed, a.device, a.output_dir.c_str(), "0", nullptr, 0, &n_rows, &secs);
    if (rc != EPI_OK) {
        std::fprintf(stderr, "engine-app: error %d: %s\n", rc, epi_last_error(nullptr));
        return 1;
    }
    return 0;
}

int run_mpi(const Args& a, const char* argv0) {
    // one process per region under torch.distributed.run; the Python launcher mirrors main.rs:131-166
    const std::string config_file = a.has_config ? a.config : "engine/config/simulation.json";  // main.rs:143-144
    std::string exe = argv0;
    const size_t slash = exe.rfind('/');
    const std::string pkg_dir = slash == std::string::npos ? "." : exe.substr(0, slash);
    const std::string root = pkg_dir + "/..";
    const char* old = std::getenv("PYTHONPATH");
    setenv("PYTHONPATH", old ? (root + ":" + old).c_str() : root.c_str(), 1);
    std::vector<std::string> cmd = {"python", "-m", "epirust_b200.engine_app", "--launch", "-m", "mpi", "-c", config_file, "-o", a.output_dir,
                                    "--seed", std::to_string(a.seed), "-t", std::to_string(a.threads)};
    if (a.nproc > 0) { cmd.push_back("--nproc"); cmd.push_back(std::to_string(a.nproc)); }
    std::vector<char*> av;
    for (auto& s : cmd) av.push_back(const_cast<char*>(s.c_str()));
    av.push_back(nullptr);
    execvp(av[0], av.data());
    std::perror("engine-app: cannot start the multi-region launcher (python)");
    return 1;
}

}  // namespace

int main(int argc, char** argv) {
    Args a;
    const int pr = parse(argc, argv, a);
    if (pr) return pr == 1 ? 0 : 2;
    if (a.mode == "mpi") return run_mpi(a, argv[0]);  // every region process prints its own mode line
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
