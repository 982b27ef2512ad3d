"""engine-app for multi-region runs: one region engine per GPU, one process per GPU (torch.distributed, NCCL).

    python -m epirust_b200.engine_app --launch -m mpi -c engine/config/simulation.json -o /tmp     # starts torchrun
    python -m torch.distributed.run --nproc-per-node R -m epirust_b200.engine_app -m mpi -c ...    # what --launch runs

Mirror of the reference's `engine-app -m mpi` (engine-app/src/main.rs:131-166): rank r reads the multi-region
`Configuration` (common/src/config/configuration.rs:221-315), validates it, takes `engine_configs[r]` and runs
`Epidemiology::run_multi_engine` (engine/src/epidemiology_simulation.rs:276-547) with the MPI transport replaced by the
on-device pack / all-to-allv / unpack of `epirust_b200.multi`.  Outputs per region, like the reference:
`<out>/output/simulation_<engine_id>_<UTC>.csv`, `..._interventions.json`, `..._outgoing_travels.csv`
(listeners/csv_service.rs:44-71, intervention_reporter.rs:28-63, travel_counter.rs:27-92).
`-m standalone` runs `EngineApp::start_standalone` through the C ABI (same as the C++ `engine-app` binary).
"""
import argparse
import datetime
import json
import math
import os
import sys
import time

import numpy as np

TRANSPORT_AREA_RELATIVE_SIZE = 0.2  # engine/src/models/constants.rs:23
COUNT_HEADER = "hour,susceptible,exposed,infected,hospitalized,recovered,deceased"
INTERVENTION_NAMES = ("lockdown", "vaccination", "build_new_hospital")


class ConfigError(ValueError):
    pass


def read_configuration(path):
    """Configuration::read (configuration.rs:237-247): {engine_configs: [{engine_id, config}], travel_plan}."""
    with open(path) as f:
        doc = json.load(f)
    for key in ("engine_configs", "travel_plan"):
        if key not in doc:
            raise ConfigError(f"missing field `{key}`")
    engines = doc["engine_configs"]
    tp = doc["travel_plan"]
    regions = list(tp["regions"])
    ids = [e["engine_id"] for e in engines]
    # TravelPlanConfig::validate_regions (travel_plan_config.rs:60-62)
    if len(ids) != len(regions) or not all(i in regions for i in ids):
        raise ConfigError("Engine names should match regions in travel plan")
    R = len(regions)

    def matrix(block, name):
        if not block.get("enabled", False):
            return None
        m = block.get("matrix")
        if m is None:
            raise ConfigError(f"travel_plan.{name}.matrix is required when enabled")
        a = np.asarray(m, dtype=np.int64)
        if a.shape != (R, R) or (a < 0).any():
            raise ConfigError(f"travel_plan.{name}.matrix must be {R}x{R} non-negative")
        return a.astype(np.uint32)

    plan = dict(n_regions=R, regions=regions, migration=matrix(tp["migration"], "migration"), commute=matrix(tp["commute"], "commute"),
                start_migration_hour=int(tp["migration"].get("start_migration_hour", 0)), end_migration_hour=int(tp["migration"].get("end_migration_hour", 0)))
    return engines, plan


def validate_configuration(engines, plan):
    """Configuration::validate (configuration.rs:249-307): transport capacity and the grid/population ratio per engine."""
    regions = plan["regions"]
    for e in engines:
        cfg = e["config"]
        pop = cfg["population"].get("Auto")
        grid_size = int(cfg["geography_parameters"]["grid_size"])
        n_agents = int(pop["number_of_agents"]) if pop else 0
        pt = float(pop["public_transport_percentage"]) if pop else 0.0
        r = regions.index(e["engine_id"])
        total_population = n_agents
        transport_cells = (math.ceil(grid_size * TRANSPORT_AREA_RELATIVE_SIZE) - 1) * grid_size
        if plan["commute"] is not None:
            incoming, outgoing = int(plan["commute"][:, r].sum()), int(plan["commute"][r, :].sum())
            if math.ceil(n_agents * pt) - outgoing + incoming > transport_cells:
                raise ConfigError(f"For engine id - {e['engine_id']}, Incoming commuters are more than engine transport capacity")
            total_population += incoming - outgoing
        if plan["migration"] is not None:
            total_population += int(plan["migration"][:, r].sum()) - int(plan["migration"][r, :].sum())
        if total_population <= 0 or (grid_size * grid_size) // total_population < 3:
            raise ConfigError(f"{e['engine_id']}: Not enough space to accumulate the migrators/commuters")


def arrival_capacity(plan, r, hours):
    """Agent slots to reserve for arrivals of region r: every commuter of a day plus the migrators of the whole window."""
    extra = 0
    if plan["commute"] is not None:
        extra += int(plan["commute"][:, r].sum())
    if plan["migration"] is not None:
        last = min(plan["end_migration_hour"], hours)
        days = max(0, (last - plan["start_migration_hour"]) // 24 + 1)
        extra += int(plan["migration"][:, r].sum()) * days
    return extra + extra // 8 + 1024


def segment_capacity(plan):
    """Records one (source, destination) segment must hold: twice the largest planned pair (migrator counts are binomial) + slack."""
    most = 0
    for k in ("migration", "commute"):
        if plan[k] is not None:
            most = max(most, int(plan[k].max()))
    return 2 * most + 4096


def output_file_format(output_dir, engine_id):
    """utils/util.rs:31-43"""
    d = os.path.join(output_dir, "output")
    os.makedirs(d, exist_ok=True)
    stamp = datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%S")
    return os.path.join(d, f"simulation_{engine_id}_{stamp}")


def write_outputs(base, rows, events, travels=None):
    with open(base + ".csv", "w") as f:
        f.write(COUNT_HEADER + "\n")
        for r in rows:
            f.write(",".join(str(int(v)) for v in r) + "\n")
    reports = []
    for hour, kind, status in events:
        data = {"status": "locked_down" if status else "lockdown_revoked"} if kind == 0 else {}
        reports.append({"hour": int(hour), "intervention": INTERVENTION_NAMES[kind], "data": data})
    with open(base + "_interventions.json", "w") as f:
        json.dump(reports, f, separators=(",", ":"))
    if travels is not None:
        with open(base + "_outgoing_travels.csv", "w") as f:
            f.write("hr,destination,susceptible,exposed,infected,recovered\n")
            for t in travels:
                f.write(",".join(str(v) for v in t) + "\n")


def run_region(args):
    """One rank = one region (main.rs:131-166 + run_multi_engine)."""
    import torch
    import torch.distributed as dist

    from . import _ffi
    from .engine import Engine, config_from_json_string
    from .multi import DistExchange, MultiRegion

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    print("in multi-engine mode", flush=True)
    engines, plan = read_configuration(args.config or "engine/config/simulation.json")
    validate_configuration(engines, plan)
    R = plan["n_regions"]
    if world > R:
        raise SystemExit(f"{world} processes for {R} regions")
    if world < R:  # fewer GPUs than regions: run the first `world` regions of the travel plan
        for k in ("migration", "commute"):
            if plan[k] is not None:
                plan[k] = np.ascontiguousarray(plan[k][:world, :world])
        plan["regions"] = plan["regions"][:world]
        plan["n_regions"] = R = world
    if not torch.cuda.is_available():
        raise SystemExit("engine-app needs a CUDA device: epirust_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # rank == index in travel_plan.regions (MpiTransport::new maps region names to ranks by that index, mpi_transport.rs:44-52);
    # the reference takes engine_configs[rank] (main.rs:146-147), which is the same engine when both lists are in the same order
    region = rank
    engine_id = plan["regions"][region]
    me = next(e for e in engines if e["engine_id"] == engine_id)
    cfg = config_from_json_string(json.dumps(me["config"]))
    hours = int(cfg.hours)
    eng = Engine(cfg, seed=args.seed + region, device=local, region=region, plan=plan, extra_capacity=arrival_capacity(plan, region, hours))
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    travels = []

    def on_outgoing(hour, kind, send_buf, counts):  # TravelCounter::outgoing_migrators_added (travel_counter.rs:84-87)
        if kind != _ffi.TRAVEL_MIGRATE or int(counts.sum()) == 0:
            return
        for dest, c in enumerate(counts):
            c = int(c)
            if dest == region:
                continue
            state = (send_buf[dest, 1:1 + c, 0] & 7).cpu().numpy() if c else np.zeros(0, np.int64)
            travels.append((hour, plan["regions"][dest], int((state == 0).sum()), int((state == 1).sum()), int((state == 2).sum()), int((state == 3).sum())))

    start = time.time()
    with torch.cuda.stream(stream):
        runner = MultiRegion([eng], plan, exchange=DistExchange(torch.device("cuda", local)), stride_records=segment_capacity(plan),
                             on_outgoing=on_outgoing)
        rows = np.zeros((1, max(hours - 1, 0), 7), np.uint32)
        done = 0
        while done < hours - 1:  # for simulation_hour in 1..config.get_hours()
            n = min(240, hours - 1 - done)
            runner.run(1 + done, n, rows_out=rows[:, done:done + n])
            done += n
            c = rows[0, done - 1]
            print(f"INFO - [{engine_id}] hour {int(c[0])}: S: {c[1]}, E:{c[2]}, I: {c[3]}, H: {c[4]}, R: {c[5]}, D: {c[6]}; "
                  f"Throughput: {done / (time.time() - start):.2f} iterations/sec", flush=True)
        eng.sync()
    elapsed = time.time() - start
    print(f"INFO - [{engine_id}] Number of iterations: {hours - 1}, Total Time taken {elapsed:.3f} seconds; Iterations/sec: {(hours - 1) / elapsed:.2f}", flush=True)
    write_outputs(output_file_format(args.output_dir, engine_id), rows[0], eng.intervention_events(), travels)
    eng.close()
    dist.destroy_process_group()


def run_standalone(args):
    from .engine import config_from_json, run_standalone as run

    cfg = config_from_json(args.config or "config/default.json")
    os.environ.setdefault("EPI_LOG", "1")
    run(cfg, seed=args.seed, device=args.device, output_dir=args.output_dir, engine_id="0")


def launch(args):
    """Start one process per region under torch.distributed.run (the reference: `mpirun -n <regions> engine-app -m mpi`)."""
    engines, plan = read_configuration(args.config or "engine/config/simulation.json")
    validate_configuration(engines, plan)
    n = args.nproc or plan["n_regions"]
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           "-m", "epirust_b200.engine_app", "-m", "mpi", "-o", args.output_dir, "--seed", str(args.seed)]
    if args.config:
        cmd += ["-c", args.config]
    os.execv(sys.executable, cmd)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="engine-app", description="EpiRust engine on B200 (multi-region launcher)")
    ap.add_argument("-c", "--config", metavar="FILE", help="Use a config file to run the simulation")
    ap.add_argument("-m", "--mode", default="standalone", choices=["kafka", "mpi", "standalone"])
    ap.add_argument("-i", "--id", help="An identifier for the engine")
    ap.add_argument("-t", "--threads", type=int, default=4, help="accepted for compatibility; the agent step runs on the GPU")
    ap.add_argument("-o", "--output-dir", default="/tmp")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--nproc", type=int, default=0, help="mpi mode: regions (= GPUs = processes) to run")
    ap.add_argument("--launch", action="store_true", help="mpi mode: start the per-region processes with torch.distributed.run")
    args = ap.parse_args(argv)
    if not args.launch:
        print({"mpi": "MPI", "kafka": "Kafka", "standalone": "Standalone"}[args.mode], flush=True)  # println!("{:?}", args.mode), main.rs:104
    if args.mode == "kafka":
        raise SystemExit("kafka mode is not available in the B200 build (needs a Kafka broker); use -m mpi")
    if args.mode == "standalone":
        return run_standalone(args)
    if args.launch:
        return launch(args)
    return run_region(args)


if __name__ == "__main__":
    main()
