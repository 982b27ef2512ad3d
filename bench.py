#!/usr/bin/env python
"""bench.py -- agent-steps/s of the per-hour agent step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 10m|1m|2m|20m]

A "step" is ONE SIMULATED DAY (24 simulated hours: the full daily routine of every agent of the region, i.e.
18 active-hour passes + the sleep-hour pass) through the C ABI of include/epi.h.  `value` = agents x simulated hours
/ second with agents and grid resident in HBM; `e2e` = the same through the C ABI from HOST buffers (population
uploaded from host memory inside the timed region, Counts rows read back to the host every simulated day).
One process per GPU; with N > 1 every rank runs one region of the same size (weak scaling).

--impl reference times the CPU restatement of the reference algorithm (oracle/, STREAM mode: hash map, parallel phase A,
sequential phase B) on the host cores, on a bounded sample of the same workload.  The Rust reference itself cannot be
built in this image (no cargo/rustc), so kind = "port".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> (agents, grid_size, exposed, interventions)   
WORKLOADS = {
    "1m": dict(n_agents=1_000_000, grid_size=2500, exposed=1000),
    "2m": dict(n_agents=2_000_000, grid_size=3550, exposed=2000),
    "10m": dict(n_agents=10_000_000, grid_size=7910, exposed=10_000, lockdown=(100_000, 0.1), hospital=10_000, vaccinate=((240, 0.2),)),
    "20m": dict(n_agents=20_000_000, grid_size=11_180, exposed=20_000),
}
WORKLOAD_NAMES = {
    "1m": "BASELINE config #2: single region, 1M synthetic agents, G=2500, no interventions",
    "2m": "BASELINE config #4 region: 2M synthetic agents, G=3550",
    "10m": "BASELINE config #3: single region, 10M synthetic agents, G=7910, lockdown + hospital build-up + vaccination",
    "20m": "BASELINE config #5 region: 20M synthetic agents, G=11180",
}
ACTIVE_BYTES, SLEEP_BYTES = 82.0, 8.0  # algorithmic bytes per agent-step (rho = 0.16)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock, power and throttle reasons of the benched GPU DURING the timed region: NVML polled from a thread every
    ~2 ms (the timed region of a default run is a fraction of a second, too short for `nvidia-smi -lms`)."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.rows, self.t, self.h, self.nv, self.err = index, [], None, None, None, None
        self.stop_flag = threading.Event()
        try:
            import pynvml as nv
            import torch

            nv.nvmlInit()
            try:  # CUDA_VISIBLE_DEVICES may renumber: find the device by UUID
                self.h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(index).uuid)).encode())
            except Exception:
                self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.nv = nv
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as ex:  # noqa: BLE001
            self.err = f"NVML unavailable: {ex}"

    def _reasons(self):
        nv = self.nv
        for name in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            f = getattr(nv, name, None)
            if f is not None:
                try:
                    return int(f(self.h))
                except Exception:  # noqa: BLE001
                    continue
        return 0

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), nv.nvmlDeviceGetPowerUsage(self.h) / 1e3, self._reasons()))
            except Exception as ex:  # noqa: BLE001
                self.err = str(ex)
                return
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "samples": 0, "reasons": [self.err or "not sampled"]}
        self.stop_flag.set()
        self.t.join(timeout=2)
        sm = sorted(r[0] for r in self.rows)
        mask = 0
        for r in self.rows:
            mask |= r[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "power_w_max": max((r[1] for r in self.rows), default=None),
                "samples": len(sm), "reasons": [n for bit, n in self.REASONS if mask & bit]}


def measured_traffic(workload):
    """DRAM bytes (dram__bytes_read + dram__bytes_write) of one movement-hour pass (k_hour + k_commit, average over the 16 movement
    hours of one simulated day) from the committed ncu capture of this build, or (None, why).  bench.py cannot read DRAM counters
    itself; the capture is tools/gpu_round.sh -> tools/day.py --json -> profiles/traffic.json."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        if t.get("workload") != workload:
            return None, "no ncu capture of this workload"
        return float(t["movement_pass_dram_bytes_per_launch"]), t.get("source", "profiles/traffic.json")
    except Exception as ex:  # noqa: BLE001
        return None, f"profiles/traffic.json unreadable: {ex}"


def algorithmic_bytes(n_agents, first_hour, n_hours):
    total = 0.0
    for h in range(first_hour, first_hour + n_hours):
        total += n_agents * (SLEEP_BYTES if 1 <= h % 24 <= 6 else ACTIVE_BYTES)
    return total


REFERENCE_SAMPLE_AGENTS = 1_000_000  # the CPU arm's region: same density and rules, population scaled so K + W whole days take minutes


def reference_config(wl, hours):
    """The workload's config for the CPU arm: the workload itself up to 1 M agents, else a 1 M-agent region of the same density
    (rho = 0.16), starting infections and intervention thresholds scaled with the population."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_ffi as O

    kw = dict(WORKLOADS[wl])
    n = kw["n_agents"]
    scale = 1.0
    if n > REFERENCE_SAMPLE_AGENTS:
        scale = REFERENCE_SAMPLE_AGENTS / n
        kw.update(WORKLOADS["1m"])
        src = WORKLOADS[wl]
        kw["exposed"] = max(1, int(round(src["exposed"] * scale)))
        if "lockdown" in src:
            kw["lockdown"] = (max(1, int(round(src["lockdown"][0] * scale))), src["lockdown"][1])
        if "hospital" in src:
            kw["hospital"] = max(1, int(round(src["hospital"] * scale)))
        if "vaccinate" in src:
            kw["vaccinate"] = src["vaccinate"]
    return O, O.make_config(hours=hours, **kw), kw["n_agents"], scale


def cpu_reference(wl, steps, warmup, threads=None):
    """The CPU restatement of the reference algorithm (oracle/, STREAM mode: hash map Point -> agent, OpenMP phase A, sequential
    phase B, process_interventions after every hour -- the structure of engine/src/allocation_map.rs:67-134).  A step = ONE WHOLE
    SIMULATED DAY, like the GPU arm; all `warmup` + `steps` days run."""
    threads = threads or os.cpu_count() or 1
    O, cfg, n, scale = reference_config(wl, 24 * (steps + warmup) + 1)
    t0 = time.time()
    eng = O.OracleEngine(cfg, seed=1, mode="stream", threads=threads)
    init_s = time.time() - t0
    per_day = []
    for d in range(warmup + steps):
        t = time.perf_counter()
        for h in range(24 * d + 1, 24 * d + 25):
            eng.step(h)
        if d >= warmup:
            per_day.append(time.perf_counter() - t)
    eng.close()
    total = sum(per_day)
    value = n * 24.0 * len(per_day) / total
    sample = (f"{n} agents" + (" (the workload's full population)" if scale == 1.0 else f" = a {scale:g} sub-sample region of the workload at the same density rho=0.16, "
              "starting infections and intervention thresholds scaled alike") + f", {len(per_day)} timed steps of one whole simulated day (24 hours, interventions applied) after {warmup} "
              f"warm-up days, oracle STREAM mode (hash map, OpenMP phase A, sequential phase B), init {init_s:.1f}s excluded")
    return value, threads, sample, total / max(1, len(per_day)), len(per_day)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 3)
    value, threads, sample, s_per_step, done = cpu_reference(args.workload, K, W)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg = {"workload": WORKLOAD_NAMES[args.workload], "agents": WORKLOADS[args.workload]["n_agents"], "grid_size": WORKLOADS[args.workload]["grid_size"],
           "step": "one simulated day (24 hours)"}
    if world > 1:  # "#4/#5 one region on CPU, scaled by region count, flagged as such"
        cfg["regions_timed"] = 1
        cfg["scaled_by_region_count"] = world
        value *= world
        sample += f"; ONE region timed on the host cores, value = that x {world} regions (the regions are independent between exchanges; the traveller exchange is not in the CPU figure)"
    line = {
        "impl": "reference", "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": done,
        "warmup": W, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def travel_plan_for(world, n_agents):
    """BASELINE config #4 pattern (engine/config/2m_100.json style, generate.py:94-97): every ordered pair exchanges
    0.1 % of a region's population as migrators (hours 48..336) and 0.05 % as commuters."""
    import numpy as np

    mig = np.full((world, world), max(1, n_agents // 1000), np.uint32)
    com = np.full((world, world), max(1, n_agents // 2000), np.uint32)
    np.fill_diagonal(mig, 0)
    np.fill_diagonal(com, 0)
    return dict(n_regions=world, migration=mig, commute=com, start_migration_hour=48, end_migration_hour=336)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from epirust_b200.engine import Engine, make_config, STATE_FIELDS
    from epirust_b200.multi import MultiRegion, max_over_ranks, share_unique_id

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: epirust_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, max(args.warmup, 3)
    kw = dict(WORKLOADS[args.workload])
    multi = world > 1
    if multi and args.workload in ("2m", "20m") and rank != 0:
        kw["exposed"] = 1  # the outbreak starts in engine1 only, the others get the default
    cfg = make_config(hours=24 * (K + W) + 49, **kw)
    n = kw["n_agents"]
    stream = torch.cuda.Stream()
    plan = travel_plan_for(world, n) if multi else None
    eng = Engine(cfg, seed=1 + rank, device=local, region=rank if multi else 0, plan=plan, extra_capacity=max(32768, 2 * (world - 1) * (n // 1000 + n // 2000)) if multi else 0)
    eng.set_stream(stream.cuda_stream)
    runner = None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def simulate(first_hour, n_hours):
        if not multi:
            got, _ = eng.simulate_hours(first_hour, n_hours)
            return got
        return runner.run(first_hour, n_hours)[0]

    sampler = ClockSampler(local)
    # ---------------- value: state resident in HBM; one event per simulated day ----------------
    with torch.cuda.stream(stream):
        if multi:
            # the ranks' NCCL communicator lives behind the C ABI (epi_comm_init); torch.distributed only carries its unique id
            runner = MultiRegion([eng], n_ranks=world, rank=rank, unique_id=share_unique_id(dist, device=torch.device("cuda", local)))
        simulate(1, 24 * W)  # warm-up days (also builds the day graph)
        barrier()
        if rank == 0:
            sampler.start()
        eng.launch_count(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        ev0.record(stream)
        rows = [simulate(24 * W + 1, 24 * K)]  # one call: the day loop (decision hours, exchanges) runs inside the library
        ev1.record(stream)
        barrier()
        t_wall1 = time.perf_counter()
        ms = ev0.elapsed_time(ev1)
        launches = eng.launch_count()
        clocks = sampler.stop() if rank == 0 else None
    got = np.concatenate(rows)
    assert len(got) == 24 * K
    last_row = [int(v) for v in got[-1]]
    ms_max, wall_ms_max = max_over_ranks(dist, [ms, (t_wall1 - t_wall0) * 1e3], device="cuda") if world > 1 else (ms, (t_wall1 - t_wall0) * 1e3)
    value = world * n * 24.0 * K / (ms_max * 1e-3)
    # which phase of the epidemic was timed: a locked-down city moves less (isolated agents skip the window load)
    locked_days = []
    events = eng.intervention_events()  # rows (hour, kind, status); kind 0 = lockdown, status 1 = locked down, 0 = revoked
    for d in range(W, W + K):
        locked = False
        for hour, kind, status in events:
            if kind == 0 and hour <= 24 * d:  # decided at the end of an earlier day's last hour
                locked = bool(status)
        locked_days.append(locked)
    phase = {"timed_days": [W + 1, W + K], "locked_down_days": int(sum(locked_days)), "infected_at_start": int(got[0][3]), "infected_at_end": int(got[-1][3])}

    # ---------------- per-kernel durations over the SAME simulated days (CUDA events around every launch, graphs off) ----
    D_t = K
    day_kernel_ms = []
    if not multi:
        eng.reset()
        eng.simulate_hours(1, 24 * W)
        eng.set_kernel_timing(True)
        for d in range(K):  # day by day: the kernel time of every day (a locked-down city moves less)
            before = sum(v[0] for v in eng.kernel_times().values())
            eng.simulate_hours(24 * (W + d) + 1, 24)
            day_kernel_ms.append(sum(v[0] for v in eng.kernel_times().values()) - before)
    else:  # a multi-region engine cannot be rewound: time the next two days of the run
        D_t = 2
        eng.set_kernel_timing(True)
        runner.run(24 * (W + K) + 1, 48)
    kt = eng.kernel_times()
    ht = eng.hour_times()
    eng.set_kernel_timing(False)
    ev_total = sum(v[0] for v in kt.values()) / D_t  # every kernel of a day, event-timed (ms)
    mov = sum(ht[h][0] + ht[h][2] for h in ht if 7 <= h <= 22) / D_t  # the 16 movement-hour passes of a day
    act = sum(ht[h][0] + ht[h][2] for h in ht) / D_t  # + h = 23 and h = 0
    # one timing source: the graph-replayed day (ms_per_step); the event-timed run only says how that day divides among the kernels
    ms_day = ms / K
    scale = ms_day / ev_total if ev_total > 0 else 1.0
    if day_kernel_ms:
        open_ms = [t * scale for t, l in zip(day_kernel_ms, locked_days) if not l]
        lock_ms = [t * scale for t, l in zip(day_kernel_ms, locked_days) if l]
        phase["ms_per_day_open"] = sum(open_ms) / len(open_ms) if open_ms else None
        phase["ms_per_day_locked_down"] = sum(lock_ms) / len(lock_ms) if lock_ms else None
        phase["note"] = "per-day figures: event-timed kernel time of each day scaled to the graph-replayed run"
    mov_pass_ms, act_pass_ms = mov * scale / 16.0, act * scale / 18.0
    peak, peak_src = peaks()
    achieved = ACTIVE_BYTES * n / (mov_pass_ms * 1e-3) / 1e9
    day_bytes = algorithmic_bytes(n, 24 * W + 1, 24)
    traffic, traffic_src = measured_traffic(args.workload) if not multi else (None, "not captured for multi-region runs")
    per_hour = {str(h): {"hour_ms": ht[h][0] / max(1, ht[h][1]), "commit_ms": ht[h][2] / max(1, ht[h][3])} for h in sorted(ht)}
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "kernel": "movement-hour pass (h%24 in 7..22) = k_hour (propose + transitions + counts + claim) + k_commit (lowest-id claim resolution); average of the 16 passes of a day",
        "algorithmic_bytes_per_launch": ACTIVE_BYTES * n, "avg_launch_ms": mov_pass_ms, "peak_source": peak_src,
        "timing_source": "ms_per_step (graph replay, CUDA events per simulated day) x the pass's share of the event-timed kernels of the same days (graphs off)",
        "event_timed_day_ms": ev_total, "graph_replay_day_ms": ms_day,
        "all_active_passes": {"passes": 18, "avg_launch_ms": act_pass_ms, "achieved": ACTIVE_BYTES * n / (act_pass_ms * 1e-3) / 1e9,
                              "frac": ACTIVE_BYTES * n / (act_pass_ms * 1e-3) / 1e9 / peak,
                              "note": "includes h=0 and h=23, whose kernels skip the grid for most agents (60 us passes credited with 82 B/agent)"},
        "per_kernel_ms": {"k_hour": kt["hour"][0] / max(1, kt["hour"][1]), "k_commit": kt["commit"][0] / max(1, kt["commit"][1]), "k_sleep": kt["sleep"][0] / max(1, kt["sleep"][1]),
                          "k_hospital_scan": kt["hospital_scan"][0] / max(1, kt["hospital_scan"][1]), "travel_kernels_per_day_ms": kt["travel"][0] / D_t},
        "per_hour_of_day_ms": per_hour,
        "whole_day": {"algorithmic_bytes": day_bytes, "achieved": day_bytes / (ms_max / K * 1e-3) / 1e9, "frac": day_bytes / (ms_max / K * 1e-3) / 1e9 / peak,
                      "note": "18 active passes x 82 B + sleep hours x 8 B per agent over the whole simulated day (max over ranks for N > 1)"},
    }

    # ---------------- e2e: host buffers -> C ABI -> host rows ----------------
    if not multi:
        # the SAME simulated days as `value`: the region is brought to the first timed day (untimed), its population goes to the
        # host (what a host-side caller would own at that point) and the timed region uploads it again and runs days W+1 .. W+K
        eng.reset()
        eng.simulate_hours(1, 24 * W)
        host_state = eng.get_state()
        pinned = {f: torch.from_numpy(host_state[f]).pin_memory() for f in STATE_FIELDS}
        host_np = {f: pinned[f].numpy() for f in STATE_FIELDS}
        h2d = sum(host_np[f].nbytes for f in STATE_FIELDS)
        barrier()
        t0 = time.perf_counter()
        eng.set_state(host_np)  # H2D of the whole population (cell, st, t0, home, work, wsa) + grid rebuild
        rows_e2e, _ = eng.simulate_hours(24 * W + 1, 24 * K)  # Counts rows D2H every simulated day
        eng.sync()
        t1 = time.perf_counter()
        e2e = {"value": n * 24.0 * K / (t1 - t0), "unit": "agent-steps/s", "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": 24 * 28,
               "rows_match_resident_run": bool(len(rows_e2e) == len(got) and (rows_e2e == got).all()),
               "note": "the population as of the first timed day uploaded from pinned host arrays once (amortised over the K days), 24 Counts rows read back per day; the same simulated days as `value`"}
    else:
        # the same K days by the host's wall clock through the public multi-region API (epi_run_multi_hours)
        e2e = {"value": world * n * 24.0 * K / (wall_ms_max * 1e-3), "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": 24 * 28 + 3 * (4 * (8 + world) + 4 * 32 * 8),
               "note": "host wall clock around the timed K days (max over ranks); the state stays in HBM (a region has no per-day host input), per day the 24 Counts rows come back and, per exchange, the exchange's scalars (TravelVars) and the running Counts totals; traveller records go GPU to GPU"}
    if runner is not None:
        runner.close()
    eng.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, threads, sample, _, _ = cpu_reference(args.workload, 3, 1)
        cpu = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample}
        if threads > 4:  # the reference's default thread count (engine-app/src/main.rs:80 `-t 4`), SURVEY.md section 8d
            v4, _, _, _, _ = cpu_reference(args.workload, 2, 1, threads=4)
            cpu["at_reference_default_threads"] = {"value": v4, "cores": 4}

    if rank == 0:
        wl = WORKLOAD_NAMES[args.workload]
        if multi:
            wl = (f"BASELINE config #4/#5 pattern: {world} regions x {n} agents (each region = {wl}), travel plan with every ordered pair "
                  f"{n // 1000} migrators/day (hours 48..336) and {n // 2000} commuters/day, one region per GPU, traveller exchange over NVLink peer memory / NCCL")
        line = {
            "metric": "agent-steps/sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl, "agents_per_gpu": n, "grid_size": kw["grid_size"], "step": "one simulated day (24 hours)",
                       "l2": "state + grids larger than L2 (no flush needed)" if n >= 5_000_000 else "working set fits the 126 MB L2; no flush (the real run is L2-resident too)",
                       "regions": world, "exchange": "epi_exchange at h%24 in {0, 7, 17}: one cooperative kernel (leave -> records pushed into the peers' memory over NVLink, CUDA IPC -> wait -> arrive); grouped ncclSend/ncclRecv when peers cannot be mapped" if multi else "n/a",
                       "phase": phase, "last_counts_row": last_row},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="10m", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
