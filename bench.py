#!/usr/bin/env python
"""bench.py -- env-steps/s (+ 7-card hand-evals/s) of the pokerl hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the unmodified reference on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # N > 1: one rank per GPU

Headline workload (BASELINE.json configs[2]): heads-up NLHE in lockstep batches of 2^20 envs, start_credits 100,
blinds 2/1, uniform random agents drawn on device from Philox(seed 0, global env id, step), auto-reset on game over.
One 2^20-env batch (92 MB of state) would fit the 126 MB L2, so each GPU holds THREE such batches ("inputs larger than
L2") and one bench "step" = one Game.step on every resident env = THREE launches of the fused sample+step kernel
(pkb::step_warp_kernel), one per batch, on one stream: every launch streams its batch from HBM and back, because the
other two batches pushed it out of L2.  Envs shard by contiguous global-id slices; there is no data-path collective
(weak scaling); one NCCL all-reduce of the payoff/statistics vector runs after the timed region.

Timing: every batch is first rolled PREROLL steps forward (untimed set-up) so that the timed steps meet the
steady-state mix of streets, hand ends and showdowns instead of 2^20 tables that were all dealt at the same moment;
then W untimed warm-up steps, then K timed steps inside a barrier + synchronize bracket, timed with one CUDA event
pair; the reported time is the max over ranks.  Consecutive step launches of a stream overlap inside the library
(programmatic dependent launch with a per-tile hand-over): `roofline.avg_launch_us` is the launch
INTERVAL of that pipeline (timed region / launches, ramp-up and drain included), and
`roofline.unpipelined_launch_us` is the duration of one launch with nothing else on the GPU.

The same JSON line carries BASELINE.json configs[3] (`nine_player`: 2^22 global envs sharded over the ranks, unequal
stacks, all-in heavy action mix) and configs[4] (`q_learning`: the batched q-learning example, 2^26 global envs with
the NCCL Q-delta reduction), the 7-card evaluator (`hand_evals`), and the CPU baselines: the unmodified reference
(baseline/_ref, one process per host core) and this repo's C port of it (oracle/).

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(num_players=2, start_credits=100, big_blind=2, small_blind=1, seed=0)
ENVS_PER_GPU = 1 << 20
RESIDENT_BATCHES = 6  # 6 x 92 MB of env state per GPU (the 126 MB L2 holds one batch and a bit); at the driver's K = 20 a timed
                      # region is 120 launches, so the pipeline's ramp-up and drain (~35 us per region) weigh 0.75 %, not 1.5 %
REPEATS = 5           # the K-step timed region is repeated this often; the median is reported
PREROLL = 64          # untimed steps every batch is rolled forward before anything is measured (steady-state hand mix)
E2E_BATCHES = int(os.environ.get("PKB_BENCH_E2E_BATCHES", "8"))  # env batches the end-to-end leg keeps in flight (6: 0.80 of `value`, 8 and 12: 0.96-0.97)
L2_FLUSH_BYTES = 256 << 20
NINE = dict(num_players=9, start_credits=[20 * (i + 1) for i in range(9)], big_blind=2, small_blind=1, seed=1,
            weights=[10, 30, 30, 5, 5, 5, 45], global_envs=1 << 22)  # BASELINE.json configs[3] / SURVEY 8(d)
QL = dict(num_players=4, global_envs=1 << 26)                        # BASELINE.json configs[4]
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def algorithmic_bytes_per_step(P: int) -> int:
    """B(P) = 2*S(P) + 6 with S(P) = 25P + 20 bytes of live state per env."""
    return 2 * (25 * P + 20) + 6


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(key: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
    except Exception:
        return None


class ClockSampler:
    """SM clock, power and throttle reasons sampled every 20 ms while the timed region runs: through NVML in a thread of
    this process (pynvml), which disturbs the GPU less than an `nvidia-smi -lms` child does (measured: the child costs an
    occasional 1 % on the 2 ms timed regions); `nvidia-smi` is the fallback where pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None
        self.thread, self.stop_flag, self.samples = None, None, []

    def _nvml_loop(self, nv, handle):
        reasons = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                   ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                   ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                   ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                mask = get_reasons(handle)
                self.samples.append((nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetPowerUsage(handle) / 1000.0,
                                     [name for name, bit in reasons if mask & bit]))
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def start(self):
        if os.environ.get("PKB_BENCH_NO_CLOCKS"):  # measurement of the sampler's own footprint only
            return
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.thread is not None:
            time.sleep(0.05)
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "source": "nvml"}
            return {"sm_mhz": statistics.median(x[0] for x in self.samples), "sm_max_mhz": max(x[1] for x in self.samples),
                    "reasons": sorted({r for x in self.samples for r in x[3]}), "samples": len(self.samples),
                    "power_w_max": max(x[2] for x in self.samples), "source": "nvml (in-process, every 20 ms)"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power), "source": "nvidia-smi -lms 20"}


# ------------------------------------------------------------------------------------------------
# CPU baselines.  (1) The UNMODIFIED reference (pure Python + numpy, installed under baseline/_ref):
# one Game per process, one process per host core, the loop of examples/random_game.py with
# auto-reset.  (2) This repo's C port of the same loop (oracle/pokerl_oracle.c, OpenMP).
# ------------------------------------------------------------------------------------------------

def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _reference_worker(job):
    """One process: the reference's Game.step loop with uniform-random valid actions for `seconds`."""
    P, seconds, seed = job
    import logging
    import random
    import signal
    sys.path.insert(0, REF_DIR)
    import numpy as np
    from pokerl import Game  # the unmodified reference (pokerl/game.py)
    logging.disable(logging.CRITICAL)
    random.seed(seed)
    np.random.seed(seed)

    class _Stuck(Exception):
        pass

    def on_alarm(*_):
        raise _Stuck()

    # the reference's `while` at game.py:607 has no termination guarantee: never let a worker outlive the sample
    signal.signal(signal.SIGALRM, on_alarm)
    signal.setitimer(signal.ITIMER_REAL, seconds + 3.0)
    game = Game(num_players=P, start_credits=100, big_blind=2, small_blind=1)
    game.reset()
    steps, in_step, stuck = 0, 0.0, 0
    clock = time.perf_counter
    t0 = clock()
    try:
        while True:
            now = clock()
            if now - t0 >= seconds:
                break
            try:
                state = game.active_state                                   # StateView construction + get_valid_actions
                action = int(np.random.choice(list(state.valid_action_indices)))  # examples/random_game.py:8
                t1 = clock()
                done, *_ = game.step(action)                                # game.py:621-700
                in_step += clock() - t1
                steps += 1
                if done:
                    game.reset()
            except AssertionError:  # game.py:473
                stuck += 1
                game.reset()
    except _Stuck:
        stuck += 1
    signal.setitimer(signal.ITIMER_REAL, 0)
    return steps, in_step, clock() - t0, stuck


def reference_rate(P: int, seconds: float, cores: int) -> dict:
    """env-steps/s of the unmodified reference over `cores` processes (full loop and Game.step-only)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context
    with ctx.Pool(cores) as pool:
        res = pool.map_async(_reference_worker, [(P, seconds, 1000 * P + i) for i in range(cores)]).get(timeout=seconds + 60)
    steps = sum(r[0] for r in res)
    wall = max(r[2] for r in res)
    return {"num_players": P, "value": steps / wall, "unit": "env-steps/s", "cores": cores,
            "per_core": steps / wall / cores, "step_only_per_core": statistics.mean(r[0] / r[1] for r in res if r[1] > 0),
            "steps": steps, "seconds": wall, "stuck_or_asserting_games": sum(r[3] for r in res)}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "pokerl"))


def cpu_baseline_reference(seconds=5.0, players=(2, 4, 9)) -> dict:
    cores = os.cpu_count() or 1
    runs = {str(P): reference_rate(P, seconds, cores) for P in players}
    head = runs[str(WORKLOAD["num_players"])]
    return {"value": head["value"], "unit": "env-steps/s", "cores": cores, "kind": "reference", "cpu_model": cpu_model(),
            "sample": f"unmodified sneppy/pokerl 0.0.6 (baseline/_ref): one Game per process, {cores} processes, "
                      f"uniform-random valid actions through Game.active_state, auto-reset on game over, "
                      f"{head['seconds']:.1f} s per table size; value = heads-up full loop",
            "by_num_players": runs}


def use_all_host_threads(ora):
    """torchrun exports OMP_NUM_THREADS=1; the CPU comparator is meant to use every host core."""
    cores = os.cpu_count() or 1
    ora.lib().ora_batch_set_threads(cores)
    return cores


def cpu_baseline_port(target_seconds=8.0) -> dict:
    """Oracle port timed on a bounded sample of the same workload."""
    from oracle import pyoracle as ora
    P = WORKLOAD["num_players"]
    cores = use_all_host_threads(ora)
    probe = ora.OracleBatch(1 << 14, P, WORKLOAD["start_credits"], WORKLOAD["big_blind"], WORKLOAD["small_blind"],
                            seed=WORKLOAD["seed"], auto_reset=True)
    probe.reset()
    probe.run_random(0, 8)
    t0 = time.perf_counter()
    n = probe.run_random(8, 32)
    rate = n / (time.perf_counter() - t0)
    envs = 1 << 16
    steps = max(16, min(4096, int(rate * target_seconds / envs)))
    b = ora.OracleBatch(envs, P, WORKLOAD["start_credits"], WORKLOAD["big_blind"], WORKLOAD["small_blind"],
                        seed=WORKLOAD["seed"], auto_reset=True)
    b.reset()
    b.run_random(0, 4)
    t0 = time.perf_counter()
    n = b.run_random(4, steps)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "env-steps/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
            "sample": f"{envs} envs x {steps} steps of the same heads-up random-agent workload "
                      f"(oracle/pokerl_oracle.c, OpenMP over {cores} host threads, {dt:.1f} s)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same metric/config."""
    if rank != 0:
        return
    seconds = float(os.environ.get("PKB_BENCH_REF_SECONDS", "6"))
    port = cpu_baseline_port()
    if reference_available():
        base = cpu_baseline_reference(seconds)
    else:  # the reference install did not travel: fall back to the port and say so
        base = dict(port, sample=port["sample"] + " [baseline/_ref missing: the C port stands in for the reference]")
    value = base["value"]
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ENVS_PER_GPU * RESIDENT_BATCHES / value * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(ENVS_PER_GPU, world, note="CPU arm: each timed sample is a bounded run of the same "
                                  "workload (tables stepped for a fixed time per process); ms_per_step = the time one "
                                  "step over all resident envs would take at the measured rate"),
        "cpu_baseline": base, "cpu_baseline_port": port,
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(envs_per_gpu, world, note=None):
    cfg = {"workload": "heads-up NLHE, random agents (BASELINE.json configs[2])", "num_players": WORKLOAD["num_players"],
           "lockstep_batch_envs": envs_per_gpu, "envs_per_gpu": envs_per_gpu * RESIDENT_BATCHES,
           "global_envs": envs_per_gpu * RESIDENT_BATCHES * world, "start_credits": WORKLOAD["start_credits"],
           "blinds": [WORKLOAD["big_blind"], WORKLOAD["small_blind"]], "agents": "uniform random over valid actions, on device",
           "auto_reset": True, "seed": WORKLOAD["seed"], "sharding": f"env-sliced x{world}, no data-path collective",
           "resident_batches_per_gpu": RESIDENT_BATCHES,
           "preroll_steps": PREROLL,
           "preroll": f"every batch is stepped {PREROLL} times before warm-up (untimed set-up): steady-state mix of streets, "
                      "hand ends and showdowns",
           "step": f"one Game.step on every resident env: {RESIDENT_BATCHES} launches (one per {envs_per_gpu}-env batch) on one stream",
           "l2": f"inputs larger than L2: {RESIDENT_BATCHES} batches of {envs_per_gpu} envs per GPU "
                 f"({RESIDENT_BATCHES * 92 * envs_per_gpu // ENVS_PER_GPU} MB of state vs 126 MB L2), every step passes over all of them"}
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------

def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from pokerl_b200 import GameBatch
    from pokerl_b200 import _native as nat
    from pokerl_b200.judger import random_hands

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nat.require_cuda()
    P, N = WORKLOAD["num_players"], args.envs
    K, W = args.steps, args.warmup
    peak, peak_src = measured_peak_gbs()

    def make_batch(j=0):
        # global env ids: every batch of every rank has its own contiguous slice
        return GameBatch(N, P, WORKLOAD["start_credits"], WORKLOAD["big_blind"], WORKLOAD["small_blind"],
                         seed=WORKLOAD["seed"], env_id_base=(rank * max(RESIDENT_BATCHES, E2E_BATCHES) + j) * N, auto_reset=True,
                         device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        """Device time (ms) of fn() between a barrier + synchronize on both sides."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        fn()
        b.record()
        barrier()
        return a.elapsed_time(b)

    batches = [make_batch(j) for j in range(RESIDENT_BATCHES)]
    for g in batches:
        g.reset()
    t_idx = 0  # step index of the device-side random agents (per batch it only has to be distinct per step)

    def run_steps(n):
        """n steps: each one Game.step on every resident env = one launch per resident batch, all on the current stream."""
        nonlocal t_idx
        for _ in range(n):
            for g in batches:
                g.step_random(t_idx)  # ONE launch: fused sampler + Game.step on the batch's N envs
            t_idx += 1

    run_steps(PREROLL)     # untimed set-up: the steady-state hand mix
    run_steps(max(W, 3))   # warm-up
    torch.cuda.synchronize()
    for g in batches:
        g.stats_reset()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    wall0 = time.perf_counter()
    # The K-step region is ~2 ms of device time at the driver's K = 20: it is timed REPEATS times, each time K steps inside
    # its own barrier + synchronize bracket, and the MEDIAN region is reported (all of them are listed in the line).
    regions = [timed(lambda: run_steps(K)) for _ in range(REPEATS)]
    dev_ms = statistics.median(regions)
    wall = time.perf_counter() - wall0
    clock_info = clocks.stop() if rank == 0 else None

    # the same K steps with plain stream order between launches: every launch alone on the GPU (the kernel's own duration)
    for g in batches:
        g.set_pipelining(False)
    unpipelined_ms = timed(lambda: run_steps(K)) / (K * RESIDENT_BATCHES)
    for g in batches:
        g.set_pipelining(True)
    stats = batches[0].stats()
    for g in batches[1:]:
        sj = g.stats()
        for k in ("steps", "hands", "showdowns", "games", "invalid", "poisoned"):
            stats[k] += sj[k]
        stats["payoff_sum"] = [a + b for a, b in zip(stats["payoff_sum"], sj["payoff_sum"])]

    # ---- for reference: ONE batch stepped back to back (its 92 MB of state stays in the 126 MB L2)
    g = batches[0]
    KF = max(16, min(K, 64))

    def one_batch():
        for i in range(KF):
            g.step_random(t_idx + i)
    resident_ms = timed(one_batch) / KF

    # ---- rollouts whose intermediate states nobody reads: FUSED steps per launch (pkb_step_random_n).  Every tile of envs
    # is loaded once, stepped FUSE times in shared memory and stored once; same six batches, same stream, K bench steps =
    # K * FUSE Game.steps per env.  State, statistics and payoffs are exactly those of single-step launches (GPU test).
    FUSE = int(os.environ.get("PKB_BENCH_FUSE", "8"))
    t_fused = t_idx + KF + 16

    def fused_steps():
        for i in range(K):
            for gb in batches:
                gb.step_random(t_fused + i * FUSE, repeat=FUSE)
    fused_steps()  # warm-up
    fused_ms = timed(fused_steps)

    # ---- end to end through the public API with HOST buffers.  Every step: H2D copy of that step's actions from
    # pinned host memory (two moves per byte), GameBatch.step_packed, D2H copy of its result byte per env (done flags +
    # valid mask) to pinned host memory.  Within a batch, copy-in -> step -> copy-out is a serial chain, so E2E_BATCHES
    # independent batches are in flight, each on its own stream; each batch's chain of R steps is captured once in a CUDA
    # graph (the copies read a different row of the pinned action recording every step) and replayed, which takes the
    # per-step Python / ctypes issue cost out of the timed region.  The actions are a recording of the device-side
    # random agents, so every one of them is valid for the state it meets.
    main = torch.cuda.current_stream(dev)
    R = int(os.environ.get("PKB_BENCH_E2E_ROUNDS", "32"))  # steps per batch; E = R * E2E_BATCHES steps in the timed region
    E = R * E2E_BATCHES
    half = (N + 1) // 2
    e2e_streams = [torch.cuda.Stream(device=dev) for _ in range(E2E_BATCHES)]
    host_actions, runs = [], []
    for j in range(E2E_BATCHES):
        rec = make_batch(j)
        rec.reset()
        ha = torch.empty((R, half), dtype=torch.uint8).pin_memory()  # half a byte per action (pkb_step_packed)
        for r in range(PREROLL):
            rec.step_random(r)
        for r in range(R):
            rec.step_random(PREROLL + r, record_actions=True)
            ha[r].copy_(GameBatch.pack_actions(rec.actions_taken))
        torch.cuda.synchronize()
        del rec
        host_actions.append(ha)
        run = make_batch(j)
        run.reset()
        for r in range(PREROLL):
            run.step_random(r)
        runs.append(run)
    dev_actions = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(E2E_BATCHES)]
    dev_out = [torch.empty(N, dtype=torch.uint8, device=dev) for _ in range(E2E_BATCHES)]
    host_out = [torch.empty((R, N), dtype=torch.uint8).pin_memory() for _ in range(E2E_BATCHES)]  # every step's result byte
    for j in range(E2E_BATCHES):  # move every handle to its own stream before the capture (a handle that changes
        with torch.cuda.stream(e2e_streams[j]):  # streams orders itself with an event, which a capture cannot contain)
            e2e_streams[j].wait_stream(main)
            runs[j].refresh_valid()
    torch.cuda.synchronize()
    graphs = []
    for j in range(E2E_BATCHES):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=e2e_streams[j]):
            for r in range(R):
                dev_actions[j].copy_(host_actions[j][r], non_blocking=True)
                runs[j].step_packed(dev_actions[j], out=dev_out[j])
                host_out[j][r].copy_(dev_out[j], non_blocking=True)
        graphs.append(gr)

    def replay_all():
        for st in e2e_streams:
            st.wait_stream(main)
        for j in range(E2E_BATCHES):
            with torch.cuda.stream(e2e_streams[j]):
                graphs[j].replay()
        for st in e2e_streams:
            st.synchronize()

    # warm-up: one untimed replay (graph upload, first touch of the pinned buffers), then the envs go back to where the
    # recorded actions apply
    saved = [run.save_state() for run in runs]
    replay_all()
    for run, st in zip(runs, saved):
        run.load_state(st)
        run.stats_reset()
    del saved
    barrier()
    t0 = time.perf_counter()
    replay_all()
    e2e_s = time.perf_counter() - t0
    # every byte that came back is a step result: R * N env-steps per batch, none rejected
    done_seen = sum(int((ho & 1).sum().item()) for ho in host_out)
    rejected = sum(run.stats()["invalid"] for run in runs)
    assert rejected == 0, f"e2e replay produced {rejected} rejected actions"
    del runs, batches, g, graphs

    # ---- BASELINE.json configs[3]: nine players, unequal stacks, all-in heavy mix, 2^22 GLOBAL envs sharded over the ranks
    from pokerl_b200 import distributed as D
    base9, n9 = D.shard_envs(NINE["global_envs"], rank, world)
    g9 = GameBatch(n9, NINE["num_players"], NINE["start_credits"], NINE["big_blind"], NINE["small_blind"],
                   seed=NINE["seed"], env_id_base=base9, auto_reset=True, device=local_rank)
    g9.reset()
    for t in range(32 + max(W, 3)):
        g9.step_random(t, NINE["weights"])
    K9 = max(4, min(K, 64))
    g9.stats_reset()

    def nine_steps():
        for i in range(K9):
            g9.step_random(1000 + i, NINE["weights"])
    nine_ms = timed(nine_steps)

    def nine_fused():
        for i in range(K9):
            g9.step_random(5000 + i * FUSE, NINE["weights"], repeat=FUSE)
    nine_fused_ms = timed(nine_fused)
    g9.set_pipelining(False)
    nine_unpipelined_ms = timed(nine_steps) / K9
    s9 = g9.stats()
    del g9

    # ---- 7-card hand evals/s on random hands resident in HBM (12 B per eval: 8 in, 4 out)
    n_hands = 1 << 30  # steady-state evals/s on >= 1 G random seven-card masks (Philox, seed 0): 8 GB in, 4 GB out
    masks = random_hands(0, n_hands, device=dev)
    out = torch.empty(n_hands, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(3):
        nat.check(nat.lib().pkb_eval_masks(nat.ptr(masks), n_hands, nat.ptr(out), stream))
    reps = 5

    def evals():
        for _ in range(reps):
            nat.check(nat.lib().pkb_eval_masks(nat.ptr(masks), n_hands, nat.ptr(out), stream))
    eval_ms = timed(evals) / reps
    del masks, out

    # ---- BASELINE.json configs[1]: all C(52,7) = 133,784,560 hands enumerated on chip (0 bytes in), histogram and the two
    # checksums of BASELINE.md s.3 compared with the values computed from the reference's judger.py
    from pokerl_b200.judger import eval_exhaustive
    EXH7 = ([0, 36768, 224848, 3473184, 4032468, 6197789, 6461665, 31433660, 58628962, 23295216, 0], 1068968016237707,
            18108495491170714949)
    eval_exhaustive(7, device=dev)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    exh = eval_exhaustive(7, device=dev)  # returns after the host has the 13 words
    exh_s = time.perf_counter() - t0
    exh_ok = (exh[0], exh[1], exh[2]) == EXH7

    # ---- BASELINE.json configs[4]: the batched q-learning example, 2^26 GLOBAL envs, NCCL Q-delta reduction per step
    ql = None
    if not args.no_q_learning:
        ql = q_learning_block(args, rank, local_rank, world, timed)

    # ---- max over ranks; NCCL reduction of the statistics vector (the only collective)
    vec = torch.tensor([dev_ms, resident_ms, e2e_s, eval_ms, unpipelined_ms, nine_ms, nine_unpipelined_ms, fused_ms, nine_fused_ms] + regions,
                       dtype=torch.float64, device=dev)
    svec = torch.tensor([stats["steps"], stats["hands"], stats["showdowns"], stats["games"], stats["invalid"],
                         stats["poisoned"]] + stats["payoff_sum"], dtype=torch.float64, device=dev)
    s9vec = torch.tensor([s9["steps"], s9["hands"], s9["showdowns"], s9["games"], s9["invalid"], s9["poisoned"]],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(svec, op=dist.ReduceOp.SUM)
        dist.all_reduce(s9vec, op=dist.ReduceOp.SUM)
    dev_ms, resident_ms, e2e_s, eval_ms, unpipelined_ms, nine_ms, nine_unpipelined_ms, fused_ms, nine_fused_ms = vec.tolist()[:9]
    regions = vec.tolist()[9:]
    dev_ms = statistics.median(regions)  # median of the per-region maxima over the ranks
    if rank != 0:
        return

    bytes_per_launch = algorithmic_bytes_per_step(P) * N
    launches = K * RESIDENT_BATCHES
    avg_launch_s = dev_ms / launches / 1e3
    achieved = bytes_per_launch / avg_launch_s / 1e9
    value = N * RESIDENT_BATCHES * world * K / (dev_ms / 1e3)
    eval_rate = n_hands * world / (eval_ms / 1e3)
    b9 = algorithmic_bytes_per_step(NINE["num_players"])
    nine_rate = NINE["global_envs"] * K9 / (nine_ms / 1e3)
    nine_gbs = b9 * (NINE["global_envs"] / world) * K9 / (nine_ms / 1e3) / 1e9  # per GPU
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(N, world),
        "roofline": {"bound": "hbm", "kernel": "pkb::step_warp_kernel<2,true>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic("step_kernel_p2"),
                     "traffic_note": ncu_traffic("step_kernel_p2_note"),
                     "algorithmic_bytes_per_env_step": algorithmic_bytes_per_step(P),
                     "bytes_per_launch": bytes_per_launch, "avg_launch_us": avg_launch_s * 1e6,
                     "avg_launch_note": "timed region / launches on ONE stream; consecutive launches overlap inside the library "
                                        "(programmatic dependent launch, tiles handed over one by one), so this is the launch "
                                        "interval of the pipeline, ramp-up and drain of the timed region included",
                     "unpipelined_launch_us": unpipelined_ms * 1e3,
                     "unpipelined_achieved": bytes_per_launch / (unpipelined_ms / 1e3) / 1e9,
                     "unpipelined_frac": bytes_per_launch / (unpipelined_ms / 1e3) / 1e9 / peak,
                     "unpipelined_note": "pkb_set_pipelining(0): every launch alone on the GPU, its tail not overlapped",
                     "peak_source": peak_src},
        "e2e": {"value": N * world * E / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": (N + 1) // 2,
                "d2h_bytes_per_step": N, "steps": E, "games_over_seen_on_host": done_seen,
                "path": "per step: half a byte per action H2D from pinned host -> GameBatch.step_packed -> pkb_step_packed -> one result "
                        "byte per env (done flags + valid mask) D2H to pinned host; "
                        f"{E2E_BATCHES} independent batches in flight on {E2E_BATCHES} streams, each batch's chain of {R} steps "
                        "captured in one CUDA graph and replayed inside the timed region; host synchronises at the end"},
        "gpu_launches": launches,  # one persistent step_warp_kernel launch per resident batch and step
        "clocks": clock_info,
        "l2_resident": {"value": N * world / (resident_ms / 1e3), "unit": "env-steps/s",
                        "frac": bytes_per_launch / (resident_ms / 1e3) / 1e9 / peak,
                        "note": "ONE batch stepped back to back on one stream: its 92 MB of state stays in the 126 MB L2"},
        "nine_player": {"metric": "env_steps_per_sec", "value": nine_rate, "unit": "env-steps/s", "steps": K9,
                        "ms_per_step": nine_ms / K9, "scaling": "strong",
                        "config": {"workload": "9-player NLHE, all-in / side-pot heavy mix (BASELINE.json configs[3])",
                                   "num_players": 9, "global_envs": NINE["global_envs"], "envs_per_gpu": NINE["global_envs"] // world,
                                   "start_credits": NINE["start_credits"], "weights_fold_check_call_r10_r25_r50_allin": NINE["weights"],
                                   "seed": NINE["seed"], "auto_reset": True, "preroll_steps": 32,
                                   "l2": "inputs larger than L2 at 1-2 GPUs (1.3 GB of state); one handle per GPU"},
                        "roofline": {"bound": "hbm", "kernel": "pkb::step_warp_kernel<9,true>", "achieved": nine_gbs, "peak": peak,
                                     "unit": "GB/s", "frac": nine_gbs / peak, "algorithmic_bytes_per_env_step": b9,
                                     "unpipelined_launch_us": nine_unpipelined_ms * 1e3, "traffic": ncu_traffic("step_kernel_p9")},
                        "fused_steps": {"steps_per_launch": FUSE, "value": NINE["global_envs"] * K9 * FUSE / (nine_fused_ms / 1e3),
                                        "unit": "env-steps/s",
                                        "frac": b9 * (NINE["global_envs"] / world) * K9 * FUSE / (nine_fused_ms / 1e3) / 1e9 / peak},
                        "stats": dict(zip(("steps", "hands", "showdowns", "games", "invalid", "poisoned"), s9vec.tolist()))},
        "hand_evals": {"metric": "hand_evals_per_sec", "value": eval_rate, "unit": "7-card evals/s", "hands": n_hands,
                       "roofline": {"bound": "hbm", "kernel": "pkb::eval_masks_fast_kernel", "achieved": 12 * eval_rate / world / 1e9,
                                    "peak": peak, "unit": "GB/s", "frac": 12 * eval_rate / world / 1e9 / peak,
                                    "traffic": ncu_traffic("eval_masks_fast_kernel"), "algorithmic_bytes_per_eval": 12}},
        "fused_steps": {"metric": "env_steps_per_sec", "value": N * RESIDENT_BATCHES * world * K * FUSE / (fused_ms / 1e3),
                        "unit": "env-steps/s", "steps_per_launch": FUSE, "launches": K * RESIDENT_BATCHES,
                        "frac": algorithmic_bytes_per_step(P) * N * FUSE / (fused_ms / 1e3 / (K * RESIDENT_BATCHES)) / 1e9 / peak,
                        "note": f"pkb_step_random_n / GameBatch.step_random(repeat={FUSE}): every launch steps each tile of envs {FUSE} "
                                "times in shared memory between one load and one store -- for rollouts with device-side agents "
                                "whose intermediate states nobody reads; the same workload, batches and stream as `value`; `frac` "
                                "counts the algorithmic bytes of every Game.step although the state crosses HBM once per launch"},
        "exhaustive_eval": {"workload": "all C(52,7) seven-card hands enumerated on chip (BASELINE.json configs[1]), rank 0",
                            "hands": 133784560, "wall_ms": exh_s * 1e3, "value": 133784560 / exh_s, "unit": "7-card evals/s",
                            "matches_reference_histogram_and_checksums": bool(exh_ok),
                            "note": "host wall time of pkb_eval_exhaustive incl. the read-back of histogram + checksums; every hand's "
                                    "value is also compared element-wise with the oracle in tests/test_gpu_eval.py"},
        "stats": {"steps": svec[0].item(), "hands": svec[1].item(), "showdowns": svec[2].item(), "games": svec[3].item(),
                  "invalid": svec[4].item(), "poisoned": svec[5].item(), "payoff_sum": svec[6:6 + P].tolist(),
                  "reduced_with": "nccl all_reduce" if world > 1 else "single rank"},
        "wall_s_timed_region": wall,
        "timed_regions_ms": regions,
        "timed_regions_note": f"the K-step region was timed {REPEATS} times (each with its own barrier + synchronize bracket, max over "
                              "ranks per region); value / ms_per_step / roofline use the median region",
    }
    if ql is not None:
        line["q_learning"] = ql
    if world == 1 and not args.no_cpu_baseline:
        port = cpu_baseline_port()
        if reference_available():
            line["cpu_baseline"] = cpu_baseline_reference(float(os.environ.get("PKB_BENCH_REF_SECONDS", "5")))
            line["cpu_baseline_port"] = port
        else:
            line["cpu_baseline"] = port
    print(json.dumps(line), flush=True)


def q_learning_block(args, rank, local_rank, world, timed):
    """BASELINE.json configs[4]: examples/q_learning.py batched to 2^26 global envs with the NCCL Q-delta reduction."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pkb_q_learning", os.path.join(ROOT, "examples", "q_learning.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    # 2^23 tables per GPU: the full 2^26 of configs[4] on eight GPUs, the same per-GPU share on fewer
    return mod.bench_block(QL["global_envs"] // 8, rank, local_rank, world, timed,
                           steps=int(os.environ.get("PKB_BENCH_QL_STEPS", "6")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (default: BASELINE configs[2], 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-q-learning", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
