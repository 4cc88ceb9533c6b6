#!/usr/bin/env python
"""bench.py -- env-steps/s (+ 7-card hand-evals/s) of the pokerl hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the CPU comparator (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # N > 1: one rank per GPU

Workload (BASELINE.json configs[2]): heads-up NLHE, 2^20 lockstep envs PER GPU, start_credits 100, blinds
2/1, uniform random agents drawn on device from Philox(seed 0, global env id, step), auto-reset on
game over.  One bench "step" = one Game.step on every env = one launch of the fused
sample+step kernel (pkb::step_warp_kernel).  Envs shard by contiguous global-id slices; there is no data-path collective
(weak scaling); one NCCL all-reduce of the payoff/statistics vector runs after the timed region.

Timing: W untimed warm-up steps, then K timed steps inside a barrier + synchronize bracket, timed with
one CUDA event pair; the reported time is the max over ranks.  One 2^20-env batch (96 MB of state)
would fit the 126 MB L2, so each GPU holds THREE such batches (289 MB, "inputs larger than L2") and
steps them round-robin: every step is one pass of the hot path over one batch of 2^20 envs whose
state was pushed out of L2 by the steps of the other two batches, i.e. it streams from HBM and back.
The batches are independent, so each one is stepped on its own CUDA stream: the drain of one step's
persistent kernel (a few tiles whose envs play several hands inside one step) overlaps the next
batch's step.  `value` is that whole-job throughput; `roofline.isolated_launch_us` is the duration of
one step launched alone on one stream (the same three batches, one stream), and `flushed_by_memset` /
`l2_resident` give the single-batch variants for reference.

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
RESIDENT_BATCHES = 3  # 3 x 96 MB of env state per GPU: larger than the 126 MB L2
E2E_BATCHES = int(os.environ.get("PKB_BENCH_E2E_BATCHES", "6"))       # env batches the end-to-end leg keeps in flight (a batch's step waits for its own copies)
L2_FLUSH_BYTES = 256 << 20


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
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 20 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
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
                "samples": len(sm), "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------------
# CPU comparator: the oracle port, all host threads (the reference itself is pure Python and does
# not exist on the GPU box; see DESIGN.md "CPU baseline")
# ------------------------------------------------------------------------------------------------

def use_all_host_threads(ora):
    """torchrun exports OMP_NUM_THREADS=1; the CPU comparator is meant to use every host core."""
    cores = os.cpu_count() or 1
    ora.lib().ora_batch_set_threads(cores)
    return cores


def cpu_rate_probe(ora, P):
    b = ora.OracleBatch(1 << 14, P, WORKLOAD["start_credits"], WORKLOAD["big_blind"], WORKLOAD["small_blind"],
                        seed=WORKLOAD["seed"], auto_reset=True)
    b.reset()
    b.run_random(0, 8)
    t0 = time.perf_counter()
    n = b.run_random(8, 32)
    return n / (time.perf_counter() - t0)


def cpu_baseline(target_seconds=12.0) -> dict:
    """Oracle port timed on a bounded sample of the same workload (rank 0, N=1 only)."""
    from oracle import pyoracle as ora
    P = WORKLOAD["num_players"]
    cores = use_all_host_threads(ora)
    rate = cpu_rate_probe(ora, P)
    envs = 1 << 16
    steps = max(16, min(4096, int(rate * target_seconds / envs)))
    b = ora.OracleBatch(envs, P, WORKLOAD["start_credits"], WORKLOAD["big_blind"], WORKLOAD["small_blind"],
                        seed=WORKLOAD["seed"], auto_reset=True)
    b.reset()
    b.run_random(0, 4)
    t0 = time.perf_counter()
    n = b.run_random(4, steps)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{envs} envs x {steps} steps of the same heads-up random-agent workload "
                      f"(oracle/pokerl_oracle.c, OpenMP over {cores} host threads, {dt:.1f} s)"}


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores, same metric/config."""
    if rank != 0:
        return
    from oracle import pyoracle as ora
    P = WORKLOAD["num_players"]
    cores = use_all_host_threads(ora)
    rate = cpu_rate_probe(ora, P)
    total_steps = args.steps + args.warmup
    # bounded sample: as many envs (<= the full 2^20) as keep the whole run near one minute
    envs = ENVS_PER_GPU
    while envs > 4096 and envs * total_steps / rate > 60.0:
        envs //= 2
    b = ora.OracleBatch(envs, P, WORKLOAD["start_credits"], WORKLOAD["big_blind"], WORKLOAD["small_blind"],
                        seed=WORKLOAD["seed"], auto_reset=True)
    b.reset()
    b.run_random(0, args.warmup)
    t0 = time.perf_counter()
    n = b.run_random(args.warmup, args.steps)
    dt = time.perf_counter() - t0
    value = n / dt
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(envs, world, note="CPU comparator on a bounded sample of the workload"),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{envs} envs x {args.steps} steps (oracle/pokerl_oracle.c port of the reference's "
                                   f"Game.step loop, OpenMP over {cores} host threads; the reference itself is pure "
                                   f"Python and is not present on the GPU box)"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(envs_per_gpu, world, note=None):
    cfg = {"workload": "heads-up NLHE, random agents (BASELINE.json configs[2])", "num_players": WORKLOAD["num_players"],
           "envs_per_gpu": envs_per_gpu, "global_envs": envs_per_gpu * world, "start_credits": WORKLOAD["start_credits"],
           "blinds": [WORKLOAD["big_blind"], WORKLOAD["small_blind"]], "agents": "uniform random over valid actions, on device",
           "auto_reset": True, "seed": WORKLOAD["seed"], "sharding": f"env-sliced x{world}, no data-path collective",
           "resident_batches_per_gpu": RESIDENT_BATCHES,
           "l2": f"inputs larger than L2: {RESIDENT_BATCHES} batches of {envs_per_gpu} envs per GPU "
                 f"({RESIDENT_BATCHES * 96 * envs_per_gpu // ENVS_PER_GPU} MB of state vs 126 MB L2) stepped round-robin, "
                 "one batch per step"}
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
    from pokerl_b200.judger import eval_masks_batch, random_hands

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nat.require_cuda()
    P, N = WORKLOAD["num_players"], args.envs
    K, W = args.steps, args.warmup

    def make_batch(j=0):
        # global env ids: every batch of every rank has its own contiguous slice
        return GameBatch(N, P, WORKLOAD["start_credits"], WORKLOAD["big_blind"], WORKLOAD["small_blind"],
                         seed=WORKLOAD["seed"], env_id_base=(rank * max(RESIDENT_BATCHES, E2E_BATCHES) + j) * N, auto_reset=True,
                         device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batches = [make_batch(j) for j in range(RESIDENT_BATCHES)]
    for g in batches:
        g.reset()
    t_idx = 0  # step index of the device-side random agents (per batch it only has to be distinct per step)

    def step_next():
        nonlocal t_idx
        batches[t_idx % RESIDENT_BATCHES].step_random(t_idx)  # ONE launch: fused sampler + Game.step on N envs
        t_idx += 1

    for _ in range(max(W, RESIDENT_BATCHES)):
        step_next()
    torch.cuda.synchronize()
    for g in batches:
        g.stats_reset()

    main = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(RESIDENT_BATCHES)]

    def run_steps(n, overlapped):
        """n steps round-robin over the resident batches; overlapped: one stream per batch, joined at the end."""
        nonlocal t_idx
        if overlapped:
            for st in streams:
                st.wait_stream(main)
        for _ in range(n):
            j = t_idx % RESIDENT_BATCHES
            if overlapped:
                with torch.cuda.stream(streams[j]):
                    batches[j].step_random(t_idx)
            else:
                batches[j].step_random(t_idx)
            t_idx += 1
        if overlapped:
            for st in streams:
                main.wait_stream(st)

    run_steps(2 * RESIDENT_BATCHES, True)
    torch.cuda.synchronize()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    wall0 = time.perf_counter()
    e_start.record()
    run_steps(K, True)
    e_stop.record()
    barrier()
    wall = time.perf_counter() - wall0
    clock_info = clocks.stop() if rank == 0 else None
    dev_ms = e_start.elapsed_time(e_stop)
    # the same K steps on ONE stream: every launch alone on the GPU (the kernel's own duration)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(K, False)
    e1.record()
    torch.cuda.synchronize()
    isolated_ms = e0.elapsed_time(e1) / K
    stats = batches[0].stats()
    for g in batches[1:]:
        sj = g.stats()
        for k in ("steps", "hands", "showdowns", "games", "invalid", "poisoned"):
            stats[k] += sj[k]
        stats["payoff_sum"] = [a + b for a, b in zip(stats["payoff_sum"], sj["payoff_sum"])]

    # ---- for reference: one batch, L2 flushed by a 256 MB memset before every step (each step timed on its own)
    g = batches[0]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    KF = min(K, 64)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(KF)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(KF)]
    barrier()
    for i in range(KF):
        flush.zero_()
        starts[i].record()
        g.step_random(t_idx + i)
        stops[i].record()
    torch.cuda.synchronize()
    flushed_ms = sum(a.elapsed_time(b) for a, b in zip(starts, stops)) / KF
    # ---- and the same batch back to back without any flush (its state stays in the 126 MB L2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(KF):
        g.step_random(t_idx + KF + i)
    e1.record()
    torch.cuda.synchronize()
    resident_ms = e0.elapsed_time(e1) / KF
    del flush

    # ---- end to end through the public API with HOST buffers.  Every step: H2D copy of that step's actions from
    # pinned host memory, GameBatch.step, D2H copy of its flags and valid masks.  As in the device-resident leg, the
    # independent resident batches take turns, each on its own stream, so one batch's copies overlap another batch's
    # step (a caller that keeps several env batches in flight does exactly this; within a batch, copy-in -> step ->
    # copy-out is a serial chain, so E2E_BATCHES of them are in flight); the actions are a recording of the
    # device-side random agents, so every one of them is valid for the state it meets.
    R = max(1, min(K, 126) // E2E_BATCHES)  # rounds; E = R * E2E_BATCHES steps
    E = R * E2E_BATCHES
    e2e_streams = [torch.cuda.Stream(device=dev) for _ in range(E2E_BATCHES)]
    host_actions, runs = [], []
    for j in range(E2E_BATCHES):
        rec = make_batch(j)
        rec.reset()
        ha = torch.empty((R, N), dtype=torch.uint8).pin_memory()  # one byte per action (pkb_step_u8)
        for r in range(R):
            rec.step_random(r, record_actions=True)
            ha[r].copy_(rec.actions_taken.to(torch.uint8))
        torch.cuda.synchronize()
        del rec
        host_actions.append(ha)
        run = make_batch(j)
        run.reset()
        runs.append(run)
    dev_actions = [torch.empty(N, dtype=torch.uint8, device=dev) for _ in range(E2E_BATCHES)]
    host_out = [torch.empty((2, N), dtype=torch.uint8).pin_memory() for _ in range(E2E_BATCHES)]  # flags, valid
    barrier()
    t0 = time.perf_counter()
    for st in e2e_streams:
        st.wait_stream(main)
    for r in range(R):
        for j in range(E2E_BATCHES):
            with torch.cuda.stream(e2e_streams[j]):
                dev_actions[j].copy_(host_actions[j][r], non_blocking=True)
                runs[j].step(dev_actions[j])
                host_out[j].copy_(runs[j].outputs, non_blocking=True)
    for st in e2e_streams:
        st.synchronize()
    e2e_s = time.perf_counter() - t0
    rejected = sum(run.stats()["invalid"] for run in runs)
    assert rejected == 0, f"e2e replay produced {rejected} rejected actions"
    del runs

    # ---- 7-card hand evals/s on random hands resident in HBM (12 B per eval: 8 in, 4 out)
    n_hands = 1 << 27
    masks = random_hands(0, n_hands, device=dev)
    out = torch.empty(n_hands, dtype=torch.int32, device=dev)
    from pokerl_b200 import _native as _n
    stream = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(3):
        _n.check(_n.lib().pkb_eval_masks(_n.ptr(masks), n_hands, _n.ptr(out), stream))
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    ev0.record()
    for _ in range(reps):
        _n.check(_n.lib().pkb_eval_masks(_n.ptr(masks), n_hands, _n.ptr(out), stream))
    ev1.record()
    torch.cuda.synchronize()
    eval_ms = ev0.elapsed_time(ev1) / reps
    del masks, out

    # ---- max over ranks; NCCL reduction of the statistics vector (the only collective)
    vec = torch.tensor([dev_ms, resident_ms, e2e_s, eval_ms, flushed_ms, isolated_ms], dtype=torch.float64, device=dev)
    svec = torch.tensor([stats["steps"], stats["hands"], stats["showdowns"], stats["games"], stats["invalid"],
                         stats["poisoned"]] + stats["payoff_sum"], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(svec, op=dist.ReduceOp.SUM)
    dev_ms, resident_ms, e2e_s, eval_ms, flushed_ms, isolated_ms = vec.tolist()
    if rank != 0:
        return

    peak, peak_src = measured_peak_gbs()
    bytes_per_launch = algorithmic_bytes_per_step(P) * N
    avg_launch_s = dev_ms / K / 1e3
    achieved = bytes_per_launch / avg_launch_s / 1e9
    value = N * world * K / (dev_ms / 1e3)
    eval_rate = n_hands * world / (eval_ms / 1e3)
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(N, world),
        "roofline": {"bound": "hbm", "kernel": "pkb::step_warp_kernel<2,true>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic("step_kernel_p2"),
                     "algorithmic_bytes_per_env_step": algorithmic_bytes_per_step(P),
                     "bytes_per_launch": bytes_per_launch, "avg_launch_us": avg_launch_s * 1e6,
                     "avg_launch_note": "timed region / launches; the launches of the three resident batches run on three streams and overlap",
                     "isolated_launch_us": isolated_ms * 1e3, "isolated_achieved": bytes_per_launch / (isolated_ms / 1e3) / 1e9,
                     "isolated_frac": bytes_per_launch / (isolated_ms / 1e3) / 1e9 / peak, "peak_source": peak_src},
        "e2e": {"value": N * world * E / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": N,
                "d2h_bytes_per_step": 2 * N, "steps": E,
                "path": "per step: one byte per action H2D from pinned host -> GameBatch.step -> pkb_step_u8 -> flags+valid D2H (one copy) to pinned host; "
                        f"{E2E_BATCHES} resident batches take turns on {E2E_BATCHES} streams, host synchronises at the end"},
        "gpu_launches": K,  # one persistent step_warp_kernel launch per step
        "clocks": clock_info,
        "l2_resident": {"value": N * world / (resident_ms / 1e3), "unit": "env-steps/s",
                        "note": "one batch stepped back to back: its 96 MB of state stays in the 126 MB L2"},
        "flushed_by_memset": {"value": N * world / (flushed_ms / 1e3), "unit": "env-steps/s", "us_per_step": flushed_ms * 1e3,
                              "note": "one batch, 256 MB memset before every step, each step timed on its own: the step "
                                      "also pays for writing the memset's dirty lines out of L2"},
        "hand_evals": {"metric": "hand_evals_per_sec", "value": eval_rate, "unit": "7-card evals/s", "hands": n_hands,
                       "roofline": {"bound": "hbm", "kernel": "pkb::eval_masks_hash_kernel", "achieved": 12 * eval_rate / world / 1e9,
                                    "peak": peak, "unit": "GB/s", "frac": 12 * eval_rate / world / 1e9 / peak,
                                    "traffic": ncu_traffic("eval_masks_hash_kernel"), "algorithmic_bytes_per_eval": 12}},
        "stats": {"steps": svec[0].item(), "hands": svec[1].item(), "showdowns": svec[2].item(), "games": svec[3].item(),
                  "invalid": svec[4].item(), "poisoned": svec[5].item(), "payoff_sum": svec[6:6 + P].tolist(),
                  "reduced_with": "nccl all_reduce" if world > 1 else "single rank"},
        "wall_s_timed_region": wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (default: BASELINE configs[2], 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
