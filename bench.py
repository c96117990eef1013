#!/usr/bin/env python
"""bench.py -- D-CBF ALIP MPC solves/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

A "step" is one pass of the hot path (dcbf_solve: problem assembly + interior-point solve) over one batch of
synthetic scenarios.  Workload at every N: BASELINE.json configs[1] -- the sig_step ALIP D-CBF MPC batched over
4096 random initial states / obstacle layouts per GPU (K = 6 circles, cold start).  With
N GPUs every rank owns its own 4096-scenario batch (weak scaling, no data-path collective); `value` is the whole-job
rate = N * 4096 * steps / max-over-ranks(device time).

  value      inputs resident in HBM, CUDA-event time of the solve kernel launches only (L2 flushed between steps)
  e2e        the same batch through the host-buffer C-ABI call dcbf_solve_host with page-locked host buffers: H2D copies of
             the inputs, kernel, D2H copies of the full result (u, plans, status, ...) every step
  roofline   FP64-pipe roofline of solve_lip_warp_kernel<1> (one problem per warp): algorithmic flop = sum_i iters_i * F_iter (SURVEY.md 8(d) formula)
             over the measured kernel time, against the FP64 DFMA peak measured on this GPU by dcbf_fp64_peak_tflops
             (MEASURED_PEAKS.json carries no FP64 figure); `hbm` sub-object: batch I/O bytes / time vs measured HBM copy
  cpu_baseline  the oracle's C port (oracle/dcbf_oracle.c) on all host threads over the same 4096 scenarios

`--impl reference` times the reference's CPU path: the reference is Python + cyipopt (not installable here: Ipopt/HSL
absent, no network), so per the task contract the arm runs the oracle port on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "configs[1]: sig_step ALIP D-CBF MPC, 4096 random states/obstacle layouts per GPU (K=6 circles, cold start)"
BATCH = 4096
SEED = 0


def f_iter(n: int, m: int, m_nl: int, m_b: int, kc: int, ke: int, form: str) -> float:
    """Algorithmic flop per interior-point iteration, SURVEY.md 8(d) (canonical formula, evaluated not re-derived)."""
    f_roll, f_step = (60, 6) if form == "dd" else (240, 30 if form == "sig_step" else 36)
    f_eval = f_roll + 130 + 3 * f_step + 3 * (14 * kc + 26 * ke)
    return (f_eval + 4 * m * n + m * n * n + m_nl * n * n + 6 * n * n + n ** 3 / 3 + 4 * n * n + 12 * (m + m_b)
            + 1.3 * (0.6 * f_eval + 2 * m))


F_ITER_SIG_K6 = f_iter(9, 30, 27, 0, 6, 0, "sig_step")   # = 8455
POOL = 8                             # distinct synthetic batches the timed steps rotate through (same pool for every N)
MAX_ITER = 200                       # dcbf_default_params (a safety cap; see DESIGN.md "iteration caps")
BUDGET = 60                          # per-problem iteration budget of the sweep's `budget` lines
CONFIG5 = dict(scenarios=1 << 20, steps=50, seed=3, n_fields=4096)    # BASELINE.json configs[4] / SURVEY.md 8(d) config 5


def workload_config():
    """the `config` object of the JSON line -- the SAME dict for both arms (bench.py and bench.py --impl reference)"""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "formulation": "sig_step", "n_circles": 6, "seed": SEED,
            "l2": "flushed between timed steps (256 MiB memset)", "max_iter": MAX_ITER,
            "batches": f"pool of {POOL} distinct batches (seeds {SEED}..{SEED + POOL - 1}); rank r solves batch (r + step) mod {POOL}"}


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture summary (profiles/, written by
    tools/ncu_lines.py / by hand from `ncu --page raw`); None if the file is missing"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return float(d["dram_bytes_per_launch"]), d.get("source", "profiles/ncu_traffic.json")
    except Exception:
        return None, None


# config 2 is a COLD start: the start vector is the reference's rule for init_guess = None, [x_k, x_k, x_k]
# (MPC_LIP_sig_step.py:185-187).  The entry points form it on the device when no start vector is passed (warm = NULL), so the
# host-buffer arms do not ship 120 bytes per scenario that are a copy of x0 (identical results: tests/test_gpu_parity.py,
# test_cold_start_rule_without_a_start_vector).
COLD_RULE_ON_DEVICE = True     # host-buffer arms (e2e): no start vector crosses PCIe
COLD_RULE_DEVICE_ARM = False   # device-resident arm (value): the start vector is resident in HBM like the other inputs (forming it in
                               # the start-order pass would only lengthen that pass)


def io_bytes_per_solve(kc: int) -> int:
    """batch I/O of one sig_step solve: x0, goal, (warm,) leg, field + the obstacle records read + all outputs."""
    return (5 + 2 + (0 if COLD_RULE_DEVICE_ARM else 15)) * 8 + 4 + 4 + 24 * kc + (15 + 15 + 9 + 1 + 1) * 8 + 4 + 4 + 1


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, all host threads), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mujoco_lip_mpc_simulation_b200 import scenarios
    from oracle import c_oracle
    cores = os.cpu_count() or 1
    sample = 1024
    # the GPU arm's workload: the same pool of POOL batches, step s solves batch s mod POOL (rank 0's rotation) -- here a bounded
    # sample of it per step, the first `sample` scenarios of that batch
    pool = [scenarios.make_batch("sig_step", BATCH, seed=SEED + j) for j in range(POOL)]
    P = c_oracle.params("sig_step", max_iter=MAX_ITER)
    sl = slice(0, sample)

    def step(s_):
        sc = pool[s_ % POOL]
        return c_oracle.solve_batch(P, sc.x0[sl], sc.goal[sl], sc.leg[sl], sc.cir, None, sc.warm[sl], field=sc.field[sl], threads=cores)
    for s_ in range(args.warmup):
        step(s_)
    t0 = time.perf_counter()
    for s_ in range(args.steps):
        step(s_)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {"impl": "reference", "metric": "D-CBF ALIP MPC solves/sec", "value": val, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"first {sample} scenarios of the step's batch (same pool and rotation as the GPU arm), oracle/dcbf_oracle.c on {cores} threads "
                                       "(reference = Python callbacks + cyipopt/Ipopt/MA57, not installable: no Ipopt, no network)"},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="also report large-batch throughput (65536 / 1M scenarios)")
    ap.add_argument("--no-config5", action="store_true", help="skip the 1M x 50 sharded closed-loop rollout (BASELINE.json configs[4])")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mujoco_lip_mpc_simulation_b200 import scenarios
    from mujoco_lip_mpc_simulation_b200.batch import DcbfSolver, SolveResult
    from mujoco_lip_mpc_simulation_b200.sharding import max_over_ranks, sum_over_ranks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)

    # ---- workload: one 4096-scenario batch per rank and step (weak scaling) ------------------------------------------
    # A step costs what its slowest problem costs, and that is data: of the batches of seeds 0..7 six take 0.46-0.47 ms, one 0.53 ms and
    # one, which holds a 53-iteration problem, 0.71 ms.  So the run works through a pool of POOL distinct batches (seeds SEED ..
    # SEED+POOL-1), the same pool for every N: rank r takes batch (r + step) mod POOL.  The value is then the throughput on the
    # scenario distribution, not on one draw of it, and no rank is stuck with the hard batch for the whole run.
    pool = [scenarios.make_batch("sig_step", BATCH, seed=SEED + j) for j in range(POOL)]
    F = pool[0].cir.shape[0]
    cir_all = np.concatenate([b_.cir for b_ in pool], axis=0)
    sc = pool[rank % POOL]                              # this rank's batch for the single-solve latency probe
    sc_last = pool[(rank + args.steps - 1) % POOL]      # the batch of the last timed step (status / agreement checks)
    solver = DcbfSolver("sig_step", device=local)
    assert int(solver.P.max_iter) == MAX_ITER
    solver.set_fields(cir_all)
    t = lambda a, dt: torch.as_tensor(a, dtype=dt, device=dev)  # noqa: E731
    dev_in = [(t(b_.x0, torch.float64), t(b_.goal, torch.float64), t(b_.leg, torch.int32), t(b_.field + j * F, torch.int32), t(b_.warm, torch.float64))
              for j, b_ in enumerate(pool)]
    B = BATCH
    out = SolveResult(torch.empty((B, 15), dtype=torch.float64, device=dev), torch.empty((B, 3, 5), dtype=torch.float64, device=dev),
                      torch.empty((B, 3, 3), dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.int32, device=dev),
                      torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.float64, device=dev),
                      torch.empty(B, dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.uint8, device=dev))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    it_acc = torch.zeros((), dtype=torch.int64, device=dev)

    def device_step(ev0, ev1, s_):
        x0, goal, leg, field, warm = dev_in[(rank + s_) % POOL]
        flush.zero_()                      # L2 flush between timed iterations (inputs are far smaller than L2)
        ev0.record()
        solver.solve_into(B, x0, goal, leg, field, None if COLD_RULE_DEVICE_ARM else warm, None, out)
        ev1.record()
        it_acc.add_(out.iters.sum())       # outside the event bracket: iterations of this step, for the roofline

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for s_ in range(args.warmup):
        device_step(w0, w1, s_)
    torch.cuda.synchronize()
    it_acc.zero_()
    fp64_peak = solver.fp64_peak_tflops(3)

    sampler = ClockSampler(local)
    sampler.start()
    # ---- timed region 1: device-resident --------------------------------------------------------------------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = solver.launches
    for s_, (e0, e1) in enumerate(evs):
        device_step(e0, e1, s_)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = solver.launches - l0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    dev_ms = max_over_ranks(float(sum(step_ms)), dev)
    iters = out.iters.cpu().numpy().astype(np.int64)
    status = out.status.cpu().numpy()
    p_plan_last = out.p_plan.cpu().numpy()      # results of the last timed step (the buffers are reused below)
    flop_per_step = float(it_acc.item()) / args.steps * F_ITER_SIG_K6     # mean over the timed steps (the batches rotate when N > 1)

    # ---- pipelined throughput: consecutive batches on rotating streams (one context each) -------------------------------------------
    # A step of `value` ends when its slowest chain of scenarios ends: 4096 scenarios on 2368 persistent warps are 1.73 per warp, so
    # three quarters of the warps run two problems while the rest run one and then idle, and a 29-iteration problem alone is most of
    # the step (DESIGN.md "the tail").  A caller with a stream of batches does
    # not have to wait: batch k+1 is launched on a second stream and its warps take over the SM slots batch k's warps vacate.
    n_lanes = int(os.environ.get("DCBF_BENCH_LANES", "3"))
    lanes = [(solver, out, torch.cuda.Stream(device=dev))]
    for _ in range(n_lanes - 1):
        sv_ = DcbfSolver("sig_step", device=local)
        sv_.set_fields(cir_all)
        o_ = SolveResult(*[torch.empty_like(t_) for t_ in (out.u, out.x_plan, out.p_plan, out.status, out.iters, out.obj, out.viol, out.close2goal)])
        lanes.append((sv_, o_, torch.cuda.Stream(device=dev)))

    def pipelined(n_steps):
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        p0.record(cur)
        for sv_, _, st_ in lanes:
            st_.wait_event(p0)
        for s_ in range(n_steps):
            sv_, o_, st_ = lanes[s_ % n_lanes]
            x0, goal, leg, field, warm = dev_in[(rank + s_) % POOL]
            with torch.cuda.stream(st_):
                sv_.solve_into(B, x0, goal, leg, field, None if COLD_RULE_DEVICE_ARM else warm, None, o_)
        for _, _, st_ in lanes:
            cur.wait_stream(st_)
        p1.record(cur)
        torch.cuda.synchronize()
        return p0.elapsed_time(p1)
    pipelined(4)
    if world > 1:
        dist.barrier()
    pipe_steps = max(args.steps, 40)
    pipe_ms = max_over_ranks(pipelined(pipe_steps), dev)
    if world > 1:
        dist.barrier()
    del lanes[1:]

    # ---- timed region 2: end to end through the host-buffer C-ABI call ----------------------------------------------
    # inputs and results live in page-locked host memory (the copies inside the timed call are DMA transfers from / to them)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    host_in = [(pin(b_.x0), pin(b_.goal), pin(b_.leg.astype(np.int32)), pin(b_.warm), pin((b_.field + j * F).astype(np.int32))) for j, b_ in enumerate(pool)]
    solver.set_fields_host(cir_all)
    hres = SolveResult(pin(np.empty((B, 15))), pin(np.empty((B, 3, 5))), pin(np.empty((B, 3, 3))), pin(np.empty(B, np.int32)),
                       pin(np.empty(B, np.int32)), pin(np.empty(B)), pin(np.empty(B)), pin(np.empty(B, np.uint8)))

    def host_step(s_):
        hx0, hgoal, hleg, hwarm, hfield = host_in[(rank + s_) % POOL]
        solver.solve_host(hx0, hgoal, hleg, None if COLD_RULE_ON_DEVICE else hwarm, field=hfield, out=hres)
    for s_ in range(3):
        host_step(s_)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s_ in range(args.steps):
        host_step(s_)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert np.array_equal(hres.status, status), "host-buffer path and device path disagree"
    # ---- the same end-to-end path for a caller with a stream of batches: dcbf_solve_host_async on three contexts used round-robin
    # (HostPipeline); every step still reads its inputs from and delivers its full result to page-locked host memory ----------------
    from mujoco_lip_mpc_simulation_b200.batch import HostPipeline
    hp_ = HostPipeline("sig_step", lanes=3, device=local)
    hp_.set_fields_host(cir_all)
    hp_outs = [SolveResult(pin(np.empty((B, 15))), pin(np.empty((B, 3, 5))), pin(np.empty((B, 3, 3))), pin(np.empty(B, np.int32)),
                           pin(np.empty(B, np.int32)), pin(np.empty(B)), pin(np.empty(B)), pin(np.empty(B, np.uint8))) for _ in range(3)]

    def host_pipelined(n_steps):
        t0_ = time.perf_counter()
        for s_ in range(n_steps):
            hx0, hgoal, hleg, hwarm, hfield = host_in[(rank + s_) % POOL]
            hp_.submit(hx0, hgoal, hleg, None if COLD_RULE_ON_DEVICE else hwarm, field=hfield, out=hp_outs[s_ % 3])    # waits for the lane's previous batch first
        hp_.drain()
        return time.perf_counter() - t0_
    host_pipelined(6)
    if world > 1:
        dist.barrier()
    e2e_pipe_steps = max(args.steps, 40)
    e2e_pipe_s = max_over_ranks(host_pipelined(e2e_pipe_steps), dev)
    last_b = (rank + e2e_pipe_steps - 1) % POOL        # the pipelined path delivers what the synchronous call delivers
    hx0, hgoal, hleg, hwarm, hfield = host_in[last_b]
    assert np.array_equal(hp_outs[(e2e_pipe_steps - 1) % 3].status, solver.solve_host(hx0, hgoal, hleg, hwarm, field=hfield).status)
    del hp_
    h2d = B * ((5 + 2 + (0 if COLD_RULE_ON_DEVICE else 15)) * 8 + 4 + 4)   # x0, goal, (start vector,) leg, field
    d2h = B * ((15 + 15 + 9 + 1 + 1) * 8 + 4 + 4 + 1)

    # ---- p50 single-solve latency (B = 1, launch to result, host buffers) ------------------------------------------------
    lat = []
    own_field = (sc.field + (rank % POOL) * F).astype(np.int32)
    one = solver.solve_host(sc.x0[:1], sc.goal[:1], sc.leg[:1], sc.warm[:1], field=own_field[:1])
    for i in range(200):
        j = i % B
        t0 = time.perf_counter()
        solver.solve_host(sc.x0[j:j + 1], sc.goal[j:j + 1], sc.leg[j:j + 1], sc.warm[j:j + 1], field=own_field[j:j + 1], out=one)
        lat.append((time.perf_counter() - t0) * 1e6)

    # ---- config 5 (BASELINE.json configs[4]): closed-loop LIP rollout, 1 M scenarios x 50 steps with warm-started re-planning,
    # sharded by batch slice over the ranks (strong scaling: 1 M / N scenarios per GPU), final result gather over NCCL -------------
    cfg5 = None
    if not args.no_config5:
        from mujoco_lip_mpc_simulation_b200.sharding import gather_rollout, rollout_shard, shard_inputs
        B5, S5 = CONFIG5["scenarios"], CONFIG5["steps"]
        sv5 = DcbfSolver("sig_step", device=local)
        warm5 = shard_inputs(sv5, 8192, seed=99, n_fields=256, rank=rank, world=world)      # warm-up: kernel load, allocator, NCCL
        gather_rollout(rollout_shard(sv5, 2, warm5), 8192)
        inp5 = shard_inputs(sv5, B5, seed=CONFIG5["seed"], n_fields=CONFIG5["n_fields"], rank=rank, world=world)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        l5 = sv5.launches
        ev[0].record()
        r5 = rollout_shard(sv5, S5, inp5)
        ev[1].record()
        g5 = gather_rollout(r5, B5)                # the only collective of the path (one NCCL all-gather of 56 B per scenario)
        ev[2].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        roll_ms = max_over_ranks(ev[0].elapsed_time(ev[1]), dev)
        gath_ms = max_over_ranks(ev[1].elapsed_time(ev[2]), dev)
        solves5 = sum_over_ranks(float(r5["steps_done"].sum().item()), dev)
        iters5 = sum_over_ranks(float(r5["total_iters"].sum().item()), dev)
        ninf5 = sum_over_ranks(float(r5["n_infeasible"].sum().item()), dev)
        if rank == 0:
            xf, sd = g5["x_final"], g5["steps_done"]
            arrived = sd < S5          # stopped early: close_2_goal of MPC_LIP_sig_step.py:110-111 (a planned step within 0.35 m of the goal)
            cfg5 = {"workload": "configs[4]: closed-loop LIP rollout, 1 M scenarios x 50 steps, warm-started re-planning, sharded by batch slice",
                    "scenarios": B5, "steps": S5, "seed": CONFIG5["seed"], "field_pool": CONFIG5["n_fields"], "scaling": "strong",
                    "scenarios_per_gpu": (B5 + world - 1) // world, "solves": int(solves5), "rollout_ms": roll_ms,
                    "solves_per_s": solves5 / (roll_ms * 1e-3), "gather_ms": gath_ms, "gather_bytes": B5 * 56,
                    "gather": "one all_gather_into_tensor of [n, 7] FP64 per rank (NCCL)" if world > 1 else "none (one rank)",
                    "mean_iters_per_replan": iters5 / solves5, "infeasible_replan_frac": ninf5 / solves5,
                    "mean_steps": float(sd.double().mean()), "arrived_frac": float(arrived.double().mean()),
                    "gpu_launches": int(sv5.launches - l5),
                    # rank-count-independent fingerprint of the gathered result (identical at N = 1, 2, 4, 8)
                    "checksum": {"steps_done_sum": int(sd.sum()), "n_infeasible_sum": int(g5["n_infeasible"].sum()),
                                 "x_final_nansum": float(torch.nan_to_num(xf).sum())}}
        del r5, g5, inp5, sv5

    # ---- the same kernel at throughput: one 65 536-scenario batch (16 x the step's batch; no tail to speak of) -- the roofline of the
    # kernel itself next to the roofline of the 4096-scenario step -------------------------------------------------------------------
    sat = None
    if rank == 0:
        Bs = 65536
        s2 = scenarios.make_batch("sig_step", Bs, seed=SEED + 1)
        sv = DcbfSolver("sig_step", device=local)
        sv.set_fields(s2.cir)
        a = [t(s2.x0, torch.float64), t(s2.goal, torch.float64), t(s2.leg, torch.int32), t(s2.warm, torch.float64), t(s2.field, torch.int32)]
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            e0.record()
            r = sv.solve(a[0], a[1], a[2], a[3], field=a[4])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[2:]))
        tf = float(r.iters.sum().item()) * F_ITER_SIG_K6 / (ms * 1e-3) * 1e-12
        sat = {"batch": Bs, "ms": ms, "solves_per_s": Bs / (ms * 1e-3), "achieved": tf, "frac": tf / fp64_peak if fp64_peak > 0 else None,
               "mean_iters": float(r.iters.float().mean()), "launch_ms": [round(x, 4) for x in ts],
               "how": "median of 5 launches after two warm-ups, L2 flushed before each, CUDA events"}
        del sv, r, a

    extra = {}
    if args.sweep and rank == 0:
        for form, Bs in (("sig_step", 65536), ("sig_step", 1 << 20), ("modi", 65536), ("dd", 4096), ("dd", 65536)):
            s2 = scenarios.make_batch(form, Bs, seed=SEED + 1)
            sv = DcbfSolver(form, device=local)
            sv.set_fields(s2.cir, s2.elp if s2.elp.shape[1] else None)
            a = [t(s2.x0, torch.float64), t(s2.goal, torch.float64), t(s2.leg, torch.int32), t(s2.warm, torch.float64)]
            lu = None if s2.last_u is None else t(s2.last_u, torch.float64)
            fld = t(s2.field, torch.int32)
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.zero_()
                e0.record()
                r = sv.solve(a[0], a[1], a[2], a[3], field=fld, last_u=lu)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            extra[f"{form}_{Bs}"] = {"solves_per_s": Bs / (best * 1e-3), "ms": best, "mean_iters": float(r.iters.float().mean()),
                                     "max_iters": int(r.iters.max())}
            if Bs == 65536:
                # the same batch with a per-problem iteration budget (dcbf_params::max_iter = 60; the reference caps its own solver at
                # 20 / 30 / 40 iterations and files what it has): a problem that crawls out of a saddle point for 100+ iterations
                # is a serial millisecond at the end of the batch; it comes back with status -1 (Maximum_Iterations_Exceeded)
                svb = DcbfSolver(form, device=local, max_iter=BUDGET)
                svb.set_fields(s2.cir, s2.elp if s2.elp.shape[1] else None)
                bestb = 1e9
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    flush.zero_()
                    e0.record()
                    rb = svb.solve(a[0], a[1], a[2], a[3], field=fld, last_u=lu)
                    e1.record()
                    torch.cuda.synchronize()
                    bestb = min(bestb, e0.elapsed_time(e1))
                extra[f"{form}_{Bs}"]["budget"] = {"max_iter": BUDGET, "solves_per_s": Bs / (bestb * 1e-3), "ms": bestb,
                                                   "hit_budget": int((rb.status == -1).sum()),
                                                   "other_results_changed": int(((rb.status != r.status) & (rb.status != -1)).sum())}
                del svb, rb
        # control tick (dcbf_tick) on the modi shape: prediction + warm-start rule + re-plan + dense plan trajectory [B,126,2]
        s3 = scenarios.make_batch("modi", 65536, seed=SEED + 1)
        sv = DcbfSolver("modi", device=local)
        sv.set_fields(s3.cir, s3.elp)
        targs = [t(s3.x0[:, 0:2], torch.float64), t(s3.x0[:, 2:4], torch.float64), t(s3.x0[:, 4], torch.float64),
                 t(np.concatenate([s3.x0[:, 0:2], np.zeros((65536, 1))], axis=1), torch.float64), t(np.full(65536, 0.1), torch.float64)]
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            e0.record()
            tk = sv.tick(targs[0], targs[1], targs[2], targs[3], targs[4], t(s3.goal, torch.float64), t(s3.leg, torch.int32), field=t(s3.field, torch.int32))
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        extra["tick_modi_65536"] = {"ticks_per_s": 65536 / (best * 1e-3), "ms": best, "pos_det_bytes": int(tk["pos_det"].numel() * 8),
                                    "mean_iters": float(tk["plan"].iters.float().mean())}
        # the tick-rate workload of the reference (main_sim_mpc.py:85-88: 40 re-plans per step): the next tick, 10 ms later, from a state
        # 5 mm / 2 cm/s away, warm-started from the plan above verbatim (mode 0 -> first barrier parameter mu_warm)
        gen = torch.Generator(device=dev); gen.manual_seed(5)
        pos2 = targs[0] + 0.005 * torch.randn(targs[0].shape, generator=gen, device=dev, dtype=torch.float64)
        vel2 = targs[1] + 0.02 * torch.randn(targs[1].shape, generator=gen, device=dev, dtype=torch.float64)
        prev, md0 = tk["plan"].x_plan.reshape(65536, 15).clone(), torch.zeros(65536, dtype=torch.uint8, device=dev)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            e0.record()
            tw = sv.tick(pos2, vel2, targs[2], targs[3], targs[4], t(s3.goal, torch.float64), t(s3.leg, torch.int32), prev_plan=prev, mode=md0,
                         field=t(s3.field, torch.int32))
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        extra["tick_warm_modi_65536"] = {"ticks_per_s": 65536 / (best * 1e-3), "ms": best, "mean_iters": float(tw["plan"].iters.float().mean()),
                                         "mu_warm": float(sv.P.mu_warm)}
        del tk, tw
        # config 5 shape, one GPU's share at 8 GPUs: closed-loop rollout, 131072 scenarios x 50 steps, no host round trips
        s5 = scenarios.make_batch("sig_step", 131072, seed=SEED + 3)
        sv = DcbfSolver("sig_step", device=local)
        sv.set_fields(s5.cir)
        a5 = [t(s5.x0, torch.float64), t(s5.goal, torch.float64), t(s5.leg, torch.int32), t(s5.field, torch.int32)]
        best = 1e9
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ro = sv.rollout(50, a5[0], a5[1], a5[2], field=a5[3], want_traj=False)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        n_solves = int(ro["steps_done"].sum())
        extra["rollout_131072x50"] = {"solves_per_s": n_solves / (best * 1e-3), "ms": best, "solves": n_solves,
                                      "mean_iters": float(ro["total_iters"].sum()) / n_solves,
                                      "infeasible_frac": float(ro["n_infeasible"].sum()) / n_solves}

    # ---- CPU baseline (rank 0, N = 1 only): oracle port on all host threads over the same 4096 scenarios ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle
        cores = os.cpu_count() or 1
        P = c_oracle.params("sig_step", max_iter=MAX_ITER)
        t0 = time.perf_counter()
        ref = c_oracle.solve_batch(P, sc_last.x0, sc_last.goal, sc_last.leg, sc_last.cir, None, sc_last.warm, field=sc_last.field, threads=cores)
        dt = time.perf_counter() - t0
        agree_cls = float(np.mean((status == 2) == (ref["status"] == 2)))
        both = (status == 0) & (ref["status"] == 0)
        dp = np.abs(p_plan_last - ref["p_plan"]).reshape(B, -1).max(axis=1)
        cpu = {"value": B / dt, "unit": "solves/s", "cores": cores, "kind": "port",
               "sample": f"the full {B}-scenario workload once, oracle/dcbf_oracle.c (C restatement, FD Hessian) on {cores} threads",
               "agreement": {"status_class": agree_cls, "solution_1e-4": float(np.mean(dp[both] <= 1e-4)), "both_converged": int(both.sum())}}

    total_solves = sum_over_ranks(float(B * args.steps), dev)
    if rank == 0:
        dev_s = dev_ms * 1e-3
        kernel_s = float(np.mean(step_ms)) * 1e-3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        achieved_tf = flop_per_step / kernel_s * 1e-12
        io_gbs = B * io_bytes_per_solve(6) / kernel_s * 1e-9
        line = {
            "metric": "D-CBF ALIP MPC solves/sec", "value": total_solves / dev_s, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "e2e": {"value": total_solves / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "transfer": ("page-locked host buffers through dcbf_solve_host; the kernels load the inputs from and store the results to the "
                                 "mapped host memory (no staging copy)" if os.environ.get("DCBF_ZEROCOPY", "1") != "0" else
                                 "page-locked host buffers through dcbf_solve_host, cudaMemcpyAsync each way"),
                    "inputs": ("x0, goal, leg, field per scenario; the cold-start vector [x_k, x_k, x_k] (reference: init_guess = None, "
                               "MPC_LIP_sig_step.py:185-187) is formed on the device (warm = NULL)" if COLD_RULE_ON_DEVICE else
                               "x0, goal, leg, field and the start vector per scenario"),
                    "pipelined": {"value": world * B * e2e_pipe_steps / e2e_pipe_s, "unit": "solves/s", "steps": e2e_pipe_steps, "lanes": 3,
                                  "how": "dcbf_solve_host_async on three contexts used round-robin (HostPipeline), wall clock over all steps incl. the "
                                         "final waits; same page-locked inputs and full results per step as `value` of this object"}},
            "gpu_launches": int(launches),
            "throughput_pipelined": {"value": world * B * pipe_steps / (pipe_ms * 1e-3), "unit": "solves/s", "steps": pipe_steps,
                                     "ms_per_step": pipe_ms / pipe_steps,
                                     "streams": n_lanes,
                                     "how": "the same batches launched back to back on rotating streams (one context each), device time over all steps; "
                                            "no L2 flush in between (the batches overlap), inputs rotate through the pool"},
            "p50_solve_us": float(np.median(lat)), "p95_solve_us": float(np.percentile(lat, 95)),
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak > 0 else None,
                         "traffic": ncu_traffic()[0], "traffic_source": ncu_traffic()[1],
                         "kernel": "solve_lip_warp_kernel<LipW, 1> (one problem per warp, 16 warps per SM)", "algorithmic_bytes": B * io_bytes_per_solve(6),
                         "peak_source": "measured on this GPU by dcbf_fp64_peak_tflops (DFMA loop); MEASURED_PEAKS.json has no FP64 figure",
                         "flop_per_iter": F_ITER_SIG_K6, "iters_per_step": flop_per_step / F_ITER_SIG_K6,
                         "saturated": sat,
                         "hbm": {"achieved": io_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": io_gbs / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "iters": {"mean": float(iters.mean()), "p50": float(np.median(iters)), "p99": float(np.percentile(iters, 99)), "max": int(iters.max()),
                      "over_40": int((iters > 40).sum()), "of": "the batch of the last timed step"},
            "status_hist": {str(int(k)): int((status == k).sum()) for k in np.unique(status)},
            "clocks": sampler.summary(),
        }
        if cpu:
            try:   # the reference's own Python path, measured where the reference exists (tools/reference_python_baseline.py)
                cpu["reference_python"] = json.load(open(os.path.join(ROOT, "profiles", "r03_reference_python.json")))
            except Exception:
                cpu["reference_python"] = None
            line["cpu_baseline"] = cpu
        if cfg5:
            line["config5"] = cfg5
        if extra:
            line["sweep"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
