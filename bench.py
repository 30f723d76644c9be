#!/usr/bin/env python
"""bench.py -- training env-steps/sec, HumanoidAMPGetup ASE pre-train, 4096 envs/GPU (BASELINE.json metric).

One "step" = one training epoch of the hot path on one batch of synthetic input:
  32-step rollout of 4096 envs (Isaac Gym bypassed with synthetic rigid-body-state tensors; observation build,
  AMP history, actor / critic inference, disc + enc rewards, GAE all run) + 6 x 8 minibatch updates
  (calc_gradients + Adam, B = 16384, B_amp = 4096)  = 131072 env-steps per GPU.

  value : inputs (rigid-body states) already resident in HBM when the timed region starts
  e2e   : same epoch through the public agent API with the rigid-body states arriving from PINNED HOST memory
          every sim step (H2D inside the timed region) and the train_result scalars read back to the host (D2H)

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
  python bench.py --impl reference ...     # the reference's own algorithm on the host CPU cores (oracle port)
  python bench.py --config 2|3|5           # BASELINE.json configs: 2 AMP-only (HumanoidAMP), 3 ASE pre-train (default, the metric's config),
                                           # 5 HRL heading task over a frozen ASE low-level controller

Measurement rules: `value` is timed with NO per-launch instrumentation (CUDA events around the K epochs only); the per-kernel GEMM time
behind `roofline` comes from ONE extra, untimed epoch run with ase_gemm_tc_profile on.  The CPU arm (`cpu_baseline`, --impl reference) runs
the oracle port on a FIXED thread count (min(32, cores in the affinity mask)) and times a step that is an exact 1/16 of an epoch in the
epoch's own proportions (2 of 32 rollout steps, 8192 of 131072 reward rows, 3 of 48 minibatch updates): ms_per_step is real wall time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "training env-steps/sec HumanoidAMPGetup 4096 envs/GPU"
UNIT = "env-steps/s"
NUM_ENVS, HORIZON, MINIBATCH, AMP_MINIBATCH, MINI_EPOCHS = 4096, 32, 16384, 4096, 6
GEMM_BACKEND = int(os.environ.get("ASE_GEMM_BACKEND", "2"))      # 2: tcgen05 3xFP16 scaled planes (default), 1: tcgen05 3xTF32
FLOP_PER_MINIBATCH = 0.9685e12      # algorithmic, fp32, fwd + bwd + gradient penalty
FLOP_ROLLOUT_PER_EPOCH = 3.13e12


CONFIG = 3
WORKLOADS = {2: ("amp", "HumanoidAMP single-clip imitation (AMP-only: no encoder / latents / diversity, MLPs [1024, 512]), synthetic rigid-body states, Isaac Gym bypassed"),
             3: ("ase", "HumanoidAMPGetup ASE pre-train (disc+encoder+diversity), synthetic rigid-body states, Isaac Gym bypassed"),
             5: ("hrl", "HumanoidHeading HRL task-train (tanh-mu HLC [1024, 512], frozen full-size ASE LLC stepped 5x per action), synthetic rigid-body states, Isaac Gym bypassed")}
CPU_FRACTION = 16                    # the CPU arm's step = 1 / 16 of an epoch


def _workload_config(n_gpus):
    return {"workload": WORKLOADS[CONFIG][1], "baseline_config": CONFIG,
            "num_envs_per_gpu": NUM_ENVS, "horizon": HORIZON, "minibatch": MINIBATCH, "amp_minibatch": AMP_MINIBATCH,
            "mini_epochs": MINI_EPOCHS, "minibatches_per_step": MINI_EPOCHS * (NUM_ENVS * HORIZON // MINIBATCH),
            "env_steps_per_step_per_gpu": NUM_ENVS * HORIZON, "parallelism": f"env-sharded dp{n_gpus}",
            "l2_policy": "inputs larger than L2 (per-epoch working set ~3 GB: experience buffers, activations, 28 MB weights x4)"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True); self._th.start(); return self

    def __exit__(self, *a):
        self._stop.set(); self._th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ ours
def _make_agent(torch, rank, world, state_source, seed):
    from ase_b200 import configs
    from ase_b200.agent import ASEAgent, AMPAgent, HRLAgent
    from ase_b200.synthetic_env import SyntheticHumanoidEnv
    dev = f"cuda:{torch.cuda.current_device()}"
    kind = WORKLOADS[CONFIG][0]
    env = SyntheticHumanoidEnv(NUM_ENVS, device=dev, seed=seed + rank, state_source=state_source,    # seed += rank (run.py:36-50)
                               local_root_obs=(kind != 'amp'), heading_task=(kind == 'hrl'))
    cfg = configs.make(kind, device=dev, vec_env=env, num_actors=NUM_ENVS, multi_gpu=world > 1, print_stats=False,
                       seed=seed, gemm_backend=GEMM_BACKEND)
    agent = {'ase': ASEAgent, 'amp': AMPAgent, 'hrl': HRLAgent}[kind]('bench', cfg)
    agent.init_tensors()
    agent.obs = agent.env_reset()
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(agent.model.params, 0)
        agent.model.params_changed()
    agent._init_train()
    return agent, env


def _plane_flags(agent):
    """gemm_backend 2: sticky scale-miss flags of the learner (0 = fine); a miss is reported in the JSON line and on stderr."""
    import ctypes as C
    from ase_b200 import lib as L
    f = C.c_int(0)
    import torch
    L.check(L.lib.ase_learner_plane_status(agent.model._h, C.byref(f), torch.cuda.current_stream().cuda_stream), 'ase_learner_plane_status')
    if f.value:
        sys.stderr.write(f"WARNING: FP16 operand-plane scale miss (flags {f.value}) -- the flagged updates are not fp32-accurate\n")
    return int(f.value)


def _timed_epochs(torch, agent, steps, world, d2h=False):
    """barrier + synchronize on both sides, CUDA events on the launching (current) stream, max over ranks."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = []
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    e0.record()
    for k in range(steps):
        agent.update_epoch()
        info = agent.train_epoch()
        if world > 1:
            agent._sync_stats()                        # the per-epoch RunningMeanStd averaging of agent.train() (hvd.sync_stats)
        if d2h:
            host.append(agent._tr_buf.cpu())           # the step's train_result series -> host (D2H inside the timed region)
        marks[k].record()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    global _EPOCH_MS
    _EPOCH_MS = [round((e0 if k == 0 else marks[k - 1]).elapsed_time(marks[k]), 2) for k in range(steps)]      # this rank's per-epoch times
    return float(ms) / 1e3, host


_EPOCH_MS = []


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL logs (version line, NCCL_DEBUG=INFO) must not mix into the one JSON line on stdout
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from ase_b200 import lib as L
    import ctypes as C
    torch.manual_seed(1234 + rank)

    agent, env = _make_agent(torch, rank, world, 'device', seed=0)
    for _ in range(args.warmup):
        agent.update_epoch(); agent.train_epoch()
        if world > 1:
            agent._sync_stats()        # part of every epoch of agent.train(); its first call also sets up NCCL's connections (~0.2 s): warm-up, not timed
    torch.cuda.synchronize()
    launches0 = L.launch_count()
    with ClockSampler(local) as clk:
        secs, _ = _timed_epochs(torch, agent, args.steps, world)
    value_epoch_ms = list(_EPOCH_MS)
    peer_adam = bool(getattr(agent, 'peer_adam', False))
    launches = L.launch_count() - launches0
    play_t, upd_t, _ = agent.epoch_times()
    # roofline numerator / denominator: ONE extra epoch, outside the timed region, with CUDA events around every tcgen05 GEMM launch
    agent.set_graphs(False)          # per-launch events need eager launches: this one epoch runs without the rollout / minibatch CUDA graphs
    L.lib.ase_gemm_tc_profile(1)
    agent.update_epoch(); agent.train_epoch()
    torch.cuda.synchronize()
    agent.set_graphs(True)
    tot_ms, nl, fl = C.c_double(), C.c_int64(), C.c_double()
    L.check(L.lib.ase_gemm_tc_profile_read(C.byref(tot_ms), C.byref(nl), C.byref(fl)), 'profile_read')
    L.lib.ase_gemm_tc_profile(0)
    prof_play_t, prof_upd_t, prof_tot = agent.epoch_times()
    plane_flags = _plane_flags(agent)      # FP16 operand-plane scale misses during the warm-up / timed epochs (read outside the timed region)
    tr = {k: float(v) for k, v in zip(L.TR_NAMES, agent._tr_buf[-1].tolist())}
    graph_rollout = getattr(agent, '_rollout_graph', None) is not None
    mb_graph = getattr(agent, '_mb_graph_state', None) is not None
    env_steps = args.steps * NUM_ENVS * HORIZON * world
    value = env_steps / secs
    del agent, env
    torch.cuda.empty_cache()

    # e2e: host-resident simulator state, H2D every sim step, D2H of the step's train_result
    agent, env = _make_agent(torch, rank, world, 'host', seed=0)
    for _ in range(max(4, args.warmup)):       # >= 3: the rollout's CUDA graph is captured on the third play_steps call
        agent.update_epoch(); agent.train_epoch()
        if world > 1:
            agent._sync_stats()
    e2e_secs, host = _timed_epochs(torch, agent, args.steps, world, d2h=True)
    e2e_value = env_steps / e2e_secs
    plane_flags |= _plane_flags(agent)
    h2d = env.h2d_bytes_per_step * HORIZON
    d2h = host[0].numel() * 4 if host else 0
    del agent, env

    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16, sustained: the kernel is timed inside a long step)"
    if peak_tf is None:
        peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    achieved_tf = (fl.value / 1e12) / (tot_ms.value / 1e3) if tot_ms.value > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    nmb = args.steps * MINI_EPOCHS * (NUM_ENVS * HORIZON // MINIBATCH)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (3xFP16 scaled hi/lo tensor-core products, fp32 accumulate)" if GEMM_BACKEND == 2 else "f32 (3xTF32 tensor-core products, fp32 accumulate)",
        "data": "synthetic",
        "config": _workload_config(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_secs * 1e3 / args.steps,
                "note": "separate timed run through the agent API: every sim step's rigid-body state arrives from pinned host memory (H2D inside "
                        "the timed region, overlapping compute on the copy engine), the epoch's train_result series is read back (D2H); it can "
                        "come out slightly ahead of `value`, whose synthetic env refreshes its state with device-to-device copies on the compute stream"},
        "gpu_launches": int(launches),
        "gpu_launches_note": "kernels launched by libase_b200.so through host calls during the timed epochs; launches replayed from the captured CUDA "
                             "graphs (rollout: ~60 per sim step; minibatch update: ~100) are NOT re-counted -- the instrumented epoch launches " + str(int(nl.value)) + " tcgen05 GEMMs eagerly",
        "clocks": clk.summary(),
        "roofline": {"bound": "tensor", "kernel": "gemm_tc2_kernel (persistent CTA pair, tcgen05.mma.cta_group::2 kind::f16, 3 MMAs per product on scaled FP16 hi/lo planes; ~80 % of the GEMM time) + gemm_tc256_kernel / gemm_tc_kernel for ragged and narrow shapes" if GEMM_BACKEND == 2
                     else "gemm_tc_kernel (tcgen05.mma kind::tf32, 3xTF32)", "achieved": achieved_tf,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                     "peak_source": peak_src, "launches_timed": int(nl.value), "kernel_ms_per_step": tot_ms.value,
                     "kernel_share_of_step": (tot_ms.value / 1e3) / prof_tot if prof_tot > 0 else None,
                     "measured_in": "one extra untimed epoch with per-launch CUDA events (the timed epochs run uninstrumented)",
                     "frac_of_step_algorithmic": (nmb / args.steps * FLOP_PER_MINIBATCH + FLOP_ROLLOUT_PER_EPOCH) / 1e12 / (secs / args.steps) / peak_tf if (peak_tf and CONFIG == 3) else None,
                     "note": "achieved = algorithmic 2*M*N*K FLOPs of the timed launches / summed CUDA-event kernel time; fp32 parity "
                             "forces 3 MMAs per product (hi.hi + lo.hi + hi.lo): the ceiling against the bf16 peak is 1/3 with FP16 planes "
                             "(backend 2), 1/6 with TF32 planes (backend 1)",
                     "learner_tflops_algorithmic": nmb * FLOP_PER_MINIBATCH / 1e12 / (upd_t * args.steps) if upd_t > 0 else None},
        "split": {"play_time_s_last_step": play_t, "update_time_s_last_step": upd_t, "rollout_cuda_graph": graph_rollout,
                  "minibatch_cuda_graph": mb_graph, "instrumented_epoch_s": {"play": prof_play_t, "update": prof_upd_t}},
        "epoch_ms_rank0": value_epoch_ms,      # the K timed epochs one by one (rank 0's CUDA events): stationarity of the timed region
        "gradient_sum": ("one kernel with Adam over NVLink peer memory (csrc/peer.cu)" if peer_adam else "NCCL allreduce (csrc/comm.cu) + Adam") if world > 1 else "none (1 GPU)",
        "plane_status": plane_flags,      # ase_learner_plane_status after both runs: 0 = every FP16 plane scale prediction held
        "train_result_last": tr,
    }
    if world == 1 and CONFIG == 3:
        line["cpu_baseline"] = cpu_baseline()
    _emit(line)


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


# ------------------------------------------------------------------------------------------------ CPU legs
def _cpu_arm(reps, warm):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_bench
    threads = cpu_bench.fixed_threads()
    st = cpu_bench.make_state(NUM_ENVS, MINIBATCH, AMP_MINIBATCH)
    for _ in range(warm):
        cpu_bench.fraction_step(st, CPU_FRACTION)
    times = [cpu_bench.fraction_step(st, CPU_FRACTION) for _ in range(reps)]
    return threads, times, cpu_bench.fraction_description(NUM_ENVS, HORIZON, MINIBATCH, AMP_MINIBATCH, MINI_EPOCHS, CPU_FRACTION)


def cpu_baseline():
    threads, times, what = _cpu_arm(reps=2, warm=1)
    t = statistics.median(times)
    return {"value": NUM_ENVS * HORIZON / CPU_FRACTION / t, "unit": UNIT, "cores": threads, "kind": "port", "sample": what,
            "step_wall_s": {"min": min(times), "median": t}, "thread_policy": "fixed: min(32, cores in the affinity mask)"}


def run_reference(args):
    """The reference's own algorithm (CPU torch fp32) on the same workload: the oracle port (the reference is Python and /root/reference does
    not exist on the GPU box).  One step = an exact 1/16 of an epoch, timed for real (no extrapolation): ms_per_step is wall time."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads, times, what = _cpu_arm(reps=args.steps, warm=min(args.warmup, 1))
    t = statistics.median(times)
    v = NUM_ENVS * HORIZON / CPU_FRACTION / t
    cfgd = _workload_config(args.gpus)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgd,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": what,
                             "step_wall_s": {"min": min(times), "median": t, "all": times}, "thread_policy": "fixed: min(32, cores in the affinity mask)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "env_steps_per_step": NUM_ENVS * HORIZON // CPU_FRACTION,
            "note": "CPU arm does not scale with --gpus: one host runs the reference algorithm for one env shard; a step is 1/16 of an epoch "
                    "(value = 8192 env-steps / measured step time)"}
    _emit(line)


_REAL_STDOUT = None


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    while data:
        data = data[os.write(fd, data):]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 5], help="BASELINE.json config (3 = the metric's config)")
    args = ap.parse_args()
    global CONFIG, _REAL_STDOUT
    CONFIG = args.config
    # stdout carries exactly ONE line (the JSON).  Libraries print there too -- NCCL's "NCCL version ..." line under NCCL_DEBUG=VERSION ignores
    # NCCL_DEBUG_FILE -- so file descriptor 1 points at stderr while the benchmark runs and _emit() writes the line to the real stdout.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        _shutdown()


if __name__ == "__main__":
    main()
