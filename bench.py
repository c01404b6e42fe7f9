#!/usr/bin/env python
"""bench.py - env-steps/sec on the foothold + obs + PPO hot path (sim stubbed), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, one process per GPU
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference on the host cores

One "step" is one OnPolicyRunner iteration: T=24 lock-step environment steps (state prep, foothold scoring,
rewards/reset, observations), 24 policy forwards, GAE, and PPO.update() = 5 epochs x 4 minibatches x (VAE step +
policy step).  env-steps/sec = world * N_envs * T * K / time, the quantity the reference logs as Perf/total_fps
(rsl_rl/runners/on_policy_runner.py:185).  Workload: BASELINE.json configs[1] (Lite3 stepping-stone heightmap, 4096
envs per GPU, ActorCriticDecoder = CE-net + 512-d terrain latent); the Isaac Gym call is replaced by synthetic
root/dof/contact/rigid-body tensors (SURVEY.md 8d distributions).

`value`  : simulator tensors already resident in HBM (a device-side pool refreshed by device-to-device copies).  From the second
           iteration on OnPolicyRunner replays the rollout (24 x 28 launches) as one CUDA graph (DTC_CUDA_GRAPH=0: eager launches).
`e2e`    : the same loop through the public Python API with the simulator tensors arriving from PINNED HOST memory every
           environment step (host->device copies inside the timed region: the 24 states of the next rollout travel on a copy stream
           into a device ring while the update runs, the graph moves ring[t] into the simulator tensors) and the iteration's
           statistics read back.
N > 1    : one process per GPU, environments sharded; the per-optimizer-step gradient all-reduce is the library's own kernel over
           NVLink peer memory (csrc/dtc_dp.cu; DTC_DP=nccl1 for the NCCL call it replaces).
`roofline`: the dominant kernel, k_gemm_tc2 (CTA-pair tcgen05 GEMM): algorithmic 2*M*N*K FLOPs / CUDA-event time around every
           launch of one extra iteration (side streams serialised for it), against the measured sustained bf16 tensor peak of
           MEASURED_PEAKS.json; `traffic` = DRAM bytes of one launch from the committed ncu capture (profiles/).  The whole GEMM
           family rides along.
`roofline_foothold`: BASELINE.json configs[4], the foothold-scoring kernel microbench at 16 384 environments (HBM bound by
           definition: 3048 B/env + the 3.9 MB map), timed alone with CUDA events, L2 flushed between launches; the kernel's
           time inside the training loop (4096 environments, warm L2) is reported beside it.
`cfg4`   : BASELINE.json configs[3] (8192 envs/GPU, rough-terrain curriculum map) as its own short run (N = 1 only).
`dp_check`: N > 1 only - a policy step sharded over the ranks (all-reduced gradients / world) against the same step on the whole
           minibatch on every rank: critic gradients must agree to fp32 round-off.
`cpu_baseline`: the CPU restatement of the reference (oracle/, pinned to golden vectors of the unmodified reference) on the host
           cores, bounded sample; `--impl reference` runs the same as its own arm (same `config`, the sample in `cpu_baseline`).
           The port runs 1.245x faster than the unmodified reference on the same inputs (tests/tools/port_vs_reference.py, measured in
           the build container where /root/reference exists: 3047 vs 2448 env-steps/s at 1024 envs on 8 cores).
Timing   : CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.  The per-iteration
           working set (717 MB rollout storage + its gathered copy) exceeds the 126 MB L2, so no explicit flush is used.
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

T_STEPS = 24
# dram__bytes_read.sum + dram__bytes_write.sum of one k_gemm_tc2 launch from the committed ncu --set full capture
# (profiles/r2_gemm_tc2_ncu_raw.csv: forward 24576 x 512 x 512, the most frequent shape: 102.84 MB read + 59.29 MB written;
# algorithmic 100.7 MB of operands (x and its TF32 companion) + 2 MB of weights + 100.7 MB of output, part of which is still
# L2-resident when the kernel ends)
TRAFFIC_GEMM_TC2 = 162128896
TRAFFIC_FOOTHOLD_16384 = 11802112  # profiles/r2_foothold_v6_ncu_raw.csv: 10.82 MB read + 0.98 MB written
PORT_OVER_REFERENCE = 1.245  # tests/tools/port_vs_reference.py (build container, 1024 envs, 8 cores)
BYTES_STATE_PER_ENV = (13 + 12 * 2 + 17 * 3 + 17 * 13) * 4  # root, dof, contact, rigid body


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def _config(envs_per_gpu, world):
    """The workload both arms are quoted on (the reference arm times a bounded sample of it, described in its cpu_baseline)."""
    return {"workload": "configs[1]: Lite3 stepping-stone heightmap, 4096 envs/GPU, ActorCriticDecoder (CE-net + 512-d terrain latent), "
                        "T=24, 5 epochs x 4 minibatches", "envs_per_gpu": envs_per_gpu, "rollout_len": T_STEPS, "epochs": 5, "minibatches": 4,
            "parallelism": f"dp{world}", "l2": "inputs larger than L2 (717 MB rollout storage per iteration), no explicit flush"}


# ------------------------------------------------------------------------------------------------ CUDA arm
def build_world(n_envs, rank, device, kind="stones"):
    import torch
    import dtc_b200  # noqa: F401
    from dtc_b200 import sim_stub
    from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg, Lite3DTCCfgPPO
    from dtc_b200.legged_gym.envs.lite3.lite3_dtc_config import class_to_dict
    from dtc_b200.rsl_rl.runners import OnPolicyRunner
    seed = 1000 + rank
    hs, tor = sim_stub.make_heightmap(kind, 0)
    layout = sim_stub.initial_env_layout(n_envs, tor, seed)
    fg = sim_stub.FakeGym(n_envs, device=device)
    cfg = Lite3DTCCfg()
    cfg.env.num_envs = n_envs
    env = LeggedRobotDTC(cfg, sim_device=device, gym=fg, height_samples=hs, terrain_origins=tor, layout=layout, seed=seed)
    g = torch.Generator().manual_seed(seed)
    pool_host = [{k: v.pin_memory() for k, v in sim_stub.synth_state(n_envs, layout[2], g).items()} for _ in range(8)]
    pool_dev = [{k: v.to(device) for k, v in s.items()} for s in pool_host]
    state = {"i": 0, "pool": pool_dev}

    def source():
        state["i"] = (state["i"] + 1) % 8
        return state["pool"][state["i"]]

    fg.source = source
    torch.manual_seed(1)  # identical initial policy on every rank (data parallel replicas)
    tc = class_to_dict(Lite3DTCCfgPPO())
    tc["runner"]["cuda_graph"] = os.environ.get("DTC_CUDA_GRAPH", "1") != "0"  # rollout replayed as one CUDA graph (A/B switch)
    runner = OnPolicyRunner(env, tc, log_dir=None, device=device)
    return env, fg, runner, state, pool_host, pool_dev


def timed(runner, iters, world, device):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    runner.learn(iters)
    e1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def foothold_microbench(device, peaks, N=16384, iters=20, warmup=3):
    """BASELINE.json configs[4]: the foothold-scoring kernel alone at 16 384 environments on the stepping-stone map."""
    import ctypes as C
    import torch
    from dtc_b200 import _lib as B, sim_stub
    from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg
    hs, tor = sim_stub.make_heightmap("stones", 0)
    layout = sim_stub.initial_env_layout(N, tor, 1)
    fg = sim_stub.FakeGym(N, device=device)
    cfg = Lite3DTCCfg()
    cfg.env.num_envs = N
    env = LeggedRobotDTC(cfg, sim_device=device, gym=fg, height_samples=hs, terrain_origins=tor, layout=layout, seed=1)
    g = torch.Generator(device=device).manual_seed(2)
    fg.load(sim_stub.synth_state(N, env.env_origins, g, device=device))
    env.reset()
    st0 = sim_stub.synth_state(N, env.env_origins, g, device=device)
    fg.queue.append(st0)
    env.step(torch.zeros(N, 12, device=device))
    fg.load(st0)  # the kernel sees one consistent simulator state (reset_idx rewrote the rows of terminated environments)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # 2 x L2
    stream = B.stream_ptr(torch.device(device))
    ts = []
    for i in range(warmup + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        B.check(env.lib.dtc_foothold_step(env._h, env.foothold_variant, C.c_void_p(0), stream), "dtc_foothold_step")
        e1.record()
        torch.cuda.synchronize(device)
        if i >= warmup:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    nbytes = 3048.0 * N + 3942400.0
    del env, fg, flush
    return {"kernel": "k_foothold_v6", "bound": "hbm", "achieved": round(nbytes / us / 1e3, 1), "peak": peaks["hbm"], "unit": "GB/s",
            "frac": round(nbytes / us / 1e3 / peaks["hbm"], 4), "us_per_launch": round(us, 2), "us_min": round(ts[0], 2), "envs": N,
            "launches_timed": iters, "algorithmic_bytes_per_launch": nbytes, "traffic": TRAFFIC_FOOTHOLD_16384,
            "workload": "configs[4]: 16384 envs x 4 legs x 45-candidate windows (7x7 lattice minus corners = every point within 0.16 m) "
                        "over the 1.5 m heightmap patch, stepping-stone map; L2 flushed (256 MB memset) before every launch",
            "traffic_note": "ncu --set full at 16384 envs: DRAM 10.82 MB read + 0.98 MB written per launch; the 46 MB of outputs stay "
                            "L2-resident for the consumer kernels, so DRAM traffic is below the algorithmic bytes; the kernel is issue-bound "
                            "(2.0 k warp instructions per environment), see profiles/README.md",
            "peak_source": peaks["source"]}


def cfg4_leg(device, rank, steps=3, warmup=2, n_envs=8192):
    """BASELINE.json configs[3]: rough-terrain curriculum map, 8192 envs per GPU, T = 24, 5 epochs x 4 minibatches."""
    import torch
    env, fg, runner, state, pool_host, pool_dev = build_world(n_envs, rank, device, kind="curriculum")
    runner.learn(warmup)
    ms = timed(runner, steps, 1, device)
    out = {"workload": "configs[3]: Lite3 rough-terrain curriculum map (stairs up / down, discrete obstacles, stepping stones; levels follow "
                       "the reference's curriculum rule), 8192 envs/GPU, T=24, 5 epochs x 4 minibatches", "envs_per_gpu": n_envs,
           "value": round(n_envs * T_STEPS * steps / (ms * 1e-3), 1), "unit": "env-steps/s", "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms / steps, 3), "mean_terrain_level": round(float(env.terrain_levels.float().mean()), 3)}
    del env, fg, runner, state, pool_host, pool_dev
    torch.cuda.empty_cache()
    return out


def dp_check(device, world, rank):
    """Sharded policy step == full-minibatch policy step (critic gradients, which have no batch-global statistic inside) and the
    KL sum that rides the same all-reduce; every rank builds the SAME synthetic rollout from one seed."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from dtc_b200 import _lib as B
    from dtc_b200.rsl_rl.algorithms import PPO
    from dtc_b200.rsl_rl.modules import ActorCriticDecoder
    from dtc_b200.rsl_rl.modules.actor_critic_decoder import STATE_KEYS
    from dtc_b200.rsl_rl.storage import RolloutStorage
    from dtc_b200.rsl_rl.utils import dp
    N, T = 64 * world, 24
    torch.manual_seed(5)
    ac = ActorCriticDecoder(53, 1389, 12).to(device)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        ac._flat.add_((torch.randn(ac._flat.numel(), generator=g) * 0.03).to(device) * (ac._flat != 0))
    ac._params_written()
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=4, entropy_coef=0.003, learning_rate=1e-3, schedule="adaptive", device=device)
    alg.init_storage(N, T, [53], [1389], [265], [12])
    st, tr = alg.storage, RolloutStorage.Transition()
    r = lambda *s: torch.randn(*s, generator=g).to(device)
    for t in range(T):
        tr.observations, tr.observation_histories, tr.privileged_observations = r(N, 53), r(N, 265), r(N, 1389)
        tr.base_vel, tr.next_observations, tr.actions = r(N, 3), r(N, 53), r(N, 12)
        tr.rewards, tr.dones = r(N), torch.zeros(N, device=device, dtype=torch.uint8)
        tr.values, tr.actions_log_prob = r(N), r(N) * 0.1 - 16.0
        tr.action_mean, tr.action_sigma = r(N, 12) * 0.1, torch.ones(N, 12, device=device)
        st.add_transitions(tr)
    st.returns.copy_(r(T, N, 1))
    st.advantages.copy_(r(T, N, 1))
    mbs = N * T // 4
    perm = torch.randperm(N * T, generator=g)
    batch = st.gather(perm.to(device))
    eps = r(mbs, 16).contiguous()
    lib, stream, hp, h = B.lib(), B.stream_ptr(torch.device(device)), alg._hparams(), ac._learner(mbs)
    b0, b1 = ac._table.ranges["policy_sync"]
    B.check(lib.dtc_ppo_step(h, C.byref(batch._c), 0, mbs, B.ptr(eps), 0, 0, C.byref(hp), 1, stream), "ppo_step full")
    full = ac._grads[b0:b1].clone()
    half = mbs // world
    e = eps[rank * half:(rank + 1) * half].contiguous()
    B.check(lib.dtc_ppo_step(h, C.byref(batch._c), rank * half, half, B.ptr(e), 0, 0, C.byref(hp), 1, stream), "ppo_step shard")
    dp.allreduce_sum_(ac._grads[b0:b1])
    shard = ac._grads[b0:b1].clone()
    kl_full, kl_sum = float(full[-4]), float(shard[-4])
    shard[:-4] *= 1.0 / world
    err = 0.0
    for k in STATE_KEYS:
        if k.startswith("critic_body."):
            idx = ac._idx[k] - b0
            a, b = full[idx].double(), shard[idx].double()
            err = max(err, float((a - b).abs().max() / a.abs().max().clamp_min(1e-30)))
    out = torch.tensor([err, abs(kl_sum - kl_full) / max(abs(kl_full), 1e-30)], device=device, dtype=torch.float64)
    dist.all_reduce(out, op=dist.ReduceOp.MAX)
    del alg, ac, st, batch
    torch.cuda.empty_cache()
    return {"critic_grad_max_rel_err": float(out[0]), "kl_sum_rel_err": float(out[1]), "tolerance": 3e-5, "rows_per_rank": half,
            "ok": bool(float(out[0]) <= 3e-5 and float(out[1]) <= 1e-3),
            "what": "dtc_ppo_step on each rank's shard + NCCL all-reduce / world vs the same step on the whole minibatch"}


def run_cuda(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))
    from dtc_b200 import _lib as B
    N = args.envs
    env, fg, runner, state, pool_host, pool_dev = build_world(N, rank, device)
    lib = B.lib()

    if args.warmup > 0:
        runner.learn(args.warmup)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = B.launch_count()
    ms = timed(runner, args.steps, world, device)
    launches = B.launch_count() - l0
    clocks = sampler.stop() if sampler else {}
    env_steps = world * N * T_STEPS * args.steps
    value = env_steps / (ms * 1e-3)

    # e2e: simulator tensors come from pinned host memory every env step; statistics are read back every iteration
    if args.profile_lite:  # launch-list runs under ncu: one timed iteration is all that is wanted
        ms_e2e, e2e_value = float("nan"), float("nan")
    else:
        state["pool"] = pool_host
        fg.enable_prefetch(True)  # host->device copies of step t+1 on a copy stream while step t's kernels run
        runner.reset_graph()      # the captured rollout baked the device pool in: one eager iteration, one capturing, then timed replays
        runner.learn(2)
        ms_e2e = timed(runner, args.steps, world, device)
        e2e_value = env_steps / (ms_e2e * 1e-3)
        fg.enable_prefetch(False)
        torch.cuda.synchronize(device)
        state["pool"] = pool_dev
        runner.reset_graph()

    # roofline of the dominant kernel (the GEMM family: > 90 % of the step), measured with CUDA events around every
    # launch of one extra iteration
    roof, fh, fh_loop = None, None, None
    if rank == 0 and not args.profile_lite:
        lib.dtc_profile_enable(1)
    if not args.profile_lite:
        runner.learn(1)  # every rank takes part (the optimizer steps all-reduce); only rank 0 records events (eager launches: the graph was reset)
    torch.cuda.synchronize(device)
    if rank == 0 and not args.profile_lite:
        import ctypes as C
        peaks = _peaks()
        flops, gms, fms, n_g, n_f = C.c_double(), C.c_double(), C.c_double(), C.c_int64(), C.c_int64()
        lib.dtc_profile_read(C.byref(flops), C.byref(gms), C.byref(n_g), C.byref(fms), C.byref(n_f))
        lib.dtc_profile_enable(0)
        fam = flops.value / (gms.value * 1e-3) / 1e12 if gms.value > 0 else 0.0
        pw, pms, pn = C.c_double(), C.c_double(), C.c_int64()
        lib.dtc_profile_kind(2, C.byref(pw), C.byref(pms), C.byref(pn))
        ach = pw.value / (pms.value * 1e-3) / 1e12 if pms.value > 0 else 0.0
        # ncu --set full of the same kernel (profiles/): DRAM bytes per launch on the learner's largest shape
        roof = {"kernel": "k_gemm_tc2 (tcgen05 cta_group::2 kind::tf32, error-compensated 3xTF32: forward / dgrad / split-K wgrad of the 256..752-wide layers)",
                "bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tensor_sustained"], "unit": "TFLOP/s",
                "frac": round(ach / peaks["tensor_sustained"], 4), "traffic": TRAFFIC_GEMM_TC2,
                "peak_source": peaks["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                "launches_per_step": pn.value, "kernel_ms_per_step": round(pms.value, 3),
                "algorithmic_flops_per_launch": round(pw.value / max(1, pn.value), 1),
                "share_of_gemm_family_time": round(pms.value / gms.value, 4) if gms.value > 0 else None,
                "gemm_family": {"launches_per_step": n_g.value, "ms_per_step": round(gms.value, 3), "tflops": round(fam, 2),
                                "algorithmic_flops_per_step": flops.value},
                "note": "algorithmic = 2*M*N*K fp32-equivalent FLOPs; the tensor pipe executes 3 TF32 MMAs per product (1e-5 parity), "
                        "i.e. %.1f TF32 TFLOP/s = %.3f of the TF32 dense peak taken as half the measured bf16 figure; per-launch CUDA "
                        "events with the step's side streams serialised" % (3 * ach, 3 * ach / (0.5 * peaks["tensor_sustained"]))}
        fh_bytes = 3048.0 * N + 3942400.0
        fh_us = fms.value * 1e3 / max(1, n_f.value)
        fh_loop = {"envs": N, "us_per_launch": round(fh_us, 2), "achieved": round(fh_bytes / (fh_us * 1e-6) / 1e9, 1),
                   "frac": round(fh_bytes / (fh_us * 1e-6) / 1e9 / peaks["hbm"], 4),
                   "note": "the same kernel inside the training loop (CUDA events around each of the step's launches, warm L2)"}
    cfg4 = dpc = None
    if not args.profile_lite:
        if world > 1:
            dpc = dp_check(device, world, rank)
        del runner, env, fg, state, pool_host, pool_dev
        torch.cuda.empty_cache()
        if rank == 0 and world == 1:  # single-GPU legs (under torchrun the other ranks are already tearing down)
            fh = foothold_microbench(device, _peaks())
            fh["in_loop"] = fh_loop
            if not args.no_cfg4:
                cfg4 = cfg4_leg(device, rank)
        elif rank == 0:
            fh = dict(fh_loop, kernel="k_foothold_v6", bound="hbm", unit="GB/s", peak=_peaks()["hbm"], traffic=None)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(sample_envs=args.cpu_envs, iters=5)  # ~2.3 s per iteration on 16 host cores: 10-12 s of CPU work

    if rank == 0:
        out = {
            "metric": "env-steps/sec (foothold+obs+PPO, sim stubbed)", "value": round(value, 1), "unit": "env-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(N, world),
            "e2e": {"value": round(e2e_value, 1), "unit": "env-steps/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                    "h2d_bytes_per_step": BYTES_STATE_PER_ENV * N * T_STEPS, "d2h_bytes_per_step": 16 * 8},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_foothold": fh, "cfg4": cfg4, "dp_check": dpc,
            "cpu_baseline": cpu,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def _oracle_world(n_envs, seed=1000):
    import torch
    import dtc_b200  # noqa: F401
    from dtc_b200 import lite3 as K, sim_stub
    from oracle import env_oracle as EO, learner_oracle as LO
    from oracle.rng import Live
    hs, tor = sim_stub.make_heightmap("stones", 0)
    layout = sim_stub.initial_env_layout(n_envs, tor, seed)
    fg = sim_stub.FakeGym(n_envs)
    g = torch.Generator().manual_seed(seed)
    pool = [sim_stub.synth_state(n_envs, layout[2], g) for _ in range(4)]
    st = {"i": 0}

    def source():
        st["i"] = (st["i"] + 1) % 4
        return pool[st["i"]]

    fg.source = source
    rng = Live(seed)
    env = EO.OracleEnv(K, n_envs, hs, layout, fg, rng)
    wenv = EO.OracleHistoryWrapper(env)
    torch.manual_seed(1)
    ac = LO.ActorCriticDecoder(53, 1389, 12, rng=rng)
    alg = LO.PPO(ac, num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
                 entropy_coef=0.003, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
                 desired_kl=0.01, rng=rng)
    alg.init_storage(n_envs, T_STEPS, [53], [1389], [265], [12])
    wenv.reset()
    return wenv, alg, LO


def _cpu_iterations(n_envs, iters, warmup):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wenv, alg, LO = _oracle_world(n_envs)
    obs = wenv.get_observations()
    for _ in range(warmup):
        obs, _ = LO.learn_iteration(wenv, alg, obs, T_STEPS)
    t0 = time.perf_counter()
    for _ in range(iters):
        obs, _ = LO.learn_iteration(wenv, alg, obs, T_STEPS)
    dt = time.perf_counter() - t0
    return n_envs * T_STEPS * iters / dt, dt, cores


def cpu_baseline(sample_envs=512, iters=1):
    v, dt, cores = _cpu_iterations(sample_envs, iters, warmup=0)
    return {"value": round(v, 1), "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{iters} full iteration(s) (24 env steps + GAE + 5x4 minibatch update) at {sample_envs} envs = "
                      f"{sample_envs * T_STEPS * iters} env-steps in {dt:.1f} s; oracle/ = CPU restatement of the reference pinned "
                      f"to golden vectors recorded from the unmodified reference",
            "port_over_reference": PORT_OVER_REFERENCE}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_envs
    v, dt, cores = _cpu_iterations(n, args.steps, args.warmup)
    sample = (f"each step = one full iteration (24 env steps + GAE + 5x4 minibatch update) on a bounded sample of {n} of the config's "
              f"{args.envs} environments ({n * T_STEPS} env-steps), torch CPU with {cores} threads; oracle/ = CPU restatement of the reference "
              f"pinned to golden vectors of the unmodified reference, which it outruns by {PORT_OVER_REFERENCE}x on the same inputs "
              f"(tests/tools/port_vs_reference.py)")
    out = {"impl": "reference", "metric": "env-steps/sec (foothold+obs+PPO, sim stubbed)", "value": round(v, 1), "unit": "env-steps/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": _config(args.envs, args.gpus),
           "cpu_baseline": {"value": round(v, 1), "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
                            "sample_envs": n, "port_over_reference": PORT_OVER_REFERENCE},
           "e2e": {"value": round(v, 1), "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dtc_b200", choices=["dtc_b200", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--cpu-envs", type=int, default=1024, help="environments of the bounded CPU-baseline sample")
    ap.add_argument("--ref-envs", type=int, default=1024, help="environments per step of the --impl reference arm (BASELINE.md section 2's stand-in)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the configs[3] leg (8192 envs, curriculum map)")
    ap.add_argument("--profile-lite", action="store_true",
                    help="for launch lists under ncu (never a bench value): warm-up as given, no e2e / roofline / CPU legs")
    args = ap.parse_args()
    if args.profile_lite:
        args.no_cpu_baseline = True
    if args.warmup < 3 and args.impl == "dtc_b200" and not args.profile_lite:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
