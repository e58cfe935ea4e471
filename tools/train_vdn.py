#!/usr/bin/env python
"""End-to-end VDN / QMIX training on GPU-resident envs.

  python tools/train_vdn.py dmfb --envs 4096 --iters 50                   # BASELINE config #1 batched
  python tools/train_vdn.py meda --envs 2048 --iters 20                   # config #4: MEDA 30x60, 4 droplets, fov 19
  python tools/train_vdn.py dmfb --alg qmix --envs 2048 --iters 50        # mixer over the global state (get_state kernel)
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_vdn.py dmfb --envs 32768 \
         --json profiles/r02_config5_8gpu.json                            # config #5: 256K envs over 8 B200s

Batched counterpart of `python train.py {dmfb,meda} --drop_num=4` (train.py:32-93): every iteration collects one
lock-step episode from each of the rank's envs (no host round trip in the env path; the step kernel writes the
observations straight into the episode buffer), stores them in the device replay buffer, and runs `--train-time` VDN
updates whose gradients are all-reduced over NCCL in one flat bucket.  Defaults follow the reference's
data-{dmfb,meda}/TrainParas/4d.yaml and common/arguments.py:57-81; the MEDA env is MEDAEnv with its base observation,
which is what common/config.py:10-16 returns for the default `--version 0.2`.

Per-phase device time (CUDA events): env_step, policy_forward (batched CRNN + epsilon-greedy), rollout_other (episode
bookkeeping), learner (sample + forward/backward + clip + Adam), grad_allreduce (inside learner)."""
import argparse
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULTS = {   # data-*/TrainParas/4d.yaml + arguments.py set_default
    "dmfb": dict(width=10, length=10, fov=9, hyper_hidden_dim=24, batch_size=128, buffer_size=5000, train_time=1,
                 anneal_steps=150000, grad_norm_clip=9.0),
    "meda": dict(width=30, length=60, fov=19, hyper_hidden_dim=32, batch_size=64, buffer_size=10000, train_time=2,
                 anneal_steps=300000, grad_norm_clip=10.0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", nargs="?", default="dmfb", choices=["dmfb", "meda"])
    ap.add_argument("--alg", choices=["vdn", "qmix"], default="vdn")
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--length", type=int, default=None)
    ap.add_argument("--drop-num", type=int, default=4)
    ap.add_argument("--fov", type=int, default=None)
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--buffer-size", type=int, default=None)
    ap.add_argument("--train-time", type=int, default=None, help="learner updates per iteration")
    ap.add_argument("--anneal-steps", type=int, default=None)
    ap.add_argument("--tf32", action="store_true", help="TF32 matmuls / convolutions in the policy and learner")
    ap.add_argument("--bf16-rollout", action="store_true",
                    help="the rollout's policy forward under bf16 autocast, channels-last (the learner stays fp32)")
    ap.add_argument("--save-dir", default="")
    ap.add_argument("--json", default="", help="write the run summary (throughput, per-phase ms) to this file")
    args = ap.parse_args()
    d = DEFAULTS[args.name]
    for k in ("width", "length", "fov", "batch_size", "buffer_size", "train_time", "anneal_steps"):
        if getattr(args, k) is None:
            setattr(args, k, d[k])
    rank, world, local = (int(os.environ.get(k, v)) for k, v in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    P = importlib.import_module("marl-dmfb_b200")
    if args.name == "dmfb":
        env = P.BatchedDMFB(args.envs, args.width, args.length, args.drop_num, fov=args.fov, device=dev, seed=1234,
                            env_base=rank * args.envs)
    else:
        if args.alg == "qmix":
            raise SystemExit("qmix needs get_state, which the MEDA env does not have (nor does the reference's)")
        env = P.BatchedMEDA(args.envs, args.width, args.length, args.drop_num, fov=args.fov, obs_version=0, device=dev,
                            seed=1234, env_base=rank * args.envs)
    info = env.get_env_info()
    timer = P.PhaseTimer()
    qmix = args.alg == "qmix"
    state_dim = 3 * args.width * args.length if qmix else 0     # getglobalobs (dmfb.py:368-392), flattened
    kw = dict(world_size=world, seed=0, hyper_hidden_dim=d["hyper_hidden_dim"], grad_norm_clip=d["grad_norm_clip"], timer=timer)
    if qmix:
        learner = P.QMIXLearner(info["obs_shape"], info["n_agents"], info["n_actions"], state_dim, dev, **kw)
    else:
        learner = P.VDNLearner(info["obs_shape"], info["n_agents"], info["n_actions"], dev, **kw)
    agents = P.BatchedAgents(learner.eval_rnn, info["n_agents"], info["n_actions"], dev, seed=100 + rank,
                             autocast_dtype=torch.bfloat16 if args.bf16_rollout else None)
    worker = P.BatchedRolloutWorker(env, agents, anneal_steps=args.anneal_steps, record_state=qmix, timer=timer)
    T, A, D, NA = info["episode_limit"], info["n_agents"], info["obs_shape"][-1], info["n_actions"]
    buf = P.ReplayBufferGPU(max(args.buffer_size, args.envs), T, A, D, NA, dev, seed=200 + rank, state_dim=state_dim)
    ep = P.EpisodeBatch(args.envs, T, A, D, NA, dev, state_dim=state_dim)
    train_step, env_steps, live_steps = 0, 0, 0
    phases = {}
    t0 = None
    for it in range(args.iters):
        if it == 1:                                  # iteration 0 is warm-up (cuDNN autotune, allocator growth)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            timer.summary()
            t0, env_steps, live_steps = time.time(), 0, 0
        with timer("rollout_total"):
            ep, stats = worker.generate_episodes(batch=ep)
        with timer("learner"):
            buf.store_episodes(ep)
            for _ in range(args.train_time):
                loss = learner.learn(buf.sample(min(buf.current_size, args.batch_size)), train_step)
                train_step += 1
        env_steps += int(stats["steps"].sum().item())            # the reference's time_steps (failures = episode_limit)
        live_steps += int((~ep.padded).sum().item())             # env steps really executed
        if rank == 0:
            print(json.dumps({"iter": it, "loss": float(loss), "epsilon": worker.epsilon,
                              "mean_reward": float(stats["reward"].mean()), "success_rate": float(stats["success"].float().mean()),
                              "mean_steps": float(stats["steps"].float().mean())}), flush=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    wall = time.time() - (t0 or time.time())
    ph = timer.summary()
    iters = max(args.iters - 1, 1)
    ph["rollout_other"] = ph.get("rollout_total", 0.0) - ph.get("env_step", 0.0) - ph.get("policy_forward", 0.0)
    summary = {
        "what": f"end-to-end {args.alg.upper()} training on GPU-resident {args.name} envs (tools/train_vdn.py)",
        "config": {"env": args.name, "chip": [args.width, args.length], "droplets": args.drop_num, "fov": args.fov,
                   "envs_per_gpu": args.envs, "gpus": world, "total_envs": args.envs * world, "episode_limit": T,
                   "batch_size": args.batch_size, "train_time": args.train_time, "tf32": bool(args.tf32), "bf16_rollout": bool(args.bf16_rollout)},
        "iters_timed": iters, "wall_s": wall,
        "env_steps_per_s_executed": world * live_steps / max(wall, 1e-9),
        "agent_steps_per_s_executed": world * live_steps * A / max(wall, 1e-9),
        "time_steps_per_s_reference_accounting": world * env_steps / max(wall, 1e-9),
        "learner_updates_per_s": iters * args.train_time / max(wall, 1e-9),
        "phase_ms_per_iter_rank0": {k: v / iters for k, v in sorted(ph.items())},
        "final": {"loss": float(loss), "epsilon": worker.epsilon, "success_rate": float(stats["success"].float().mean())},
    }
    if rank == 0:
        print(json.dumps(summary), flush=True)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(summary, f, indent=1)
        if args.save_dir:
            learner.save_model(args.save_dir, 0)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
