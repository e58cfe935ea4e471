#!/usr/bin/env python
"""End-to-end VDN / QMIX training on GPU-resident envs (BASELINE config #5 when launched on 8 GPUs).

  python tools/train_vdn.py --envs 4096 --iters 50
  python tools/train_vdn.py --alg qmix --envs 2048 --iters 50      # mixer over the global state (get_state kernel)
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_vdn.py --envs 32768

Batched counterpart of `python train.py dmfb --drop_num=4 --fov=9` (train.py:32-93): every iteration collects one
lock-step episode from each of the rank's envs (no host round trip in the env path), stores them in the device replay
buffer, and runs `--train-time` VDN updates whose gradients are all-reduced over NCCL in one flat bucket."""
import argparse
import importlib
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--alg", choices=["vdn", "qmix"], default="vdn")
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--width", type=int, default=10)
    ap.add_argument("--length", type=int, default=10)
    ap.add_argument("--drop-num", type=int, default=4)
    ap.add_argument("--fov", type=int, default=9)
    ap.add_argument("--batch-size", type=int, default=128)
    ap.add_argument("--buffer-size", type=int, default=5000)
    ap.add_argument("--train-time", type=int, default=1)
    ap.add_argument("--anneal-steps", type=int, default=150000)
    ap.add_argument("--save-dir", default="")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    P = importlib.import_module("marl-dmfb_b200")
    env = P.BatchedDMFB(args.envs, args.width, args.length, args.drop_num, fov=args.fov, device=dev, seed=1234,
                        env_base=rank * args.envs)
    info = env.get_env_info()
    qmix = args.alg == "qmix"
    state_dim = 3 * args.width * args.length if qmix else 0     # getglobalobs (dmfb.py:368-392), flattened
    if qmix:
        learner = P.QMIXLearner(info["obs_shape"], info["n_agents"], info["n_actions"], state_dim, dev, world_size=world, seed=0)
    else:
        learner = P.VDNLearner(info["obs_shape"], info["n_agents"], info["n_actions"], dev, world_size=world, seed=0)
    agents = P.BatchedAgents(learner.eval_rnn, info["n_agents"], info["n_actions"], dev, seed=100 + rank)
    worker = P.BatchedRolloutWorker(env, agents, anneal_steps=args.anneal_steps, record_state=qmix)
    buf = P.ReplayBufferGPU(max(args.buffer_size, args.envs), info["episode_limit"], info["n_agents"], info["obs_shape"][-1],
                            info["n_actions"], dev, seed=200 + rank, state_dim=state_dim)
    ep = P.EpisodeBatch(args.envs, info["episode_limit"], info["n_agents"], info["obs_shape"][-1], info["n_actions"], dev,
                        state_dim=state_dim)
    train_step, t0, env_steps = 0, time.time(), 0
    for it in range(args.iters):
        ep, stats = worker.generate_episodes(batch=ep)
        buf.store_episodes(ep)
        env_steps += int(stats["steps"].sum().item())
        for _ in range(args.train_time):
            loss = learner.learn(buf.sample(min(buf.current_size, args.batch_size)), train_step)
            train_step += 1
        if rank == 0:
            print(json.dumps({"iter": it, "loss": float(loss), "epsilon": worker.epsilon,
                              "mean_reward": float(stats["reward"].mean()), "success_rate": float(stats["success"].float().mean()),
                              "mean_steps": float(stats["steps"].float().mean()),
                              "env_steps_per_s": world * env_steps / (time.time() - t0)}), flush=True)
    if rank == 0 and args.save_dir:
        learner.save_model(args.save_dir, 0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
