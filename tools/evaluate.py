#!/usr/bin/env python
"""Batched counterparts of the reference's evaluation drivers, on GPU-resident envs.

  python tools/evaluate.py dmfb --drop-num 10 --chip-size 20 --evaluate-task 100          # evaluate.py:7-25
  python tools/evaluate.py dmfb --chip-size 50 --drop-num 10 --degrade-sweep --evaluate-epoch 40 --chips 5
                                                                                           # evaDegre.py:14-56

evaluate: `evaluate_task` greedy episodes on `evaluate_task` independent chips in ONE lock-step rollout; prints the
averages rollout.py:69-85 returns (reward, steps with failures charged episode_limit, constraints, success rate).

degrade sweep: `chips` independent degrading chips (b_degrade=True, per_degrade=1.0, evaDegre.py:37-38).  The reference
runs the tasks of an epoch sequentially on one chip so that wear accumulates; here each chip is one env, an epoch is
`evaluate_task` consecutive episodes on it, and all chips advance in lock step.  Health is snapshotted at the start of every epoch (evaDegre.py:21) and the four arrays are written like
evaDegre.py:52-56: rewards/steps/success (chips, epochs) and health (chips, epochs, W, L).
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", choices=["dmfb", "meda"])
    ap.add_argument("--drop-num", "-d", type=int, default=4)
    ap.add_argument("--chip-size", "--width", "-w", type=int, default=None)
    ap.add_argument("--length", "-l", type=int, default=None)
    ap.add_argument("--fov", type=int, default=None)
    ap.add_argument("--evaluate-task", type=int, default=100)
    ap.add_argument("--evaluate-epoch", type=int, default=20)
    ap.add_argument("--degrade-sweep", action="store_true")
    ap.add_argument("--chips", type=int, default=5)
    ap.add_argument("--model", default="", help="rnn_net_params.pkl written by the reference or by tools/train_vdn.py")
    ap.add_argument("--out", default="DegreData")
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    P = importlib.import_module("marl-dmfb_b200")
    dev = torch.device("cuda:0")
    if args.name == "dmfb":                                  # arguments.py:57-81 defaults
        W = args.chip_size or 10
        L = args.length or W
        fov = args.fov or 9
        make = lambda n, deg: P.BatchedDMFB(n, W, L, args.drop_num, fov=fov, b_degrade=deg, per_degrade=1.0,  # noqa: E731
                                            device=dev, seed=args.seed)
    else:
        W = args.chip_size or (80 if args.drop_num == 10 else 30)
        L = args.length or (80 if args.drop_num == 10 else 60)
        fov = args.fov or 19
        make = lambda n, deg: P.BatchedMEDA(n, W, L, args.drop_num, fov=fov, b_degrade=deg, per_degrade=1.0,  # noqa: E731
                                            obs_version=2, device=dev, seed=args.seed)
    n_envs = args.chips if args.degrade_sweep else args.evaluate_task
    env = make(n_envs, args.degrade_sweep)
    info = env.get_env_info()
    net = P.CRNN(info["obs_shape"], info["n_actions"]).to(dev)
    if args.model:
        net.load_state_dict(torch.load(args.model, map_location=dev))
    else:
        print("no --model given: evaluating a randomly initialised policy")
    worker = P.BatchedRolloutWorker(env, P.BatchedAgents(net, info["n_agents"], info["n_actions"], dev, seed=args.seed))
    t0 = time.time()
    if not args.degrade_sweep:
        _, st = worker.generate_episodes(evaluate=True)
        print("time:", time.time() - t0)
        print("The average total_rewards is  {}".format(float(st["reward"].mean())))
        print("The average total_steps is: {}".format(float(st["steps"].float().mean())))
        print("The average constraints is: {}".format(float(st["constraints"].float().mean())))
        print("The successful rate is: {}".format(float(st["success"].float().mean())))
        return
    E = args.evaluate_epoch
    rewards, steps, success = (np.zeros((n_envs, E)) for _ in range(3))
    health = np.zeros((n_envs, E, W, L))
    for epoch in range(E):
        health[:, epoch] = env.health.cpu().numpy()                           # evaDegre.py:21
        acc = {k: torch.zeros(n_envs, device=dev) for k in ("reward", "steps", "success")}
        for _ in range(args.evaluate_task):
            _, st = worker.generate_episodes(evaluate=True)
            for k in acc:
                acc[k] += st[k].float()
        rewards[:, epoch], steps[:, epoch], success[:, epoch] = (
            (acc[k] / args.evaluate_task).cpu().numpy() for k in ("reward", "steps", "success"))
        print(f"epoch {epoch}: success {success[:, epoch].mean():.3f} steps {steps[:, epoch].mean():.1f} "
              f"min health {health[:, epoch].min():.3g}", flush=True)
    path = os.path.join(args.out, "{}by{}-{}d0b".format(W, L, args.drop_num))
    os.makedirs(path, exist_ok=True)
    for name, arr in (("rewards", rewards), ("steps", steps), ("success", success), ("health", health)):
        np.save(os.path.join(path, name + ".npy"), arr)
    print("wrote", path, "in", time.time() - t0, "s")


if __name__ == "__main__":
    main()
