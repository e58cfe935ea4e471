"""C3: is the staggered auto-reset number slower because of the resets, or because the chip has started to degrade?"""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("marl-dmfb_b200")
N, SUB, W, A, T = 65536, int(os.environ.get("TK_SUB", "4")), 50, 10, 200
env = pkg.BatchedDMFB(N, W, W, A, fov=9, b_degrade=True, per_degrade=1.0, device="cuda:0", seed=1234, sub_batches=SUB)
slots = 16
obs = torch.empty(slots + 1, N, A, env.D, dtype=torch.int8, device="cuda:0")
act = torch.randint(0, 5, (slots, N, A), device="cuda:0", dtype=torch.int8)
s = torch.cuda.Stream()

def graph_time(fn, n, reps=5):
    with torch.cuda.stream(s):
        for i in range(3): fn(i)
        env.join(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n): fn(i)
            env.join()
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps): g.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n)

def frac():
    env.join(); torch.cuda.synchronize()
    return float((env.health != 1.0).float().mean())

env.reset()
print(f"fresh chip: no reset {graph_time(lambda i: env.step(act[i % slots], out=obs[i % slots + 1], join=SUB == 1), slots):.2f} us, degraded cells {frac():.4f}")
for rnd in range(3):
    env.step_count.copy_(torch.arange(N, device='cuda:0', dtype=torch.int32) % T) if rnd == 0 else None
    us = graph_time(lambda i: env.step(act[i % slots], auto_reset=True, out=obs[i % slots + 1], join=SUB == 1), T)
    f = frac()
    env.reset()          # new=False: the chips keep their health
    us0 = graph_time(lambda i: env.step(act[i % slots], out=obs[i % slots + 1], join=SUB == 1), slots)
    print(f"after {1203 * (rnd + 1) + 99 * rnd} more steps: staggered auto-reset {us:.2f} us; degraded cells {f:.4f}; "
          f"no-reset step on these chips {us0:.2f} us")
    env.step_count.copy_(torch.arange(N, device='cuda:0', dtype=torch.int32) % T)
