"""auto_reset flag on, but (almost) no env terminates: 20-step graphs from a fresh reset, each replay timed on its own"""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pkg = importlib.import_module("marl-dmfb_b200")
N, SUB = 65536, int(os.environ.get("TK_SUB", "1"))
W, A, DEG = (50, 10, True) if os.environ.get("TK_CFG", "c2") == "c3" else (20, 10, False)
env = pkg.BatchedDMFB(N, W, W, A, fov=9, b_degrade=DEG, per_degrade=1.0, device="cuda:0", seed=1234, sub_batches=SUB)
slots = 16
obs = torch.empty(slots + 1, N, 10, env.D, dtype=torch.int8, device="cuda:0")
act = torch.randint(0, 5, (slots, N, 10), device="cuda:0", dtype=torch.int8)
s = torch.cuda.Stream()
for ar in (False, True):
    with torch.cuda.stream(s):
        env.reset()
        for i in range(3): env.step(act[i], auto_reset=ar, out=obs[i + 1], join=SUB == 1)
        env.join(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(slots): env.step(act[i], auto_reset=ar, out=obs[i + 1], join=SUB == 1)
            env.join()
        ts = []
        for rep in range(6):
            env.reset(); g.replay(); s.synchronize()   # warm, busy start
            env.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); g.replay(); g.replay(); e1.record(s); s.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / (2 * slots))
    print(f"sub={SUB} auto_reset={ar}: no terminations, {sorted(ts)[len(ts)//2]:.2f} us per step (32 steps from reset)")
