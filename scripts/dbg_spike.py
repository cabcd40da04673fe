import sys, gc, time, torch
sys.path.insert(0, "/root/repo")
import bench
from starst3r_b200 import gs
dev = torch.device("cuda:0")
params, states, truth, cams = bench.make_workload(dev, 0)
plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
def run(tag, n=24):
    evs = []
    host = []
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        gs.train_step(params, states, truth, cams, bench.W, bench.H, i + 1, plan=plan)
        e1.record()
        host.append((time.perf_counter() - t0) * 1e3)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    print(tag, "gpu", [round(a.elapsed_time(b), 2) for a, b in evs])
    print(tag, "host", [round(h, 2) for h in host])
run("warm")
run("gc-on")
gc.disable()
run("gc-off")
gc.enable()
