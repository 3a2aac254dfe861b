import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import torch, numpy as np
import bench_configs as bc
from leniax_b200 import runner
def run(cfg, steps):
    fn = bc.config_d if cfg == 'D' else bc.config_e
    # monkeypatch timed to also record host issue time
    def timed(f, reps=2):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(reps): out = f()
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(cfg, steps, 'host issue ms/rep', (t1 - t0) * 1e3 / reps, 'gpu ms/rep', e0.elapsed_time(e1) / reps, 'gpu us/step', e0.elapsed_time(e1) / reps / steps * 1e3, flush=True)
        return e0.elapsed_time(e1) / reps, out
    bc.timed = timed
    fn(steps)
for s in (64, 256, 1024): run('D', s)
for s in (16, 64, 256): run('E', s)
