import os, sys, subprocess
code = r'''
import os, sys, torch
sys.path.insert(0, "/root/repo")
from air_hockey_training_environment_b200 import BatchedAirHockey
n = 1 << 20
env = BatchedAirHockey(n, seed=1234)
acts = [torch.randint(0, 4, (n,), dtype=torch.int8, device="cuda") for _ in range(4)]
out1 = env.make_outputs(1, ("reward", "done", "score", "game_over", "obs_next"))
for i in range(2000): env.step(acts[i % 4], K=1, out=out1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(500): env.step(acts[i % 4], K=1, out=out1)
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / 500 * 1000
print("K1 us", round(us, 2), "GB/s", round(n * 176 / us / 1e3, 1), env.stats()["frames"], env.puck[:2].tolist())
'''
for tag, ev in (("prefetch", {}), ("no-prefetch", {"AH_NO_PREFETCH": "1"})):
    e = dict(os.environ); e.update(ev)
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    print(tag, "->", r.stdout.strip(), r.stderr.strip()[-300:])
