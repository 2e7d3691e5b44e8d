import os, sys, subprocess, json
sys.path.insert(0, "/root/repo")
code = r'''
import os, sys, torch
sys.path.insert(0, "/root/repo")
from air_hockey_training_environment_b200 import BatchedAirHockey
n, K = 1 << 20, 8
env = BatchedAirHockey(n, seed=1234)
out = env.make_outputs(K, ("reward", "done", "score", "game_over", "obs_next"))
acts = [torch.randint(0, 4, (K, n), dtype=torch.int8, device="cuda") for _ in range(4)]
for i in range(260): env.step(acts[i % 4], K=K, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(200): env.step(acts[i % 4], K=K, out=out)
b.record(); torch.cuda.synchronize()
print(env.launch_info()["blocks"], a.elapsed_time(b) / 200 * 1000)
'''
for blocks in sys.argv[1:]:
    e = dict(os.environ); e["AH_GRID_BLOCKS"] = blocks
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    print("AH_GRID_BLOCKS", blocks, "->", r.stdout.strip(), r.stderr.strip()[-200:])
