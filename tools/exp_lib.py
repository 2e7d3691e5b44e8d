import os, sys, subprocess
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
for i in range(300): env.step(acts[i % 4], K=K, out=out)
b.record(); torch.cuda.synchronize()
us8 = a.elapsed_time(b) / 300 * 1000
out1 = env.make_outputs(1, ("reward", "done", "score", "game_over", "obs_next"))
a.record()
for i in range(300): env.step(acts[i % 4][0], K=1, out=out1)
b.record(); torch.cuda.synchronize()
print(env.launch_info(), "K8 us", round(us8, 2), "K1 us", round(a.elapsed_time(b) / 300 * 1000, 2), env.stats()["frames"])
'''
for lib in sys.argv[1:]:
    e = dict(os.environ)
    if lib != "default": e["AH_B200_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True)
    print(os.path.basename(lib), "->", r.stdout.strip(), r.stderr.strip()[-300:])
