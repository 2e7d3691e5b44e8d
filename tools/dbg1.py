import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import oracle
from air_hockey_training_environment_b200 import BatchedAirHockey
sys.path.insert(0, "/root/repo/tests")
from test_gpu_parity import put_state, get_state13, get_scores
for dtype in (torch.float32, torch.float64):
    env = BatchedAirHockey(8, dtype=dtype)
    P = torch.tensor([[340, 230], [300, 190],[10, 300], [100, 300],[890, 300], [800, 300],[10, 240], [900, 240]], dtype=dtype)
    env.puck[:, 0:2] = P.to(env.device)
    reward, score, done = env.rewards()
    print(dtype, done.tolist(), score.tolist(), reward.tolist())
n, T = 4096, 100
rng = np.random.default_rng(5)
actions = rng.integers(0, 4, (T, n)).astype(np.int8)
table = rng.uniform(-10, 10, (n, 16, 2)).astype(np.float32).astype(np.float64)
env = BatchedAirHockey(n, dtype=torch.float32)
dev = env.device
tab = torch.as_tensor(table).to(dev)
st, sc = oracle.initial_state(n)
cur = np.zeros(n, np.int32)
warm = rng.integers(0, 4, (400, n)).astype(np.int8)
oracle.run_frames(st, sc, 400, actions=warm, reset_table=None, seed=3, reset_cursor=cur, want=())
np.set_printoptions(precision=6, linewidth=200, suppress=True)
big = 0
for t in range(T):
    st = st.astype(np.float32).astype(np.float64)
    before = st.copy()
    cur[:] = 0
    put_state(env, st, sc, cur)
    out = env.make_outputs(1, ("reward", "done", "score", "game_over"))
    env.step(torch.as_tensor(actions[t]).to(dev), K=1, out=out, reset_table=tab)
    ref = oracle.run_frames(st, sc, 1, actions=actions[t:t + 1], reset_table=table, reset_cursor=cur, want=("reward", "score", "done", "game_over"))
    got = get_state13(env)
    d = np.abs(got[:, :12] - st[:, :12]).max(axis=1)
    ev = (ref["reward"][0] != out.reward[0].cpu().numpy().astype(np.int32))
    for i in np.nonzero(d > 5e-4)[0][:3]:
        big += 1
        if big < 12:
            print("t", t, "env", i, "maxdiff", d[i], "evmismatch", ev[i]); print(" before", before[i]); print(" oracle", st[i]); print(" gpu   ", got[i])
print("count big", big)
