"""Where does a Rainbow training step spend its time?  (dev tool, GPU box)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from pfrl_b200.replay_buffers import PrioritizedReplayBuffer
from pfrl_b200.envs import SyntheticAtariVectorEnv

dev = torch.device('cuda', 0)
cap, B = 200000, 512
buf = PrioritizedReplayBuffer(cap, alpha=0.5, beta0=0.4, betasteps=None, normalize_by_max="memory",
                              num_steps=3, device=0, max_batch=512, part_capacity=cap + 4096)
g = torch.Generator(device=dev); g.manual_seed(0)
rng = np.random.RandomState(0)
done = 0
while done < cap + 2:
    m = min(1 << 16, cap + 2 - done)
    fr = torch.randint(0, 256, (m + 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
    term = np.zeros(m, bool); term[-1] = True
    buf.append_trajectory(fr, rng.randint(0, 18, m).astype(np.int64), rng.randint(-1, 2, m).astype(float), term)
    done += m
torch.backends.cudnn.allow_tf32 = False
agent = bench.make_rainbow_agent(buf, 0, B)
env = SyntheticAtariVectorEnv(16, device=dev, seed=1)
bench.rainbow_loop(agent, env, 3)
torch.cuda.synchronize()

def timeit(fn, n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3

obs = env.reset()
print('batch_act ms', timeit(lambda: agent.batch_act(obs), 20))
print('env.step ms', timeit(lambda: env.step(None), 20))
def upd():
    agent.update(agent.replay_buffer.sample(B))
print('sample+update ms', timeit(upd, 20))
def smp():
    e = agent.replay_buffer.sample(B); agent.replay_buffer.update_errors(torch.ones(B, device=dev))
print('sample+update_errors ms', timeit(smp, 20))
from pfrl_b200.replay_buffer import batch_experiences
def smpg():
    e = agent.replay_buffer.sample(B); b = batch_experiences(e, dev, agent.phi, 0.99); agent.replay_buffer.update_errors(torch.ones(B, device=dev))
print('sample+gather+update_errors ms', timeit(smpg, 20))
e = agent.replay_buffer.sample(B); batch = batch_experiences(e, dev, agent.phi, 0.99); agent.replay_buffer.update_errors(torch.ones(B, device=dev))
def fwdbwd():
    loss, d = agent._compute_loss(dict(batch), want_errors=True)
    agent.optimizer.zero_grad(); loss.backward(); agent.optimizer.step()
print('loss fwd+bwd+adam ms', timeit(fwdbwd, 20))
def fwd():
    with torch.no_grad(): agent.model(batch['state'])
print('one forward B=512 ms', timeit(fwd, 20))
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
print('one forward B=512 TF32 ms', timeit(fwd, 20))
print('loss fwd+bwd+adam TF32 ms', timeit(fwdbwd, 20))
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
# observe path: appends for 16 envs
o2, r, d, info = env.step(None)
agent.batch_act(o2)
def obs_only():
    o3, r, d, info = env.step(None)
    agent.batch_last_obs = list(o2); agent.batch_last_action = [0] * 16
    ri = agent.replay_updater.update_interval; agent.replay_updater.update_interval = 10 ** 9
    agent.batch_observe(o3, r, d, np.zeros(16, bool))
    agent.replay_updater.update_interval = ri
print('env.step + observe(append only) ms', timeit(obs_only, 20))

# ---- kernel-level breakdown of one update (torch profiler) -------------------
from torch.profiler import profile, ProfilerActivity
for bench_flag in (False, True):
    torch.backends.cudnn.benchmark = bench_flag
    for _ in range(3): fwdbwd()
    print('cudnn.benchmark', bench_flag, 'loss fwd+bwd+adam ms', timeit(fwdbwd, 20))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): fwdbwd()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
