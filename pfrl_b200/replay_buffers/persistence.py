"""save / load of device replay buffers (AbstractReplayBuffer.save/load,
pfrl/replay_buffer.py:69-85).  The live window is exported through the gather
kernel in chunks (raw observations, per-step rewards, priorities) into one
.npz archive and re-appended on load."""
import numpy as np
import torch


def save_buffer(buf, filename):
    buf._flush()
    n = len(buf)
    meta = dict(n=n, num_steps=buf.num_steps, prioritized=buf._prioritized)
    if n == 0:
        np.savez(filename if str(filename).endswith(".npz") else open(filename, "wb"), **meta)
        return
    from pfrl_b200.replay_buffers.device_buffer import DeviceExperiences

    chunks = {k: [] for k in ("state", "next_state", "action", "step_rewards", "len", "term")}
    for lo in range(0, n, 4096):
        m = min(4096, n - lo)
        idx = torch.arange(lo, lo + m, dtype=torch.int64, device=buf.device)
        out = buf._gather(DeviceExperiences(buf, m, index=idx), 1.0, None, raw=True,
                          want_steps=True)
        chunks["state"].append(out["state"].cpu().numpy())
        chunks["next_state"].append(out["next_state"].cpu().numpy())
        chunks["action"].append(out["action"].cpu().numpy())
        chunks["step_rewards"].append(out["step_rewards"].cpu().numpy())
        chunks["len"].append(out["len"].cpu().numpy())
        chunks["term"].append(out["is_state_terminal"].cpu().numpy())
    arrays = {k: np.concatenate(v) for k, v in chunks.items()}
    if buf._prioritized:
        arrays["priority"] = buf.store.read_priorities()
        arrays["max_priority"] = np.float64(buf.store.info()["max_priority"])
    with open(filename, "wb") as f:
        np.savez(f, **meta, **arrays)


def load_buffer(buf, filename):
    with open(filename, "rb") as f:
        z = np.load(f, allow_pickle=False)
        z = {k: z[k] for k in z.files}
    n = int(z["n"])
    assert int(z["num_steps"]) == buf.num_steps
    if buf.store is not None:
        buf.store.close()
    buf.store = None
    buf._n_total = 0
    buf._pend_parts, buf._pend_exp = [], []
    buf._part_head = 0
    buf._part_cache.clear()
    buf._live_min_seq.clear()
    buf.last_n_transitions.clear()
    if n == 0:
        return
    pr = z.get("priority")
    for k in range(n):
        L = int(z["len"][k])
        s, ns = z["state"][k], z["next_state"][k]
        if buf.store is None:
            buf._create_store(s, z["action"][k])
        s_slots, s_min = buf._parts_of(s)
        n_slots, n_min = buf._parts_of(ns)
        act = buf.layout._action_array(z["action"][k]).tobytes()
        buf._pend_exp.append((s_slots, n_slots, act, list(z["step_rewards"][k][:L]), L,
                              bool(z["term"][k]), None if pr is None else float(pr[k])))
        buf._n_total += 1
        buf._live_min_seq.append(min(s_min, n_min))
        if len(buf._pend_exp) >= 4096:
            buf._flush()
    buf._flush()
