"""Replay the scripted traces of tests/golden/*.npz (made by
oracle/gen_golden.py from the real reference) against any implementation of
the replay-buffer API and compare with what the reference produced."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OP_APPEND, OP_STOP, OP_SAMPLE = 0, 1, 2


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def make_state(sid, shape):
    n = int(np.prod(shape))
    return ((np.arange(n, dtype=np.int64) * 31 + sid * 17 + (sid * sid) % 251) % 256).astype(
        np.uint8).reshape(shape)


def per_kwargs(g):
    seed, capacity, num_steps, n_envs, batch, lazy = [int(x) for x in g["meta"]]
    alpha, beta0, betasteps, gamma = [float(x) for x in g["params"]]
    nb = str(g["normalize_by_max"])
    nb = {"True": True, "False": False}.get(nb, nb)
    return dict(capacity=capacity, alpha=alpha, beta0=beta0,
                betasteps=None if np.isnan(betasteps) else betasteps, normalize_by_max=nb,
                num_steps=num_steps), seed, gamma, bool(lazy)


def replay_per_trace(g, make_buffer, batch_fn, lazy_cls, indices_of=None, rtol=2e-6):
    """make_buffer(**kw) -> buffer; batch_fn(exps, gamma) -> dict of numpy
    arrays (state/next_state already float32 scaled by 1/255);
    indices_of(buffer, exps) -> sampled logical indices (or None to skip)."""
    kw, seed, gamma, lazy = per_kwargs(g)
    obs_shape = tuple(int(x) for x in g["obs_shape"])
    frame_shape = (1,) + obs_shape[1:]
    k = obs_shape[0]
    buf = make_buffer(**kw)
    np.random.seed(seed)
    frames, cur = {}, {}

    def obs_of(e, sid, reset=False):
        # mirrors oracle/gen_golden.py: next_state of step t is the very same
        # object as the state of step t+1; frames are shared between stacks
        if not lazy:
            return make_state(sid, obs_shape)
        f = make_state(sid, frame_shape)
        frames[e] = [f] * k if reset else frames[e][1:] + [f]
        return lazy_cls(list(frames[e]), stack_axis=0)

    ai = si = 0
    off = 0
    n_samples = 0
    errs = g["errors"]
    for row in g["ops"]:
        op, e, sid, nsid, action, terminal, n = [int(x) for x in row]
        if op == OP_APPEND:
            if cur.get(e) is None or cur[e][0] != sid:
                cur[e] = (sid, obs_of(e, sid, reset=True))
            s = cur[e][1]
            ns = obs_of(e, nsid)
            buf.append(s, action, float(g["rewards"][ai]), ns, None, bool(terminal), env_id=e)
            cur[e] = (nsid, ns)
            ai += 1
        elif op == OP_STOP:
            buf.stop_current_episode(env_id=e)
            cur[e] = None
            assert len(buf) == int(g["length"][ai - 1])
        else:
            exps = buf.sample(n)
            sl = slice(off, off + n)
            if indices_of is not None:
                got = np.asarray(indices_of(buf, exps))
                assert np.array_equal(got, g["idx"][sl]), "sampled indices differ at sample %d" % si
            b = batch_fn(exps, gamma)
            np.testing.assert_allclose(b["weights"], g["weight"][sl].astype(np.float32), rtol=rtol)
            np.testing.assert_allclose(b["reward"], g["reward"][sl], rtol=1e-6, atol=1e-7)
            assert np.array_equal(b["discount"], g["discount"][sl])
            assert np.array_equal(b["is_state_terminal"], g["terminal"][sl])
            assert np.array_equal(b["action"], g["action"][sl])
            ssum = b["state"].reshape(n, -1).astype(np.float64).sum(1)
            nsum = b["next_state"].reshape(n, -1).astype(np.float64).sum(1)
            np.testing.assert_allclose(ssum, g["state_sum"][sl], rtol=1e-6)
            np.testing.assert_allclose(nsum, g["next_sum"][sl], rtol=1e-6)
            buf.update_errors([float(x) for x in errs[sl]])
            off += n
            si += 1
            n_samples += 1
    assert n_samples == len(g["sample_sizes"])
    return buf
