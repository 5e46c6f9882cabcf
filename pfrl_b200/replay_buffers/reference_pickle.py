"""Read a replay checkpoint written by the reference.

``ReplayBuffer.save`` pickles ``self.memory`` (pfrl/replay_buffers/
replay_buffer.py:85-94): a ``RandomAccessQueue`` (two Python lists,
collections/random_access_queue.py:13-18) for the uniform buffer, a whole
``PrioritizedBuffer`` (deque of experiences + the two nested-list trees +
``max_priority``, collections/prioritized.py:21-37) for the prioritised one;
releases before v0.2 pickled a bare ``collections.deque``.  ``DQN.save_snapshot``
stores it as ``<dir>/replay_buffer.pkl`` (agents/dqn.py:794-810).

The file references classes of the ``pfrl`` package.  This reader does not
need that package: every ``pfrl.*`` class is mapped onto an attribute bag and
the content is pulled out of it structurally -- experiences oldest first,
leaf priorities by walking the sum tree, ``max_priority``.  Object identity
inside the pickle (LazyFrames frames shared by consecutive observations, the
observation shared by ``next_state`` of step t and ``state`` of step t+1)
survives unpickling, which is what the device buffers' de-duplication keys on.
"""
import collections
import itertools
import pickle

import numpy as np

from pfrl_b200.utils.lazy_frames import LazyFrames


class _Bag(object):
    """Stand-in for an instance of a reference class: just its ``__dict__``."""


_BAGS = {}


def _bag_class(module, name):
    key = (module, name)
    if key not in _BAGS:
        _BAGS[key] = type(name, (_Bag,), {"_ref_module": module})
    return _BAGS[key]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "pfrl" or module.startswith("pfrl."):
            if name == "LazyFrames":
                return LazyFrames
            return _bag_class(module, name)
        return super().find_class(module, name)


def _tree_leaves(tree):
    """Leaf values 0 .. length-1 of a pickled TreeQueue (nodes are
    ``[left, right, value]`` lists, an absent subtree is ``[]``; the top node
    covers ``bounds``; collections/prioritized.py:140-242)."""
    n = tree.length
    out = np.zeros(n, dtype=np.float64)
    if n == 0:
        return out
    seen = 0
    todo = [(tree.bounds[0], tree.bounds[1], tree.root)]
    while todo:
        lo, hi, node = todo.pop()
        if not node or hi <= 0 or lo >= n:
            continue
        if hi - lo == 1:
            out[lo] = node[2]
            seen += 1
            continue
        mid = (lo + hi) // 2
        todo.append((lo, mid, node[0]))
        todo.append((mid, hi, node[1]))
    if seen != n:
        raise ValueError("priority tree holds %d leaves, expected %d" % (seen, n))
    return out


ReferenceReplay = collections.namedtuple(
    "ReferenceReplay", "experiences priorities max_priority capacity")


def read(filename):
    """-> ReferenceReplay(experiences (list, oldest first; each a list of 1..n
    transition dicts), priorities (float64 array or None), max_priority
    (float or None), capacity (int or None))."""
    with open(filename, "rb") as f:
        memory = _Unpickler(f).load()
    if isinstance(memory, (collections.deque, list)):
        return ReferenceReplay(list(memory), None, None, getattr(memory, "maxlen", None))
    kind = type(memory).__name__
    if kind == "RandomAccessQueue":
        items = list(itertools.chain(reversed(memory._queue_front), memory._queue_back))
        return ReferenceReplay(items, None, None, memory.maxlen)
    if kind == "PrioritizedBuffer":
        if memory.flag_wait_priority:
            raise ValueError(
                "the checkpoint was written between sample() and update_errors(): the sampled "
                "leaves hold 0.0 and their priorities are lost")
        pri = _tree_leaves(memory.priority_sums)
        items = list(memory.data)
        if len(items) != len(pri):
            raise ValueError("%d experiences but %d priorities" % (len(items), len(pri)))
        return ReferenceReplay(items, pri, float(memory.max_priority), memory.capacity)
    raise TypeError("unrecognised replay checkpoint content: %s" % kind)


def looks_like_pickle(filename):
    with open(filename, "rb") as f:
        head = f.read(2)
    return len(head) == 2 and head[0] == 0x80 and 2 <= head[1] <= 5
