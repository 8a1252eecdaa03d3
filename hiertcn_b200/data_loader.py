"""Batch layout of the hot path (host side, numpy).

Mirrors the tensor contract of the reference's hierarchical loader
(``Dataloader_hier_model_xing.get_batch`` -> ``dequeue``, reference data_loader.py:233-272):

    get_batch() -> (x_list, y_list, mask_list, info_list), each an S-list
      x_list[s]    [B, L_s]    = [0, a_1 .. a_{L_s-1}]   (session shifted right, null id first)
      y_list[s]    [B, L_s]    = [a_1 .. a_{L_s}]        (zero right-padded; ids 1..N-1, 0 = null)
      mask_list[s] [B, 1]      0 iff this was the user's last session (state reset), else 1
      info_list[s] [B, L_s, 5] = [user_id, item_id, type, timestamp, session_no]
      L_s = max over the batch of the session length in slot s      (float64 arrays, like the reference)

``pack_batch`` turns that S-list into the packed device layout the C ABI takes
(x_id / y_id ``[B, T]`` int32 with ``T = sum L_s`` and ``slot_off[S+1]``), which is exactly the
``y_id = concat(y_list, axis=1)`` the reference feeds (run_hier_xing.py:278).

CSV parsing (reference data_prepare.py) is out of scope; ``make_synthetic_interactions`` produces
an interaction table of the same shape (``viewer_table`` / ``viewer_data``) for the queue loader.
"""
from __future__ import annotations

from collections import deque

import numpy as np


# --------------------------------------------------------------------------------------
# synthetic XING-shaped ids
# --------------------------------------------------------------------------------------


class ItemSampler:
    """Zipf(alpha) over ids 1..N-1 (heavy-tailed like XING after the >=50-interaction filter,
    reference data_prepare.py:32) or uniform (worst-case gather locality)."""

    def __init__(self, item_num: int, dist: str = "zipf", alpha: float = 1.05):
        self.n = int(item_num)
        self.dist = dist
        if dist == "zipf":
            p = np.arange(1, self.n, dtype=np.float64) ** (-alpha)
            self.cdf = np.cumsum(p / p.sum())
            self.cdf[-1] = 1.0
        elif dist != "uniform":
            raise ValueError(dist)

    def sample(self, rng: np.random.Generator, size) -> np.ndarray:
        if self.dist == "uniform":
            return rng.integers(1, self.n, size=size, dtype=np.int64)
        # rank r (0-based, most popular first) -> a fixed pseudo-random id so popular rows are scattered
        r = np.searchsorted(self.cdf, rng.random(size=size), side="left").astype(np.int64)
        return (r * 2654435761 % (self.n - 1)) + 1


def synthetic_batch(B: int, S: int = 10, L: int = 20, item_num: int = 20778, seed: int = 0,
                    lengths: str = "dense", id_dist: str = "zipf", mask_keep: float = 0.9,
                    sampler: ItemSampler | None = None):
    """One batch in the ``dequeue()`` layout, generated directly (no queues).

    lengths: 'dense'  -- every session has exactly L events (defines the algorithmic-work denominator)
             'ragged' -- geometric, mean ~6, clipped to [1, L], padded to the per-slot max
    Returns (x_list, y_list, mask_list) with float64 arrays like the reference loader.
    """
    rng = np.random.default_rng(seed)
    sampler = sampler or ItemSampler(item_num, id_dist)
    x_list, y_list, mask_list = [], [], []
    for _ in range(S):
        if lengths == "dense":
            n = np.full(B, L, dtype=np.int64)
        elif lengths == "ragged":
            n = np.clip(rng.geometric(1.0 / 6.0, size=B), 1, L).astype(np.int64)
        else:
            raise ValueError(lengths)
        Ls = int(n.max())
        items = sampler.sample(rng, (B, Ls))
        valid = np.arange(Ls)[None, :] < n[:, None]
        padded = np.zeros((B, Ls + 1), dtype=np.float64)            # [0, a_1..a_n, 0..]
        padded[:, 1:] = np.where(valid, items, 0)
        x_list.append(padded[:, :-1].copy())
        y_list.append(padded[:, 1:].copy())
        mask_list.append((rng.random((B, 1)) < mask_keep).astype(np.float64))
    return x_list, y_list, mask_list


def pack_batch(x_list, y_list, mask_list):
    """S-lists -> packed arrays for the device path.

    Returns dict(x_id [B,T] int32, y_id [B,T] int32, slot_off [S+1] int32, mask [S,B] float32)."""
    S = len(x_list)
    lens = [int(np.shape(x)[1]) for x in x_list]
    slot_off = np.zeros(S + 1, dtype=np.int32)
    slot_off[1:] = np.cumsum(lens)
    x_id = np.ascontiguousarray(np.concatenate([np.asarray(x) for x in x_list], axis=1).astype(np.int32))
    y_id = np.ascontiguousarray(np.concatenate([np.asarray(y) for y in y_list], axis=1).astype(np.int32))
    mask = np.ascontiguousarray(np.stack([np.asarray(m).reshape(-1) for m in mask_list]).astype(np.float32))
    return dict(x_id=x_id, y_id=y_id, slot_off=slot_off, mask=mask)


# --------------------------------------------------------------------------------------
# queue loader with the reference's semantics (enqueue / enqueue_loop / dequeue)
# --------------------------------------------------------------------------------------


def make_synthetic_interactions(num_users: int, item_num: int, seed: int = 0, id_dist: str = "zipf",
                                max_sessions: int = 12, mean_session_len: float = 6.0):
    """An interaction table shaped like ``load_xing``'s output (reference data_prepare.py:15-93):
    viewer_data [n_events, 5] = [user_id, item_id, type, timestamp, session_no], sorted by user then
    time; viewer_table [n_users, 2] = [first row, row count]."""
    rng = np.random.default_rng(seed)
    sampler = ItemSampler(item_num, id_dist)
    rows, table, start = [], [], 0
    for u in range(num_users):
        n_sess = int(rng.integers(2, max_sessions + 1))
        t = float(rng.integers(0, 10 ** 6))
        cnt = 0
        for s in range(n_sess):
            n = int(np.clip(rng.geometric(1.0 / mean_session_len), 1, 40))
            items = sampler.sample(rng, n)
            for it in items:
                t += float(rng.integers(1, 600))
                rows.append((u, int(it), int(rng.integers(1, 4)), t, s))
            t += 1800.0 + float(rng.integers(0, 86400))            # > 30 min gap = new session
            cnt += n
        table.append((start, cnt))
        start += cnt
    return np.asarray(table, dtype=np.float64), np.asarray(rows, dtype=np.float64)


class Dataloader_hier_model_xing:
    """Per-slot user queues of sessions; each ``get_batch`` dequeues ``max_session_num`` sessions per
    slot.  Same behaviour as reference data_loader.py:115-272 (users are appended to the shortest
    queue; a user's last session carries mask 0; sessions are clipped to ``max_activity_len``;
    ``done`` reports that the user index wrapped during this call), re-implemented with deques."""

    def __init__(self, args, type="train", data=None):  # noqa: A002  (reference spelling)
        if data is None:
            raise ValueError("pass data=(viewer_table, viewer_data); CSV parsing is out of scope")
        self.args = args
        self.done = False
        table, self.viewer_data = data
        n = table.shape[0]
        lo, hi = {"train": (0, int(n * 0.8)), "validate": (int(n * 0.8), int(n * 0.9)),
                  "test": (int(n * 0.9), n)}[type]                  # reference :141-146
        self.viewer_table = table[lo:hi]
        self._rng = np.random.default_rng(getattr(args, "seed", 0))
        self.refresh()
        B = args.batch_size
        self.data = [deque() for _ in range(B)]
        self.info = [deque() for _ in range(B)]
        self.mask = [deque() for _ in range(B)]
        self.queue_len = np.zeros(B)

    def refresh(self, i=None):
        n = self.viewer_table.shape[0]
        self.index = self._rng.permutation(n) if self.args.shuffle else np.arange(n)
        self.index_pointer = 0

    def enqueue(self):
        row = self.index[self.index_pointer]
        start = int(self.viewer_table[row, 0])
        end = start + int(self.viewer_table[row, 1])
        ev = self.viewer_data[start:end]
        # session boundaries: wherever session_no changes between consecutive events
        cut = np.flatnonzero(np.diff(ev[:, 4]) != 0) + 1
        bounds = np.concatenate([[0], cut, [len(ev)]])
        cap = self.args.max_activity_len
        sess_items = [ev[a:b, 1][:cap].astype(np.int64) for a, b in zip(bounds[:-1], bounds[1:])]
        sess_info = [ev[a:b][:cap] for a, b in zip(bounds[:-1], bounds[1:])]
        q = int(np.argmin(self.queue_len))
        self.data[q].extend(sess_items)
        self.info[q].extend(sess_info)
        self.mask[q].extend([1] * (len(sess_items) - 1) + [0])
        self.queue_len[q] += len(sess_items)
        self.index_pointer += 1
        if self.index_pointer >= self.index.shape[0]:
            self.refresh()
            return True
        return False

    def enqueue_loop(self):
        stop = False
        while True:
            stop = self.enqueue() or stop
            if np.amin(self.queue_len) > self.args.max_session_num:
                return stop

    def dequeue(self):
        B = self.args.batch_size
        x_out, y_out, info_out, mask_out = [], [], [], []
        for _ in range(self.args.max_session_num):
            heads = [self.data[b].popleft() for b in range(B)]
            infos = [self.info[b].popleft() for b in range(B)]
            flags = [self.mask[b].popleft() for b in range(B)]
            width = max(len(h) for h in heads) + 1
            padded = np.zeros((B, width))
            info = np.zeros((B, width, 5))
            for b in range(B):
                padded[b, 1:len(heads[b]) + 1] = heads[b]
                info[b, 1:len(heads[b]) + 1, :] = infos[b]
            self.queue_len -= 1
            x_out.append(padded[:, :-1])
            y_out.append(padded[:, 1:])
            info_out.append(info[:, 1:, :])
            mask_out.append(np.asarray(flags, dtype=np.float64).reshape(B, 1))
        return x_out, y_out, mask_out, info_out

    def get_batch(self):
        self.done = self.enqueue_loop()
        return self.dequeue()

    # cursor of the loader (not saved by the reference, run_hier_xing.py:317-321: a resumed run restarts the epoch)
    def state_dict(self):
        return dict(index=[int(i) for i in self.index], index_pointer=int(self.index_pointer),
                    rng=self._rng.bit_generator.state,
                    data=[[[int(v) for v in s] for s in q] for q in self.data],
                    info=[[np.asarray(s).tolist() for s in q] for q in self.info],
                    mask=[[int(v) for v in q] for q in self.mask], queue_len=[float(v) for v in self.queue_len])

    def load_state_dict(self, d):
        self.index = np.asarray(d["index"], dtype=np.int64)
        self.index_pointer = int(d["index_pointer"])
        self._rng.bit_generator.state = d["rng"]
        self.data = [deque(np.asarray(s, dtype=np.int64) for s in q) for q in d["data"]]
        self.info = [deque(np.asarray(s, dtype=np.float64).reshape(-1, 5) for s in q) for q in d["info"]]
        self.mask = [deque(int(v) for v in q) for q in d["mask"]]
        self.queue_len = np.asarray(d["queue_len"], dtype=np.float64)
