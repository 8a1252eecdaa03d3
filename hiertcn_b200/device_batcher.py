"""Device-side batch assembly: the interaction log and a per-slot session schedule live in HBM and every
batch is assembled by ``htcn_assemble_batch`` -- no host packing, no H2D of ids.

The schedule replays the queue discipline of the reference loader (data_loader.py:170-231: users, in index order, are
appended to the currently shortest of the B slot queues; a user's last session carries the reset mask) once on the host;
``Dataloader_hier_model_xing`` with the same args / seed yields the same sessions in the same slots, so the two can be
compared batch by batch (tests/test_gpu_train.py).  Slots are padded to L = max_activity_len instead of the batch maximum
(results are identical: padded positions are never scored and a causal stack cannot see them).
"""
from __future__ import annotations

import numpy as np

from . import _cabi as cabi


def build_schedule(viewer_table, viewer_data, batch_size, shuffle=False, seed=0, passes=1):
    """-> (items int32 [n_events], sess_off int32 [n_sessions+1], sched_sess int32 [B,P], sched_last uint8 [B,P], lens [B])"""
    items = np.ascontiguousarray(viewer_data[:, 1], dtype=np.int32)
    sess_no = viewer_data[:, 4]
    n_users = viewer_table.shape[0]
    starts, first_sess, n_sess = [], np.zeros(n_users, np.int64), np.zeros(n_users, np.int64)
    for u in range(n_users):
        s0, cnt = int(viewer_table[u, 0]), int(viewer_table[u, 1])
        cut = np.flatnonzero(np.diff(sess_no[s0:s0 + cnt]) != 0) + 1
        first_sess[u] = len(starts)
        starts.extend([s0] + [s0 + int(c) for c in cut])
        n_sess[u] = 1 + len(cut)
    sess_off = np.zeros(len(starts) + 1, np.int32)
    sess_off[:-1] = starts
    # the end of a session is the start of the next one of the same user, or the end of the user's rows
    ends = np.asarray(starts[1:] + [0], dtype=np.int64)
    for u in range(n_users):
        ends[first_sess[u] + n_sess[u] - 1] = int(viewer_table[u, 0]) + int(viewer_table[u, 1])
    # sessions of consecutive users are contiguous in viewer_data, so one offsets array serves: check it
    assert np.all(ends[:-1] == np.asarray(starts[1:])), "viewer_data must hold users back to back"
    sess_off[-1] = ends[-1]
    rng = np.random.default_rng(seed)
    slots = [[] for _ in range(batch_size)]
    last = [[] for _ in range(batch_size)]
    qlen = np.zeros(batch_size)
    for _ in range(passes):
        order = rng.permutation(n_users) if shuffle else np.arange(n_users)
        for u in order:
            q = int(np.argmin(qlen))
            k = int(n_sess[u])
            slots[q].extend(range(int(first_sess[u]), int(first_sess[u]) + k))
            last[q].extend([0] * (k - 1) + [1])
            qlen[q] += k
    lens = np.asarray([len(s) for s in slots], np.int64)
    P = int(lens.max())
    sched = np.zeros((batch_size, P), np.int32)
    flag = np.zeros((batch_size, P), np.uint8)
    for b in range(batch_size):
        sched[b, :lens[b]] = slots[b]
        flag[b, :lens[b]] = last[b]
    return items, sess_off, sched, flag, lens


class DeviceBatcher:
    """``next(state)`` returns the dict ``HierTCN.forward(staged=...)`` / ``HierTCNTrainer.forward_backward(staged=...)``
    consume.  The following batch is assembled on a side stream while the current one is being used; the only value that
    travels to the host is Q, the number of scored positions (a pinned 4-byte read behind an event)."""

    def __init__(self, args, data, device=None, passes=1, type="train"):  # noqa: A002
        import torch
        if not torch.cuda.is_available():
            raise cabi.HtcnError("DeviceBatcher needs a CUDA device; there is no CPU fallback")
        cabi.load()
        table, viewer_data = data
        n = table.shape[0]
        lo, hi = {"train": (0, int(n * 0.8)), "validate": (int(n * 0.8), int(n * 0.9)), "test": (int(n * 0.9), n),
                  "all": (0, n)}[type]
        self.B, self.S, self.L = int(args.batch_size), int(args.max_session_num), int(args.max_activity_len)
        items, sess_off, sched, flag, lens = build_schedule(table[lo:hi], viewer_data, self.B, bool(args.shuffle),
                                                            getattr(args, "seed", 0), passes)
        self.n_batches = int(lens.min()) // self.S
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)  # noqa: E731
        self.items, self.sess_off, self.sched, self.flag = up(items), up(sess_off), up(sched), up(flag)
        self.pitch = int(sched.shape[1])
        self.T = self.S * self.L
        self.slot_off = np.arange(self.S + 1, dtype=np.int32) * self.L
        self.stream = torch.cuda.Stream(device=self.device)
        i32, f32 = torch.int32, torch.float32
        R = self.B * self.T
        self.sets = []
        for _ in range(2):
            self.sets.append(dict(x_id=torch.empty((self.B, self.T), dtype=i32, device=self.device),
                                  y_id=torch.empty((self.B, self.T), dtype=i32, device=self.device),
                                  mask=torch.empty((self.S, self.B), dtype=f32, device=self.device),
                                  row_of=torch.empty(R, dtype=i32, device=self.device),
                                  y_rows=torch.empty(R, dtype=i32, device=self.device),
                                  n_valid=torch.zeros(1, dtype=i32, device=self.device),
                                  n_valid_host=torch.zeros(1, dtype=i32).pin_memory(),
                                  scratch=torch.empty(int(cabi.load().htcn_batcher_scratch_ints(self.B, self.T)), dtype=i32,
                                                      device=self.device),
                                  event=torch.cuda.Event(), used=None))
        self.cursor = 0          # index of the next batch to hand out
        self._launched = -1
        self.done = False
        self._launch(0)

    def _launch(self, k):
        import torch
        if k >= self.n_batches or k <= self._launched:
            return
        s = self.sets[k & 1]
        with torch.cuda.stream(self.stream):
            if s["used"] is not None:
                self.stream.wait_event(s["used"])              # the consumer of batch k-2 is done with these buffers
            cabi.call("htcn_assemble_batch", self.items.data_ptr(), self.sess_off.data_ptr(), self.sched.data_ptr(),
                      self.flag.data_ptr(), self.pitch, k * self.S, self.B, self.S, self.L, s["x_id"].data_ptr(),
                      s["y_id"].data_ptr(), s["mask"].data_ptr(), s["row_of"].data_ptr(), s["y_rows"].data_ptr(),
                      s["n_valid"].data_ptr(), s["scratch"].data_ptr(), self.stream.cuda_stream)
            s["n_valid_host"].copy_(s["n_valid"], non_blocking=True)
            s["event"].record(self.stream)
        self._launched = k

    def next(self, state=None):
        import torch
        if self.cursor >= self.n_batches:
            raise StopIteration("schedule exhausted: build the DeviceBatcher with more passes")
        k = self.cursor
        s = self.sets[k & 1]
        s["event"].synchronize()                               # Q of batch k is on the host
        Q = int(s["n_valid_host"][0])
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(s["event"])
        if state is None:
            state = torch.zeros((self.B, 256), dtype=torch.float32, device=self.device)
        elif not hasattr(state, "data_ptr"):
            state = torch.from_numpy(np.ascontiguousarray(state, dtype=np.float32)).to(self.device)
        self.cursor += 1
        self.done = self.cursor >= self.n_batches
        # everything enqueued so far on the consumer's stream -- in particular the work on batch k-1, whose buffers batch
        # k+1 reuses -- precedes this event; the side stream waits for it before assembling k+1
        ev = torch.cuda.Event()
        ev.record(cur)
        self.sets[(k + 1) & 1]["used"] = ev
        self._launch(k + 1)
        return dict(x_id=s["x_id"], y_id=s["y_id"], mask=s["mask"], row_of=s["row_of"], y_rows=s["y_rows"][:max(Q, 0)],
                    state=state, B=self.B, T=self.T, S=self.S, Q=Q, slot_off=self.slot_off, h2d_bytes=0)
