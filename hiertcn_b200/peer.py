"""Catalog-sharded scoring with the exchanges done by our own kernels over NVLink peer memory (csrc/peer.cu) instead of
NCCL collectives.  One process per GPU; torch.distributed is used ONCE, to hand the CUDA IPC handles of the symmetric
buffers around -- no collective runs on the data path afterwards.

``PeerShardedCatalogScorer.score`` is ``dist.ShardedCatalogScorer.score`` (same arithmetic kernels, same merge order, so the
results are bit-identical -- tests/test_gpu_multi.py) with

    all_gather_into_tensor x2            ->  htcn_peer_exchange   (query rows + target ids stored into every peer's buffer)
    all_reduce(target logits)            ->  htcn_peer_bcast_owned (the owner of a target id writes its logit everywhere)
    all_to_all_single x5 + 5 transposes  ->  htcn_peer_exchange   (each shard's partials stored where the owner's merge reads)

each a single launch that publishes its stores with an epoch flag and waits for the peers' flags.  The reference has no
distributed code; this is the B200-native form of the north star's catalog sharding.
"""
from __future__ import annotations

import ctypes as C

from .dist import ShardedCatalogScorer

KIND_GATHER, KIND_TARGET, KIND_PARTIALS = 0, 1, 2
_FLAG_BYTES = 256          # flags[3 kinds][8 ranks] uint32, padded
_LOCAL_BYTES = 256         # done[8] uint32 + err int32 (private to the rank, kept in the same allocation)


def _align(n, a=256):
    return -(-n // a) * a


class _Raw:
    """a device allocation torch did not make, exposed through __cuda_array_interface__"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerBuffer:
    """One symmetric device buffer per rank (same size and layout everywhere), every peer's copy mapped into this process."""

    def __init__(self, dist, rank, world, nbytes, device):
        import torch
        from . import _cabi as cabi
        cabi.load()
        self.cabi, self.rank, self.world, self.nbytes = cabi, rank, world, int(nbytes)
        # every rank walks through the same collectives whatever fails locally, so a failure (no IPC in this container, no
        # peer access between two GPUs) raises on ALL ranks instead of leaving the others inside a collective
        self.own, self.base, err, handle = None, [], None, bytes(64)
        try:
            own = C.c_void_p()
            cabi.call("htcn_peer_alloc", self.nbytes, C.pointer(own))
            self.own = own.value
            hb = (C.c_uint8 * 64)()
            cabi.call("htcn_peer_export", self.own, C.cast(hb, C.c_void_p))
            handle = bytes(hb)
        except Exception as e:          # noqa: BLE001
            err = repr(e)
        got = [None] * world
        dist.all_gather_object(got, (err, handle))
        if err is None and all(g[0] is None for g in got):
            try:
                for r in range(world):
                    if r == rank:
                        self.base.append(self.own)
                        continue
                    p = C.c_void_p()
                    hb = (C.c_uint8 * 64).from_buffer_copy(got[r][1])
                    cabi.call("htcn_peer_import", C.cast(hb, C.c_void_p), C.pointer(p))
                    self.base.append(p.value)
                self.mem = torch.as_tensor(_Raw(self.own, self.nbytes), device=device)     # uint8 view of my own buffer
            except Exception as e:      # noqa: BLE001
                err = repr(e)
        got2 = [None] * world
        dist.all_gather_object(got2, err)                          # doubles as the barrier: everybody has mapped everybody
        bad = [g[0] for g in got if g[0]] + [g for g in got2 if g]
        if bad:
            self.close()
            raise RuntimeError("peer buffer setup failed: " + bad[0])

    def view(self, offset, shape, dtype):
        import torch
        n = int(torch.tensor([], dtype=dtype).element_size())
        for s in shape:
            n *= int(s)
        return self.mem[offset:offset + n].view(dtype).view(*shape)

    def close(self):
        if getattr(self, "own", None) is None:
            return
        for p in self.base:
            if p != self.own:
                self.cabi.call("htcn_peer_unimport", p)
        self.base, self.mem = [], None
        self.cabi.call("htcn_peer_free", self.own)
        self.own = None


class PeerShardedCatalogScorer(ShardedCatalogScorer):
    """``ShardedCatalogScorer`` whose three exchanges are peer-memory kernels.  ``ops`` must be a ``CudaScoreOps``."""

    def __init__(self, ops, dist, rank, world, n_items, n_split=1):
        super().__init__(ops, dist, rank, world, n_items, n_split)
        self.buf = None
        self.cap = None
        self.epoch = 0

    # ---- symmetric buffer
    def _ensure(self, Ql, k, row_bytes):
        import torch
        W, ns = self.world, self.n_split
        want = (Ql, k, row_bytes, ns)
        if self.cap == want:
            return
        if self.buf is not None:
            torch.cuda.synchronize()
            self.dist.barrier()
            self.buf.close()
        Q = W * Ql
        off, o = {}, _FLAG_BYTES + _LOCAL_BYTES
        for name, n in (("h", Q * row_bytes), ("y", Q * 4), ("zy", Q * 4), ("pm", W * ns * Ql * 4), ("ps", W * ns * Ql * 4),
                        ("pc", W * ns * Ql * 4), ("tv", W * Ql * max(k, 1) * 4), ("ti", W * Ql * max(k, 1) * 4)):
            off[name] = o
            o += _align(n)
        self.off = off
        self.buf = PeerBuffer(self.dist, self.rank, W, o, self.ops.m.device)
        self.cap = want
        self.epoch = 0
        b, m = self.buf, self.ops.m
        f32, i32 = torch.float32, torch.int32
        adt = torch.bfloat16 if row_bytes == 256 else torch.float32
        self.h_all = b.view(off["h"], (Q, 128), adt)
        self.y_all = b.view(off["y"], (Q,), i32)
        self.zy = b.view(off["zy"], (Q,), f32)
        self.recv = dict(pm=b.view(off["pm"], (W * ns, Ql), f32), ps=b.view(off["ps"], (W * ns, Ql), f32),
                         pc=b.view(off["pc"], (W * ns, Ql), i32), tv=b.view(off["tv"], (W, Ql, max(k, 1)), f32),
                         ti=b.view(off["ti"], (W, Ql, max(k, 1)), i32))
        self.zy_own = torch.zeros(Q, dtype=f32, device=m.device)
        # the shard's partials live in persistent buffers: the exchange that ships them is then the same launch every call
        # (its argument arrays are built once) and nothing of score() waits for the host
        self.part = dict(pm=torch.empty((ns, Q), dtype=f32, device=m.device), ps=torch.empty((ns, Q), dtype=f32, device=m.device),
                         pc=torch.empty((ns, Q), dtype=i32, device=m.device),
                         tv=torch.empty((1, Q, max(k, 1)), dtype=f32, device=m.device),
                         ti=torch.empty((1, Q, max(k, 1)), dtype=i32, device=m.device),
                         ovf=torch.zeros(1, dtype=i32, device=m.device))
        self.overflowed = torch.zeros(1, dtype=i32, device=m.device)    # sticky: any call since the last check()
        self._xargs = {}
        self._done = b.own + _FLAG_BYTES
        self._err = b.own + _FLAG_BYTES + 64
        self.err_view = b.view(_FLAG_BYTES + 64, (1,), i32)

    def _flags(self, kind):
        b = self.buf
        remote, keep = self.ops.cabi.ptr_array([b.base[p] + (kind * 8 + self.rank) * 4 for p in range(self.world)])
        return remote, keep, b.own + kind * 8 * 4

    def _exchange(self, kind, key, segs_of_peer):
        """segs_of_peer(p) -> [(src_ptr, dst_offset_in_peer_buffer, row_bytes, n_rows, src_pitch, dst_pitch), ...]; the host
        argument arrays are cached under ``key`` (the source pointers and shapes they were built from)"""
        cabi, b, W = self.ops.cabi, self.buf, self.world
        args = self._xargs.get(kind)
        if args is None or args[0] != key:
            src, dst, rb, nr, sp, dp = [], [], [], [], [], []
            n_seg = None
            for p in range(W):
                segs = segs_of_peer(p)
                n_seg = len(segs)
                for s in segs:
                    src.append(s[0]); dst.append(b.base[p] + s[1]); rb.append(s[2]); nr.append(s[3]); sp.append(s[4]); dp.append(s[5])
            arrays = [cabi.ptr_array(src), cabi.ptr_array(dst), cabi.long_array(rb), cabi.int_array(nr), cabi.long_array(sp),
                      cabi.long_array(dp)]
            fr, keep, fl = self._flags(kind)
            args = (key, [a[0] for a in arrays], n_seg, fr, fl, arrays, keep)
            self._xargs[kind] = args
        _, (srcp, dstp, rbp, nrp, spp, dpp), n_seg, fr, fl = args[:5]
        cabi.call("htcn_peer_exchange", srcp, dstp, rbp, nrp, spp, dpp, n_seg, W, self.rank, fr, fl, self._done, self._err,
                  self.epoch, self.ops.m.stream_ptr())

    def check(self):
        """Synchronises.  Raises if a wait of an exchange kernel ran into its spin limit (a peer that never arrived), or if the
        two-pass top-k of any call since the last check overflowed a candidate list (mass ties at a row's threshold): score()
        does not read that flag itself -- it would stall the stream every call -- so such results must be redone with
        ``exact_topk=True`` (the always-exact heap sweep)."""
        if self.buf is None:
            return
        if int(self.err_view.item()):
            raise RuntimeError("peer exchange: a peer did not arrive within the spin limit")
        if int(self.overflowed.item()):
            self.overflowed.zero_()
            raise RuntimeError("sharded top-k: a candidate list overflowed; call score(..., exact_topk=True)")

    def close(self):
        if self.buf is not None:
            self.buf.close()
            self.buf, self.cap = None, None

    # ---- scoring
    def score(self, hout, y_id, k: int = 0, ce: bool = True, rank_metric: bool = True, exact_topk: bool = False):
        """as ShardedCatalogScorer.score; nothing in here waits for the host (see check())"""
        o, W, r = self.ops, self.world, self.rank
        Ql = hout.shape[0]
        need_t = (ce or rank_metric) and y_id is not None
        if W == 1 or not need_t or Ql % 4:
            return super().score(hout, y_id, k, ce, rank_metric)            # top-k only / ragged row counts: the NCCL path
        hout, y_id = hout.contiguous(), y_id.contiguous()
        row_bytes = 128 * hout.element_size()
        self._ensure(Ql, k, row_bytes)
        off, ns, Q = self.off, self.n_split, W * Ql
        self.epoch += 1
        self._mark("start")
        self.zy.zero_()
        # 1. every shard must see every query row: store mine into every rank's h_all / y_all
        self._exchange(KIND_GATHER, (hout.data_ptr(), y_id.data_ptr()),
                       lambda p: [(hout.data_ptr(), off["h"] + r * Ql * row_bytes, Ql * row_bytes, 1, 0, 0),
                                  (y_id.data_ptr(), off["y"] + r * Ql * 4, Ql * 4, 1, 0, 0)])
        self._mark("allgather_queries")
        # 2. target logits: the owner of a target id fills its entry on every rank
        o.target_logit(self.h_all, self.y_all, self.n0, self.n1, self.zy_own)
        self._mark("target_logit")
        dst, keep = o.cabi.ptr_array([self.buf.base[p] + off["zy"] for p in range(W)])
        fr, keep2, fl = self._flags(KIND_TARGET)
        o.cabi.call("htcn_peer_bcast_owned", self.zy_own.data_ptr(), self.y_all.data_ptr(), Q, self.n0, self.n1, dst, W, fr, fl,
                    self._done, self._err, self.epoch, o.m.stream_ptr())
        self._mark("allreduce_target")
        # 3. local sweep (+ exact redo of the rows whose target-referenced partial sum left fp32 range in this shard)
        part = o.sweep(self.h_all, self.y_all, self.zy, self.n0, self.n1, k, ns, ce, rank_metric, out=self.part,
                       sync_overflow=False)
        if k:
            if exact_topk:
                o._heap_topk(self.h_all, self.n0, self.n1, k, ns, part)
            else:
                self.overflowed |= part["ovf"]
        if ce and hasattr(o, "repair_parts"):
            o.repair_parts(self.h_all, self.n0, self.n1, part["pm"], part["ps"])
        self._mark("sweep")
        # 4. the partials of rank p's rows go straight into rank p's merge buffers, slot (my rank, split)
        names = [n for n in ("pm", "ps", "pc") if part.get(n) is not None]

        def segs(p):
            out = [(part[n].data_ptr() + p * Ql * 4, off[n] + r * ns * Ql * 4, Ql * 4, ns, Q * 4, Ql * 4) for n in names]
            if k:
                out += [(part[n].data_ptr() + p * Ql * k * 4, off[n] + r * Ql * k * 4, Ql * k * 4, 1, 0, 0) for n in ("tv", "ti")]
            return out

        self._exchange(KIND_PARTIALS, (tuple(names), k), segs)
        self._mark("alltoall_partials")
        # 5. merge
        rv = self.recv
        zy_local = self.zy[r * Ql:(r + 1) * Ql]
        out = dict(o.finish(rv["pm"] if "pm" in names else None, rv["ps"] if "ps" in names else None,
                            rv["pc"] if "pc" in names else None, y_id, zy_local))
        out["target_logit"] = zy_local.clone()
        if k:
            out.update(o.topk_merge(rv["tv"], rv["ti"], k))
        self._mark("merge")
        return out
