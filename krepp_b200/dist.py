"""Multi-GPU plumbing for the read-sharded mode (mode A): reads are independent units, the index image is
replicated per GPU, every rank takes one contiguous slice of the read stream and there is NO collective on the data
path.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is only used for the barrier, for max-over-ranks timing and
for gathering per-rank counters."""
from __future__ import annotations

import os


def env_rank_world() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of n_items for `rank`; slices of all ranks tile [0, n_items) in order."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (e.g. elapsed seconds); identity when not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(values, device=None):
    """Element-wise sum of a small list of per-rank counters (e.g. reads, records, algorithmic bytes)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(v) for v in values]
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


def gather_text(text: str) -> list[str]:
    """Rank-ordered list of every rank's text block (output concatenation in read order); identity when single."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [text]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, text)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Mode B: the index is larger than one GPU's HBM, so the table is split by LSH bucket (row) range over the
# ranks and the data path has two real exchange steps -- lookups to the shard that owns their bucket, hit entries back to
# the read's home rank.  The kernels and the cut points live in the library (krepp_shard_lookup / _join / _finish,
# include/krepp_b200.h); this file only moves the bytes: torch.distributed all_to_all_single (NCCL over NVLink on GPUs),
# or, for several logical ranks inside one process (tests on a single GPU), plain tensor hand-over.

TUPLE_WORDS = 4  # a lookup tuple and a hit entry are both 4 x u32


def exchange_v(send, send_counts, group=None, recv_counts=None):
    """All-to-all-v of the rows of `send` (a [n, w] tensor, rank g's rows contiguous, send_counts[g] of them).
    Returns (recv, recv_counts): the rows every rank sent here, in rank order.  recv_counts, when the caller knows them (the
    row-boundary slices have the same sizes every batch), saves the exchange of the counts and its host round trip."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if recv_counts is None:
        sc = torch.as_tensor([int(c) for c in send_counts], dtype=torch.int64, device=send.device)
        rc = torch.empty(world, dtype=torch.int64, device=send.device)
        dist.all_to_all_single(rc, sc, group=group)
        recv_counts = [int(x) for x in rc.tolist()]
    recv = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send[:sum(int(c) for c in send_counts)], output_split_sizes=recv_counts,
                           input_split_sizes=[int(c) for c in send_counts], group=group)
    return recv, recv_counts


class _Lane:
    """One batch in flight on a rank: a batch slot (its own CUDA stream and result buffers) and the exchange buffers."""

    def __init__(self, rank: "ShardRank", max_reads: int, max_bases: int, batch_kw: dict):
        import numpy as np
        import torch
        from .capi import IBatch
        info = rank.index.info
        self.slot = IBatch(rank.index, np.zeros((0, 1), np.uint8), capacity=(max_reads, max_bases), **batch_kw)
        self.row_begin = torch.empty(rank.nrows + 1, dtype=torch.int32, device=rank.dev)
        # 2 strands x one window per base, of which (r+1)/m are eligible on average
        frac = (info.r + 1) / info.m if info.frac else 1.0 / info.m
        self.tuples = torch.empty((int(max_bases * 2 * frac * 1.1) + 4096, TUPLE_WORDS), dtype=torch.int32, device=rank.dev)
        self.hits = torch.empty((max(96 * max_reads, 65536), TUPLE_WORDS), dtype=torch.int32, device=rank.dev)
        self.keep = None


class ShardRank:
    """One rank of a mode-B job: its shard of the index on its GPU and `lanes` batches in flight (each a batch slot plus the
    device buffers of the two exchanges).  The three phases below are the library's; `ShardedJob` strings them together with
    the exchanges.  With two lanes the tail of batch i (regrouping, resolve, solve: enqueued, not waited for) runs while the
    host drives batch i + 1 through its lookups and exchanges."""

    def __init__(self, index_dir: str, device: int, rank: int, world: int, max_reads: int, max_bases: int, lanes: int = 1, **batch_kw):
        import torch
        from .capi import Index
        self.torch, self.rank, self.world = torch, rank, world
        self.dev = torch.device("cuda", device)
        with torch.cuda.device(self.dev):
            self.index = Index(index_dir, device, shard=rank, nshards=world)
            self.nrows = int(self.index.info.nrows)
            self.splits = [int(x) for x in self.index.row_splits]
            self.lanes = [_Lane(self, max_reads, max_bases, batch_kw) for _ in range(max(1, lanes))]
        self.slot = self.lanes[0].slot

    # phase 1 (home): reads -> tuples grouped by row; what goes to every owner
    def lookup(self, d_bases, d_offsets, n_reads: int, lane: int = 0):
        from .capi import CapacityError
        torch, ln = self.torch, self.lanes[lane]
        with torch.cuda.device(self.dev):
            torch.cuda.current_stream().synchronize()  # the reads may still be on their way (torch's stream); the library has its own
            while True:
                try:
                    so = ln.slot.shard_lookup(d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, d_bases.numel(), ln.tuples.data_ptr(),
                                              ln.tuples.shape[0], ln.row_begin.data_ptr())
                    break
                except CapacityError as e:
                    ln.tuples = torch.empty((e.demand + e.demand // 10 + 4096, TUPLE_WORDS), dtype=torch.int32, device=self.dev)
            counts = [int(so[g + 1] - so[g]) for g in range(self.world)]
            # owner g needs row_begin[splits[g] .. splits[g+1]] inclusive: neighbouring slices share one word, so they are
            # laid out one after the other for the all-to-all
            rb = torch.cat([ln.row_begin[self.splits[g]:self.splits[g + 1] + 1] for g in range(self.world)])
            rb_counts = [self.splits[g + 1] - self.splits[g] + 1 for g in range(self.world)]
        return ln.tuples, counts, rb, rb_counts

    # phase 2 (owner): every sender's tuples against this shard -> hit entries, sender by sender
    def join(self, recv_tuples, recv_counts, recv_rb, lane: int = 0):
        from .capi import CapacityError
        torch, ln = self.torch, self.lanes[lane]
        nloc = self.splits[self.rank + 1] - self.splits[self.rank] + 1
        tp, rp, at = [], [], 0
        for s in range(self.world):
            tp.append(recv_tuples.data_ptr() + 16 * at)  # a tuple is 16 bytes
            rp.append(recv_rb.data_ptr() + 4 * nloc * s)
            at += recv_counts[s]
        with torch.cuda.device(self.dev):
            torch.cuda.current_stream().synchronize()  # the exchange ran on torch's stream, the library has its own
            while True:
                try:
                    ho = ln.slot.shard_join(tp, rp, ln.hits.data_ptr(), ln.hits.shape[0])
                    break
                except CapacityError as e:
                    ln.hits = torch.empty((e.demand + e.demand // 8 + 4096, TUPLE_WORDS), dtype=torch.int32, device=self.dev)
        return ln.hits, [int(ho[s + 1] - ho[s]) for s in range(self.world)]

    # phase 3 (home): the batch's hit entries from all owners -> records (enqueued; results() waits)
    def finish(self, recv_hits, lane: int = 0):
        self.lanes[lane].keep = recv_hits  # must stay alive until the wait
        with self.torch.cuda.device(self.dev):
            self.torch.cuda.current_stream().synchronize()
            self.lanes[lane].slot.shard_finish(recv_hits.data_ptr(), recv_hits.shape[0])

    def results(self, rows: bool = True, lane: int = 0) -> dict:
        ln = self.lanes[lane]
        with self.torch.cuda.device(self.dev):
            r = ln.slot.wait() if rows else ln.slot.wait_device()
        ln.keep = None
        return r

    def close(self):
        for ln in self.lanes:
            ln.slot.close()
        self.index.close()


class ShardedJob:
    """Runs batches through mode B.  `ranks` is the list of ShardRank objects living in THIS process: one (its peers are
    other processes, exchanges go through torch.distributed) or all of them (logical ranks on one GPU, exchanges are
    tensor hand-overs)."""

    def __init__(self, ranks, group=None):
        self.ranks, self.group = ranks, group
        self.local = len(ranks) > 1 or ranks[0].world == 1
        if self.local:
            assert [r.rank for r in ranks] == list(range(ranks[0].world)), "an in-process job holds every rank"
        self.bytes_exchanged = 0

    def _exchange(self, sends, recv_counts=None):
        """sends[i] = (tensor, counts) of in-process rank i -> [(recv tensor, recv counts)] per in-process rank."""
        import torch
        if not self.local:
            t, c = sends[0]
            recv, rc = exchange_v(t, c, self.group, recv_counts)
            self.bytes_exchanged += recv.numel() * recv.element_size()
            return [(recv, rc)]
        out = []
        for j in range(len(self.ranks)):
            parts, rc = [], []
            for t, c in sends:
                o = sum(c[:j])
                parts.append(t[o:o + c[j]])
                rc.append(c[j])
            recv = torch.cat(parts)
            self.bytes_exchanged += recv.numel() * recv.element_size()
            out.append((recv, rc))
        return out

    def start(self, batches, lane: int = 0):
        """Phases 1-3 of one batch on `lane`, up to the enqueue of its last phase (no wait)."""
        p1 = [r.lookup(*b, lane=lane) for r, b in zip(self.ranks, batches)]
        tup = self._exchange([(t, c) for t, c, _, _ in p1])
        r0 = self.ranks[0]  # every sender sends this rank its own slice of row boundaries: the sizes are known
        rbs = self._exchange([(rb.view(-1, 1), rc) for _, _, rb, rc in p1], None if self.local else [r0.splits[r0.rank + 1] - r0.splits[r0.rank] + 1] * r0.world)
        p2 = [r.join(t, c, rb.view(-1), lane=lane) for r, (t, c), (rb, _) in zip(self.ranks, tup, rbs)]
        back = self._exchange(p2)
        for r, (h, _) in zip(self.ranks, back):
            r.finish(h, lane=lane)

    def run(self, batches, rows: bool = True):
        """batches[i] = (d_bases uint8 tensor, d_offsets int64 tensor, n_reads) of in-process rank i.  Returns the result
        dicts (krepp_batch_wait; krepp_batch_wait_device when rows is False) in the same order."""
        self.start(batches)
        return [r.results(rows) for r in self.ranks]

    def run_stream(self, stream, rows: bool = True, consume=None):
        """A sequence of batches (each as for run()) through the ranks' lanes: batch i + 1 is started while batch i's last
        phase still runs, and a batch is waited for only when its lane is needed again.  Every rank of the job must pass the
        same number of batches (the exchanges are collectives).  consume(i, results, lane) is called for every batch in order,
        while its result buffers are still valid (they are reused by the batch that takes the lane next); returns the number of batches."""
        nl = min(len(r.lanes) for r in self.ranks)
        pending = []  # (batch number, lane)
        n = 0
        for batches in stream:
            lane = n % nl
            if len(pending) == nl:
                i, l = pending.pop(0)
                res = [r.results(rows, lane=l) for r in self.ranks]
                if consume:
                    consume(i, res, l)
            self.start(batches, lane=lane)
            pending.append((n, lane))
            n += 1
        for i, l in pending:
            res = [r.results(rows, lane=l) for r in self.ranks]
            if consume:
                consume(i, res, l)
        return n
