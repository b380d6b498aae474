"""Multi-GPU plumbing of the path, one process per GPU:

  level 1  across merges -- sibling sub-trees of the guide tree are independent (merge_graphs only reads its two children):
           `TreeSchedule` is a ready-queue over the guide tree that replaces the reference's serial post-order walk
           (packages/pangraph/src/tree/clade.rs:49-71, driven from commands/build/build_run.rs:111); every ready merge is
           handed to the least-loaded rank.  `rounds_of_rank` is its degenerate form for a flat list of leaf merges.
  level 2  inside one round -- mm_map is a pure function of (index, query, name, options)
           (packages/pangraph/src/align/minimap2_lib/align_with_minimap2_lib.rs:64-74 is the loop that is sharded):
           every rank builds the (deterministic) index of the round, maps its share of the queries (`shard_queries`:
           longest first onto the least-loaded rank) and the hit records go back to rank 0 in query-index order.

The only collective on the path is the gather of variable-length match records (mm_reg1_t 80 B + CIGAR words) to rank 0:
`gather_rounds` exchanges the byte counts (one all_gather of an int64) and then moves exactly those bytes with point-to-point
sends (NCCL over NVLink on GPUs; gloo in the CPU tests) -- no padding to the largest rank."""
import heapq
import struct

import torch
import torch.distributed as dist


def rounds_of_rank(n_rounds, rank, world):
    """Round r belongs to rank r % world (pair index modulo the number of GPUs)."""
    return list(range(rank, n_rounds, world))


def pack_rounds(payloads):
    """[(round_index, bytes)] -> one buffer: u64 count, then per round u64 index, u64 size, payload."""
    parts = [struct.pack("<Q", len(payloads))]
    for idx, blob in payloads:
        parts.append(struct.pack("<QQ", idx, len(blob)))
        parts.append(blob)
    return b"".join(parts)


def unpack_rounds(buf):
    (n,) = struct.unpack_from("<Q", buf, 0)
    off, out = 8, []
    for _ in range(n):
        idx, size = struct.unpack_from("<QQ", buf, off)
        off += 16
        out.append((idx, bytes(buf[off:off + size])))
        off += size
    return out


def gather_rounds(payloads, device, group=None):
    """Every rank contributes [(index, bytes)]; rank 0 gets all of them sorted by index, the others get None.
    Sizes first (all_gather of one int64 per rank), then every rank > 0 sends exactly its bytes to rank 0."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    blob = pack_rounds(payloads)
    size = torch.tensor([len(blob)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    if rank != 0:
        buf = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
        dist.send(buf, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return None
    bufs = [None] + [torch.empty(sizes[r], dtype=torch.uint8, device=device) for r in range(1, world)]
    reqs = [dist.irecv(bufs[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group) for r in range(1, world)]
    rounds = unpack_rounds(blob)
    for r, q in enumerate(reqs, start=1):
        q.wait()
        rounds.extend(unpack_rounds(bufs[r].cpu().numpy().tobytes()))
    rounds.sort(key=lambda t: t[0])
    return rounds


# ---------------- level 2: the queries of one round across ranks ----------------

def shard_queries(lens, world):
    """Longest-processing-time dealing: queries by decreasing length (ties: lower index first) onto the rank with the
    least bases so far (ties: lower rank).  Deterministic on every rank.  Returns world lists of query indices, ascending."""
    load = [(0, r) for r in range(world)]
    heapq.heapify(load)
    shards = [[] for _ in range(world)]
    for i in sorted(range(len(lens)), key=lambda i: (-lens[i], i)):
        bases, r = heapq.heappop(load)
        shards[r].append(i)
        heapq.heappush(load, (bases + lens[i], r))
    return [sorted(s) for s in shards]


def map_round_sharded(lens, map_fn, device, group=None):
    """One alignment round with its queries sharded over the ranks of `group`.
    map_fn(query_indices) -> [bytes per query] maps the given queries against the round's index, which the caller has built
    on THIS rank (replicated; the index build is deterministic).  Rank 0 returns the per-query records in query-index order
    (what the reference's `-j 1` loop yields), the other ranks None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_queries(lens, world)[rank]
    blobs = map_fn(mine) if mine else []
    assert len(blobs) == len(mine)
    got = gather_rounds(list(zip(mine, blobs)), device, group)
    if got is None:
        return None
    assert [i for i, _ in got] == list(range(len(lens))), "a query was mapped twice or not at all"
    return [b for _, b in got]


# ---------------- level 1: a ready-queue over the guide tree ----------------

class TreeSchedule:
    """Ready-queue over a binary guide tree given as children[node] = (left, right) or None for a leaf.
    A merge is READY when both children are done (leaves are done from the start).  `next_wave(world)` returns the ready
    merges dealt to ranks -- heaviest first onto the least-loaded rank, weight = number of leaves below the node, a proxy
    for the consensus length the round aligns -- and `complete(nodes)` releases their parents.  The serial post-order of the
    reference (clade.rs:49-71) is one valid execution of this schedule (world = 1 yields exactly that order restricted to
    internal nodes when waves are taken one node at a time); any order that respects the tree gives the same graph because
    merge_graphs reads nothing but its two children (graph_merging.rs:26-72)."""

    def __init__(self, children):
        self.children = list(children)
        n = len(self.children)
        self.parent = [-1] * n
        for v, ch in enumerate(self.children):
            if ch is not None:
                for c in ch:
                    assert self.parent[c] == -1, "node with two parents"
                    self.parent[c] = v
        roots = [v for v in range(n) if self.parent[v] == -1]
        assert len(roots) == 1, "guide tree must have one root"
        self.root = roots[0]
        self.n_leaves = [0] * n
        self.order = self.postorder()
        for v in self.order:
            ch = self.children[v]
            self.n_leaves[v] = 1 if ch is None else self.n_leaves[ch[0]] + self.n_leaves[ch[1]]
        self.done = [ch is None for ch in self.children]
        self.issued = list(self.done)

    def postorder(self):
        """The reference's visiting order (left, right, node), iteratively."""
        out, stack = [], [(self.root, False)]
        while stack:
            v, seen = stack.pop()
            if seen or self.children[v] is None:
                out.append(v)
                continue
            stack.append((v, True))
            stack.append((self.children[v][1], False))
            stack.append((self.children[v][0], False))
        return out

    def ready(self):
        return [v for v in self.order if not self.issued[v] and all(self.done[c] for c in self.children[v])]

    def next_wave(self, world):
        """[(node, rank)] for every merge that is ready now; marks them issued."""
        load = [(0, r) for r in range(world)]
        heapq.heapify(load)
        wave = []
        for v in sorted(self.ready(), key=lambda v: (-self.n_leaves[v], v)):
            w, r = heapq.heappop(load)
            wave.append((v, r))
            heapq.heappush(load, (w + self.n_leaves[v], r))
            self.issued[v] = True
        return wave

    def complete(self, nodes):
        for v in nodes:
            assert self.issued[v] and not self.done[v]
            self.done[v] = True

    def finished(self):
        return self.done[self.root]


def run_tree(children, leaf_payload, merge_fn, device, group=None, shard_below=0):
    """Drives merges up a guide tree over the ranks of `group`.
    leaf_payload(node) -> bytes for a leaf; merge_fn(node, left_bytes, right_bytes, sub_group) -> bytes runs ONE merge on the
    calling rank (sub_group is None) -- or, when a wave has fewer ready merges than `shard_below` x ranks (near the root), on all
    ranks together with the round's queries sharded (sub_group = group; merge_fn must return the result on rank 0).  After every
    wave the results are gathered on rank 0 and broadcast, so every rank holds the payload of every finished node (graphs near
    the leaves are small; the reweave that produces them is serial on rank 0 in the reference).  Returns {node: bytes}."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sched = TreeSchedule(children)
    have = {v: leaf_payload(v) for v, ch in enumerate(children) if ch is None}
    while not sched.finished():
        wave = sched.next_wave(world)
        if len(wave) < shard_below * world:  # too few merges for the ranks: all ranks work on each merge (level 2)
            results = []
            for v, _ in wave:
                out = merge_fn(v, have[children[v][0]], have[children[v][1]], group)
                results.append((v, out if rank == 0 else b""))
            mine = results if rank == 0 else []
        else:
            mine = [(v, merge_fn(v, have[children[v][0]], have[children[v][1]], None)) for v, r in wave if r == rank]
        got = gather_rounds(mine, device, group)
        blob = [pack_rounds(got)] if rank == 0 else [None]
        dist.broadcast_object_list(blob, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        for v, b in unpack_rounds(blob[0]):
            have[v] = b
        sched.complete([v for v, _ in wave])
    return have
