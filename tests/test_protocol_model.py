"""Randomised model check of the peer-memory flag protocol (csrc/dev_common.cuh, coll_kernels.cuh).

The reference documents its races instead of detecting them ("TODO: we need a lock here",
nccl_collective_group.py:127,400,415).  The new layer's main risk is the
flag protocol, so this test executes a faithful model of it under random interleavings and
asserts the two safety properties the kernels rely on:

  S1  a staging region is never overwritten while a rank that still has to read its previous
      content has not read it (double buffering by sequence parity + the `arrive` rule of
      coll_prologue());
  S2  every read observes the content written for exactly the reader's own op.

Modelled: W ranks, B blocks per kernel, kernels of one rank strictly ordered (one stream), blocks of
a kernel in arbitrary order, cross-rank flags monotonically increasing, ops = two-shot allreduce,
one-shot allreduce, staged NVLS allreduce (phase-synchronised and round-pipelined), broadcast (root runs
ahead without waiting for anyone; unicast and scatter + multicast-allgather rounds), reduce (non-roots wait
for the root's release), barrier, and the LL (packed data+flag, no prologue wait) small-message allreduce.  Each rank's block executes the same step list the CUDA code
does.  The last test removes the `arrive` wait and shows that the checker then finds a violation —
i.e. the rule is necessary and the model can see that.
"""
import random

import pytest


class Violation(AssertionError):
    pass


class World:
    ROUNDS = 3   # rounds per block of the round-pipelined kernels
    LANE_SLOTS = 3   # staging ring slots of the lane kernel (csrc kLaneSlots)

    def __init__(self, W, B, ops, use_arrive_rule=True):
        self.W, self.B, self.ops, self.rule = W, B, ops, use_arrive_rule
        self.arrive = [[0] * W for _ in range(W)]                       # arrive[rank][src]
        self.flagA = [[[0] * W for _ in range(B)] for _ in range(W)]    # flagA[rank][block][src]
        self.flagB = [[[0] * W for _ in range(B)] for _ in range(W)]
        self.pipeA = [[[0] * W for _ in range(B)] for _ in range(W)]    # per-round flags of the pipelined kernels
        self.pipeB = [[[0] * W for _ in range(B)] for _ in range(W)]
        self.laneIn = [[0] * B for _ in range(W)]                        # lane kernel: copy CTA k -> switch CTA (local)
        self.content = {}     # region -> seq of the data it holds
        self.pending = {}     # region -> set of (rank, block) that still have to read that data
        self.op_idx = [0] * W                                            # kernel each rank is in
        self.pc = [[0] * B for _ in range(W)]                            # step index of each block
        self.steps = [[self._program(r, b, 0) for b in range(B)] for r in range(W)]

    # ---- what the kernels do, as step lists ---------------------------------------------------
    def _program(self, r, b, k):
        if k >= len(self.ops):
            return []
        kind, root = self.ops[k]
        q, W, h = k + 1, self.W, (k + 1) & 1
        peers = [j for j in range(W) if j != r]
        everyone = list(range(W))
        st = []
        R = self.ROUNDS
        # host-side counters every rank advances identically: round-flag epoch and LL op number
        e = R * sum(1 for kk, _ in self.ops[:k] if kk in ("nvls_rounds", "bcast_rounds", "nvls_lanes"))
        ll_no = 1 + sum(1 for kk, _ in self.ops[:k] if kk == "ll")
        if kind == "ll":
            # no prologue wait; data and flag travel together; arrive is published last
            lh = ("ll", ll_no & 1)
            st += [("write", (j, lh, r, b), ("ll", ll_no), {(j, b)}) for j in peers]
            st += [("ll_read", (r, lh, s, b), ("ll", ll_no)) for s in peers]
            if b == 0:
                st += [("set_arrive", j, q) for j in peers]
            return st
        if kind == "barrier":
            if b == 0:
                st += [("set_arrive", j, q) for j in peers] + [("wait_arrive", j, q) for j in peers]
            return st
        # coll_prologue: block 0 publishes arrive = q; every block waits arrive >= q-1 from every peer
        if b == 0:
            st += [("set_arrive", j, q) for j in peers]
        if self.rule:
            st += [("wait_arrive", j, q - 1) for j in peers]
        if kind == "twoshot":
            st += [("write", (j, h, r, b), q, {(j, b)}) for j in peers]                      # A: push chunk j to rank j's slot r
            st += [("sigA", j, q) for j in peers] + [("waitA", j, q) for j in peers]
            st += [("read", (r, h, s, b), q) for s in peers]                                 # B: fold the W contributions
            st += [("write", (r, h, r, b), q, {(j, b) for j in peers})]                      #    result -> own slot (peers pull it)
            st += [("sigB", j, q) for j in peers] + [("waitB", j, q) for j in peers]
            st += [("read", (j, h, j, b), q) for j in peers]                                 # C: pull the reduced chunks
        elif kind == "oneshot":
            st += [("write", (j, h, r, b), q, {(j, b)}) for j in peers]
            st += [("sigA", j, q) for j in peers] + [("waitA", j, q) for j in peers]
            st += [("read", (r, h, s, b), q) for s in peers]
        elif kind == "nvls":
            everyone = list(range(W))
            st += [("write", (r, h, "in", b), q, {(j, b) for j in everyone})]                # A: stage in (own arena)
            st += [("sigA", j, q) for j in peers] + [("waitA", j, q) for j in peers]
            st += [("read", (j, h, "in", b), q) for j in everyone]                           # B: the switch reads every rank's copy
            st += [("write", (j, h, ("out", r), b), q, {(j, b)}) for j in everyone]          #    ... and multicasts the result back
            st += [("sigB", j, q) for j in peers] + [("waitB", j, q) for j in peers]
            st += [("read", (r, h, ("out", j), b), q) for j in everyone]                     # C: stage out
        elif kind == "nvls_rounds":
            # software pipeline of k_allreduce_nvls_rounds: in(q+1) | waitA(q) switch(q) sigB(q) | waitB(q-1) out(q-1)
            def stage_in(x):
                return [("write", (r, h, ("in", x), b), q, {(j, b) for j in everyone})] + [("sigPA", j, e + x + 1) for j in peers]

            def stage_out(x):
                return [("waitPB", j, e + x + 1) for j in peers] + [("read", (r, h, ("out", j, x), b), q) for j in everyone]

            st += stage_in(0)
            for x in range(R):
                if x + 1 < R:
                    st += stage_in(x + 1)
                st += [("waitPA", j, e + x + 1) for j in peers]
                st += [("read", (j, h, ("in", x), b), q) for j in everyone]
                st += [("write", (j, h, ("out", r, x), b), q, {(j, b)}) for j in everyone]
                st += [("sigPB", j, e + x + 1) for j in peers]
                if x >= 1:
                    st += stage_out(x - 1)
            st += stage_out(R - 1)
        elif kind == "nvls_lanes":
            # one lane: block 0 is the switch CTA, blocks 1..B-1 are copy CTAs; the staging ring has 3 slots that are
            # rewritten every three rounds (k_allreduce_nvls_lanes).  Regions: ("ring", slot, chunk, share k).
            Kc = self.B - 1
            if Kc == 0:
                return st  # a lane needs at least one copy CTA; the host never launches such a grid
            tag = lambda x: (q, x)  # noqa: E731  data of round x of this op
            if b == 0:
                for x in range(R):
                    st += [("wait_lane_in", kk, e + x + 1) for kk in range(Kc)]
                    st += [("sigPA0", j, e + x + 1) for j in peers] + [("waitPA0", j, e + x + 1) for j in peers]
                    st += [("read", (j, h, ("ring", x % self.LANE_SLOTS, r, kk), 0), tag(x)) for j in everyone for kk in range(Kc)]
                    # the switch writes the reduced chunk r back into every rank's slot: the next readers are that
                    # rank's copy CTAs (copy-out)
                    st += [("write", (j, h, ("ring", x % self.LANE_SLOTS, r, kk), 0), ("red", q, x), {(j, kk + 1)}) for j in everyone for kk in range(Kc)]
                    st += [("sigPB0", j, e + x + 1) for j in everyone]
            else:
                kk = b - 1
                for x in range(R + 2):
                    if x < R:
                        # my share of every chunk's granule; readers: the switch CTA (block 0) of the chunk's owner
                        st += [("write", (r, h, ("ring", x % self.LANE_SLOTS, j, kk), 0), tag(x), {(j, 0)}) for j in everyone]
                        st += [("set_lane_in", kk, e + x + 1)]
                    if x >= 2:
                        st += [("waitPB0", j, e + x - 1) for j in everyone]
                        st += [("read", (r, h, ("ring", (x - 2) % self.LANE_SLOTS, j, kk), 0), ("red", q, x - 2)) for j in everyone]
        elif kind == "bcast_rounds":
            others = [j for j in everyone if j != root]
            if r == root:
                for x in range(R):
                    st += [("write", (j, h, ("sc", x), b), q, {(j, b)}) for j in peers]                 # unicast scatter
                    st += [("write", (j, h, ("mc", r, x), b), q, {(j, b)}) for j in peers]              # own chunk multicast
                    st += [("sigPA", j, e + x + 1) for j in peers] + [("sigPB", j, e + x + 1) for j in peers]
            else:
                for x in range(R + 1):
                    if x < R:
                        st += [("waitPA", root, e + x + 1), ("read", (r, h, ("sc", x), b), q)]
                        st += [("write", (j, h, ("mc", r, x), b), q, ({(j, b)} if j in others and j != r else set())) for j in everyone if j != r]
                        st += [("sigPB", j, e + x + 1) for j in peers]
                    if x >= 1:
                        st += [("waitPB", j, e + x) for j in peers]
                        st += [("read", (r, h, ("mc", j, x - 1), b), q) for j in everyone if j != r]
        elif kind == "broadcast":
            if r == root:
                st += [("write", (j, h, 0, b), q, {(j, b)}) for j in peers]                  # root pushes, waits for nobody
                st += [("sigA", j, q) for j in peers]
            else:
                st += [("waitA", root, q), ("read", (r, h, 0, b), q)]
        elif kind == "reduce":
            if r != root:
                st += [("write", (root, h, r, b), q, {(root, b)}), ("sigA", root, q), ("waitB", root, q)]
            else:
                st += [("waitA", j, q) for j in peers] + [("read", (r, h, s, b), q) for s in peers]
                st += [("sigB", j, q) for j in peers]
        else:
            raise ValueError(kind)
        return st

    # ---- execution ----------------------------------------------------------------------------
    def runnable(self):
        out = []
        for r in range(self.W):
            for b in range(self.B):
                if self.pc[r][b] < len(self.steps[r][b]) and self._ready(r, b, self.steps[r][b][self.pc[r][b]]):
                    out.append((r, b))
        return out

    def _ready(self, r, b, step):
        op = step[0]
        if op == "wait_arrive":
            return self.arrive[r][step[1]] >= step[2]
        if op == "waitA":
            return self.flagA[r][b][step[1]] >= step[2]
        if op == "waitB":
            return self.flagB[r][b][step[1]] >= step[2]
        if op == "waitPA":
            return self.pipeA[r][b][step[1]] >= step[2]
        if op == "waitPB":
            return self.pipeB[r][b][step[1]] >= step[2]
        if op == "wait_lane_in":
            return self.laneIn[r][step[1]] >= step[2]
        if op == "waitPA0":
            return self.pipeA[r][0][step[1]] >= step[2]
        if op == "waitPB0":
            return self.pipeB[r][0][step[1]] >= step[2]
        if op == "ll_read":  # the receiver polls the slot itself: ready once the flag of THIS op is there
            return self.content.get(step[1]) == step[2]
        return True

    def step(self, r, b):
        s = self.steps[r][b][self.pc[r][b]]
        op = s[0]
        if op == "set_arrive":
            self.arrive[s[1]][r] = max(self.arrive[s[1]][r], s[2])
        elif op == "sigA":
            self.flagA[s[1]][b][r] = s[2]
        elif op == "sigB":
            self.flagB[s[1]][b][r] = s[2]
        elif op == "sigPA":
            assert s[2] > self.pipeA[s[1]][b][r], "round flags must increase monotonically"
            self.pipeA[s[1]][b][r] = s[2]
        elif op == "sigPB":
            assert s[2] > self.pipeB[s[1]][b][r], "round flags must increase monotonically"
            self.pipeB[s[1]][b][r] = s[2]
        elif op == "set_lane_in":
            assert s[2] > self.laneIn[r][s[1]]
            self.laneIn[r][s[1]] = s[2]
        elif op == "sigPA0":
            assert s[2] > self.pipeA[s[1]][0][r]
            self.pipeA[s[1]][0][r] = s[2]
        elif op == "sigPB0":
            assert s[2] > self.pipeB[s[1]][0][r]
            self.pipeB[s[1]][0][r] = s[2]
        elif op == "write":
            region, q, readers = s[1], s[2], s[3]
            # Conservative aliasing: one-shot and two-shot lay slots and tiles out differently inside a
            # half, so a write may land on ANY older data of the same (rank, half).  Data of the same op
            # is disjoint by construction (distinct slot / tile per writer).
            def op_of(tag):  # data tags of the lane kernel are (op, round) / ("red", op, round)
                return tag if not isinstance(tag, tuple) else (tag[1] if tag[0] == "red" else tag[0])

            ring = isinstance(region[2], tuple) and region[2][0] == "ring"
            for other, seq in self.content.items():
                if ring and isinstance(other[2], tuple) and other[2][0] == "ring" and op_of(seq) == op_of(q):
                    # same op, ring addressing is exact: only the very same slot/chunk/share aliases
                    if other == region and seq != q and self.pending.get(other):
                        raise Violation(f"S1: rank {r} block {b} overwrites ring region {region} (now {q}) while "
                                        f"{sorted(self.pending[other])} still have to read {seq}")
                    continue
                if other[:2] == region[:2] and op_of(seq) != op_of(q) and self.pending.get(other):
                    raise Violation(f"S1: rank {r} block {b} op {q} writes {region} while {sorted(self.pending[other])} "
                                    f"still have to read op {seq} data in {other} of the same staging half")
            self.content[region], self.pending[region] = q, set(readers)
        elif op in ("read", "ll_read"):
            region, q = s[1], s[2]
            if self.content.get(region) != q:
                raise Violation(f"S2: rank {r} block {b} op {q} reads {region} holding op {self.content.get(region)}")
            self.pending[region].discard((r, b))
        self.pc[r][b] += 1
        # kernel boundary: the next kernel of this rank starts only when every block of this one is done
        if all(self.pc[r][x] >= len(self.steps[r][x]) for x in range(self.B)) and self.op_idx[r] < len(self.ops):
            self.op_idx[r] += 1
            for x in range(self.B):
                self.steps[r][x] = self._program(r, x, self.op_idx[r])
                self.pc[r][x] = 0

    def run(self, rng, bias=None):
        """Random scheduler; `bias(r, b)` may weight the choice (e.g. make one rank slow)."""
        n = 0
        while True:
            if all(self.op_idx[r] >= len(self.ops) for r in range(self.W)):
                return n
            cands = self.runnable()
            if not cands:
                raise Violation(f"deadlock: op_idx={self.op_idx}")
            if bias is not None:
                weights = [bias(r, b) for r, b in cands]
                r, b = rng.choices(cands, weights=weights)[0]
            else:
                r, b = rng.choice(cands)
            self.step(r, b)
            n += 1


def random_ops(rng, W, n):
    kinds = ["twoshot", "oneshot", "nvls", "nvls_rounds", "nvls_lanes", "ll", "broadcast", "bcast_rounds", "reduce", "barrier"]
    return [(k, rng.randrange(W)) for k in (rng.choice(kinds) for _ in range(n))]


@pytest.mark.parametrize("W,B", [(2, 1), (2, 3), (3, 2), (4, 2), (8, 1)])
def test_protocol_is_safe_under_random_interleavings(W, B):
    rng = random.Random(1000 * W + B)
    total = 0
    for trial in range(60):
        ops = random_ops(rng, W, 12)
        slow = rng.randrange(W)
        bias = [None, lambda r, b: 0.05 if r == slow else 1.0, lambda r, b: 20.0 if r == slow else 1.0][trial % 3]
        total += World(W, B, ops).run(rng, bias)
    assert total > 0


def test_broadcast_root_runs_ahead_but_never_two_ops():
    """A producer-only root may be one op ahead of a slow reader (other staging half), never two."""
    rng = random.Random(7)
    ops = [("broadcast", 0)] * 10
    for _ in range(50):
        w = World(3, 2, ops)
        lead = 0
        while not all(i >= len(ops) for i in w.op_idx):
            cands = w.runnable()
            # starve rank 2: it only runs when nothing else can
            pick = [c for c in cands if c[0] != 2] or cands
            w.step(*rng.choice(pick))
            lead = max(lead, w.op_idx[0] - w.op_idx[2])
        assert 1 <= lead <= 2  # "2" = root has *entered* kernel q+2's prologue and is parked on arrive


def test_checker_has_teeth_without_the_arrive_rule():
    """Drop the prologue wait: a run-ahead broadcast root overwrites a staging half a slow reader has
    not read yet.  The model must catch it, otherwise the test above proves nothing."""
    rng = random.Random(3)
    ops = [("broadcast", 0)] * 6
    caught = 0
    for _ in range(40):
        try:
            World(3, 1, ops, use_arrive_rule=False).run(rng, lambda r, b: 50.0 if r == 0 else 1.0)
        except Violation as e:
            assert str(e).startswith(("S1", "S2"))
            caught += 1
    assert caught > 0


def test_mixed_algorithms_share_the_staging_safely():
    """one-shot and two-shot lay their slots out differently inside the same half; alternate them with
    asymmetric ops in between and with one rank much slower than the others."""
    rng = random.Random(11)
    ops = [("twoshot", 0), ("broadcast", 1), ("oneshot", 0), ("reduce", 2), ("twoshot", 0), ("broadcast", 0), ("oneshot", 0)] * 3
    for slow in range(4):
        World(4, 2, ops).run(rng, lambda r, b: 0.02 if r == slow else 1.0)


def test_lane_kernel_ring_reuse_is_safe_and_two_slots_would_not_be():
    """The lane kernel rewrites a three-slot staging ring every three rounds.  Many rounds, a slow and a fast
    rank, one to three copy CTAs per lane."""
    rng = random.Random(31)
    old = World.ROUNDS
    World.ROUNDS = 7
    try:
        ops = [("nvls_lanes", 0), ("ll", 0), ("nvls_lanes", 0), ("broadcast", 1), ("nvls_lanes", 0), ("twoshot", 0)]
        for W, B in ((2, 2), (3, 3), (4, 4), (8, 2)):
            for slow in range(min(W, 2)):
                for weight in (0.03, 25.0):
                    World(W, B, ops).run(rng, lambda r, b: weight if r == slow else 1.0)
        # teeth: with only two slots in(q) would overwrite the slot out(q-2) still has to read
        World.LANE_SLOTS = 2
        caught = 0
        for _ in range(20):
            try:
                World(2, 2, [("nvls_lanes", 0)]).run(rng)
            except Violation:
                caught += 1
        assert caught > 0
    finally:
        World.ROUNDS = old
        World.LANE_SLOTS = 3


def test_ll_and_round_pipelined_kernels_between_asymmetric_ops():
    """LL skips the prologue wait and the round kernels share one flag epoch: interleave them with
    producer-only ops (a root that runs ahead) and a very slow / very fast rank."""
    rng = random.Random(23)
    ops = [("ll", 0), ("broadcast", 1), ("ll", 0), ("ll", 0), ("nvls_rounds", 0), ("bcast_rounds", 2), ("ll", 0),
           ("bcast_rounds", 0), ("nvls_rounds", 0), ("twoshot", 0), ("ll", 0), ("reduce", 1), ("ll", 0)] * 2
    for W, B in ((3, 2), (4, 1), (8, 1)):
        for slow in range(min(W, 3)):
            for weight in (0.02, 30.0):
                World(W, B, ops).run(rng, lambda r, b: weight if r == slow else 1.0)


# ---------------------------------------------------------------------------------------------
# p2p ring (k_send / k_recv): cell k of a pair's lifetime lives at ring position k % CELLS and
# carries flag value k+1; the sender may reuse a position only after the receiver acked its
# previous occupant.  Blocks of a kernel take cells i, i+grid, ...
# ---------------------------------------------------------------------------------------------
SEND_BATCH = 4   # csrc/byte_kernels.cuh kSendBatch: cells a sender block publishes per release fence


def _run_ring(rng, cells, messages, send_grid, recv_grid, recv_bias):
    ready, ack, ring = [0] * cells, [0] * cells, [None] * cells
    unread = [False] * cells
    # per-block work lists for the whole message sequence, kernel by kernel.  A sender block takes
    # SEND_BATCH cells per pass (stride grid), waits for the acks of ALL of them, copies, then raises
    # their ready flags together; the host caps the sender grid at cells // SEND_BATCH so that one pass
    # over the grid fits in the ring.  A receiver block takes one cell at a time.
    def send_kernels(grid):
        out, first = [], 0
        for n in messages:
            g = max(1, min((n + SEND_BATCH - 1) // SEND_BATCH, grid, max(1, cells // SEND_BATCH)))
            per_block = []
            for b in range(g):
                mine = [first + i for i in range(b, n, g)]
                per_block.append([mine[p:p + SEND_BATCH] for p in range(0, len(mine), SEND_BATCH)])
            out.append(per_block)
            first += n
        return out

    def recv_kernels(grid):
        out, first = [], 0
        for n in messages:
            g = min(n, grid, cells)
            out.append([[first + i for i in range(b, n, g)] for b in range(g)])
            first += n
        return out
    S, R = send_kernels(send_grid), recv_kernels(recv_grid)
    si = ri = 0
    spc, rpc = [0] * len(S[0]), [0] * len(R[0])
    got = []
    while si < len(S) or ri < len(R):
        cands = []
        if si < len(S):
            for b, lst in enumerate(S[si]):
                if spc[b] < len(lst) and all(k < cells or ack[k % cells] >= k + 1 - cells for k in lst[spc[b]]):
                    cands.append(("s", b))
        if ri < len(R):
            for b, lst in enumerate(R[ri]):
                if rpc[b] < len(lst) and ready[lst[rpc[b]] % cells] >= lst[rpc[b]] + 1:
                    cands.append(("r", b))
        assert cands, "p2p ring deadlock"
        side, b = rng.choices(cands, weights=[recv_bias if c[0] == "r" else 1.0 for c in cands])[0]
        if side == "s":
            for k in S[si][b][spc[b]]:
                pos = k % cells
                assert not unread[pos], f"cell {k} overwrites ring position {pos} before it was consumed"
                ring[pos], unread[pos], ready[pos] = k, True, k + 1
            spc[b] += 1
            if all(spc[x] >= len(S[si][x]) for x in range(len(S[si]))):
                si += 1
                spc = [0] * len(S[si]) if si < len(S) else []
        else:
            k = R[ri][b][rpc[b]]
            pos = k % cells
            assert ring[pos] == k, f"receiver expected cell {k} at position {pos}, found {ring[pos]}"
            got.append(k)
            unread[pos], ack[pos] = False, k + 1
            rpc[b] += 1
            if all(rpc[x] >= len(R[ri][x]) for x in range(len(R[ri]))):
                ri += 1
                rpc = [0] * len(R[ri]) if ri < len(R) else []
    assert sorted(got) == list(range(sum(messages)))


@pytest.mark.parametrize("cells", [4, 16])
def test_p2p_ring_flow_control(cells):
    rng = random.Random(cells)
    for trial in range(200):
        messages = [rng.randint(1, 3 * cells) for _ in range(rng.randint(1, 6))]
        _run_ring(rng, cells, messages, send_grid=rng.randint(1, cells), recv_grid=rng.randint(1, cells),
                  recv_bias=[0.05, 1.0, 20.0][trial % 3])
