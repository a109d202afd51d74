"""Model check of the multi-stream staged NVLS pipeline (csrc/b200coll.cu: allreduce_streams).

Every rank runs, per piece i, three kernels on three in-order streams chained by events —
copy-in(i) [user -> region i % R], switch(i) [the only kernel that waits for peers: flag A = "my region is
staged", flag B = "I reduced my chunk and multicast it into every rank's region"], copy-out(i) — between an
opening and a closing barrier op on the caller's stream, and whatever the caller enqueues next (here: a
two-shot allreduce that pushes into the peers' staging) follows the closing barrier.

The model executes exactly that dependency structure under random interleavings (any runnable kernel of any
rank may go next; a rank can be made slow or fast) and asserts, for every region of every rank:
  H1  copy-in(i) never overwrites a region whose previous piece has not been copied out by this rank and
      reduced by every peer;
  H2  the switch stage of piece i reads, on every rank, the data staged for piece i;
  H3  copy-out(i) reads the results of piece i (all W chunks);
  H4  the op behind the pipeline never writes a peer's staging while that peer still has a copy-out to do.
It also shows that dropping the closing barrier (H4) or the copy-out -> copy-in event (H1) is caught.
"""
import random

import pytest


class Violation(AssertionError):
    pass


def run(rng, W, P, R, closing_barrier=True, reuse_event=True, bias=None):
    # per rank: three FIFO streams of kernels; a kernel = (name, piece).  State per rank/region.
    IN, SW, OUT = 0, 1, 2
    streams = [[[("in", i) for i in range(P)], [("sw", i) for i in range(P)], [("out", i) for i in range(P)]] for _ in range(W)]
    pos = [[0, 0, 0] for _ in range(W)]
    staged = [[None] * R for _ in range(W)]      # piece whose INPUT is staged in the region
    result = [[set() for _ in range(R)] for _ in range(W)]   # (piece, chunk owner) results present in the region
    read_in = [[set() for _ in range(R)] for _ in range(W)]  # peers that have read the staged input of the region's piece
    sw_phase = [[0] * P for _ in range(W)]       # 0 = not started, 1 = flag A raised (waiting for peers), 2 = done (flag B raised)
    copied_out = [[False] * P for _ in range(W)]
    open_bar = [False] * W                       # opening barrier passed (every peer has started it)
    started = [False] * W
    closed = [False] * W                         # closing barrier passed
    next_op_done = [False] * W
    done = lambda r: all(p >= P for p in pos[r])  # noqa: E731

    def runnable():
        c = []
        for r in range(W):
            if not started[r]:
                c.append((r, "start"))
                continue
            if not open_bar[r]:
                if all(started):
                    c.append((r, "open"))
                continue
            if pos[r][IN] < P:
                i = pos[r][IN]
                if not reuse_event or i < R or copied_out[r][i - R]:
                    c.append((r, "in"))
            if pos[r][SW] < P:
                i = pos[r][SW]
                if pos[r][IN] > i:                                   # event: copy-in(i) done
                    if sw_phase[r][i] == 0:
                        c.append((r, "swA"))
                    elif sw_phase[r][i] == 1 and all(sw_phase[j][i] >= 1 for j in range(W)):
                        c.append((r, "swB"))
            if pos[r][OUT] < P:
                i = pos[r][OUT]
                # event: MY switch(i) kernel is done, and it only finishes after every peer's flag B
                if sw_phase[r][i] == 2 and all(sw_phase[j][i] == 2 for j in range(W)):
                    c.append((r, "out"))
            if not closed[r]:
                # the closing barrier sits on the caller's stream behind the last copy-out ...
                if (pos[r][OUT] >= P if closing_barrier else pos[r][SW] >= P) and all(
                        (pos[j][OUT] >= P if closing_barrier else pos[j][SW] >= P) for j in range(W)):
                    c.append((r, "close"))
            elif not next_op_done[r]:
                c.append((r, "next"))
        return c

    steps = 0
    while not all(next_op_done):
        cands = runnable()
        if not cands:
            raise Violation("deadlock")
        r, what = rng.choices(cands, weights=[bias(x[0]) for x in cands])[0] if bias else rng.choice(cands)
        steps += 1
        if what == "start":
            started[r] = True
        elif what == "open":
            open_bar[r] = True
        elif what == "in":
            i = pos[r][IN]
            reg = i % R
            prev = staged[r][reg]
            if prev is not None:
                if not copied_out[r][prev]:
                    raise Violation(f"H1: rank {r} copy-in({i}) overwrites region {reg} before copy-out({prev})")
                if len(read_in[r][reg]) < W:
                    raise Violation(f"H1: rank {r} copy-in({i}) overwrites region {reg} before every rank reduced piece {prev}")
            staged[r][reg], result[r][reg], read_in[r][reg] = i, set(), set()
            pos[r][IN] += 1
        elif what == "swA":
            sw_phase[r][pos[r][SW]] = 1
        elif what == "swB":
            i = pos[r][SW]
            reg = i % R
            for j in range(W):   # the switch reads every rank's staged region ...
                if staged[j][reg] != i:
                    raise Violation(f"H2: rank {r} switch({i}) reads rank {j} region {reg} holding piece {staged[j][reg]}")
                read_in[j][reg].add(r)
            for j in range(W):   # ... and multicasts chunk r of the result into every rank's region
                result[j][reg].add((i, r))
            sw_phase[r][i] = 2
            pos[r][SW] += 1
        elif what == "out":
            i = pos[r][OUT]
            reg = i % R
            if result[r][reg] != {(i, j) for j in range(W)}:
                raise Violation(f"H3: rank {r} copy-out({i}) finds {sorted(result[r][reg])} in region {reg}")
            copied_out[r][i] = True
            pos[r][OUT] += 1
        elif what == "close":
            closed[r] = True
        elif what == "next":
            # a two-shot behind the pipeline pushes into every peer's staging (any region)
            for j in range(W):
                if j != r and pos[j][OUT] < P:
                    raise Violation(f"H4: rank {r}'s next op writes rank {j}'s staging while copy-out({pos[j][OUT]}) is pending")
            next_op_done[r] = True
    return steps


@pytest.mark.parametrize("W,P,R", [(2, 9, 4), (4, 6, 3), (8, 5, 4), (3, 12, 3)])
def test_pipeline_is_safe_under_random_interleavings(W, P, R):
    rng = random.Random(100 * W + P)
    for trial in range(60):
        slow = rng.randrange(W)
        bias = [None, lambda r: 0.05 if r == slow else 1.0, lambda r: 20.0 if r == slow else 1.0][trial % 3]
        assert run(rng, W, P, R, bias=bias) > 0


def test_model_has_teeth():
    rng = random.Random(5)
    caught = {"H4": 0, "H1": 0}
    for _ in range(80):
        try:
            run(rng, 3, 6, 3, closing_barrier=False, bias=lambda r: 30.0 if r == 0 else 1.0)
        except Violation as e:
            caught["H4"] += str(e).startswith("H4")
        try:
            run(rng, 3, 8, 3, reuse_event=False, bias=lambda r: 30.0 if r == 0 else 1.0)
        except Violation as e:
            caught["H1"] += str(e).startswith("H1")
    assert caught["H4"] > 0 and caught["H1"] > 0, caught
