"""CPU simulation of the hand-off protocol inside one CTA of the tensor-core decode matvec (xalm_b200/csrc/matvec_mma.cuh).

Actors: ONE producer (bulk copies into an NS-stage ring; `full[s]` completes when the bytes land, `empty[s]` when all 16
multiplying warps have released the stage), 16 multiplying warps (wait full -> multiply -> arrive empty; per tile: wait
`pfree[pb]` if the buffer was used before, write their partial sums into `part[pb]`, arrive `pbar[pb]`), ONE epilogue warp
(wait `pbar[pb]` -> read the 16 partial sums -> arrive `pfree[pb]`).  All waits are mbarrier PARITY waits, i.e. a waiter only
learns "the phase with parity q has completed" — which is ambiguous if it can fall two phases behind.

The simulation runs the actors under random schedules with mbarriers modelled as (phase counter, pending arrivals) and checks
what the kernel relies on:
  * nobody ever reads a ring stage or a partial-sum buffer that holds another tile's data, nothing is overwritten before it is read;
  * no parity wait is ever answered by the wrong phase (a waiter is never two completions behind);
  * everybody terminates (no deadlock) for every (NS, stages per tile, tiles) tried, including fewer tiles than buffers.
PB = max(8, NS + 2) as in mma_pb().
"""
import random

import pytest

NCW = 16


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def passed(self, parity):
        """mbarrier.try_wait.parity: true iff the current phase's parity differs from `parity`."""
        return (self.phase & 1) != parity


def simulate(NS, kranges, n_tiles, seed, slow=None):
    """slow: index of an actor that is scheduled 50x less often than the others (0 = producer, 1..16 = multiplying warps,
    17 = epilogue warp) — the adversarial cases: a starved epilogue warp is what the pfree barriers exist for."""
    rng = random.Random(seed)
    PB = max(8, NS + 2)
    full = [MBar(1) for _ in range(NS)]
    empty = [MBar(NCW) for _ in range(NS)]
    pbar = [MBar(NCW) for _ in range(PB)]
    pfree = [MBar(1) for _ in range(PB)]
    ring = [None] * NS                              # (tile, stage) held by a slot
    part = [[None] * NCW for _ in range(PB)]        # tile whose partial sum warp w parked in buffer pb
    n_stages = n_tiles * kranges
    in_flight = []                                  # bulk copies issued, not landed: (slot, tile, stage)

    # ---- actors as generators: yield a condition to wait for, or None for "one step done" ----
    def producer():
        slot, phase = 0, 0
        for g in range(n_stages):
            while not empty[slot].passed(phase ^ 1):
                yield
            # what the parity wait must mean: every warp released the previous use of this slot (g // NS - 1)
            assert empty[slot].phase == g // NS, "producer answered by the wrong phase of empty[]"
            in_flight.append((slot, g // kranges, g % kranges))
            slot += 1
            if slot == NS:
                slot, phase = 0, phase ^ 1
            yield

    def worker(w):
        slot, phase = 0, 0
        pb, pph = 0, 0
        for tt in range(n_tiles):
            for kr in range(kranges):
                g = tt * kranges + kr
                while not full[slot].passed(phase):
                    yield
                assert full[slot].phase == g // NS + 1, "worker answered by the wrong phase of full[]"
                assert ring[slot] == (tt, kr), f"warp {w} read stage {ring[slot]} instead of {(tt, kr)}"
                yield                                # multiply
                empty[slot].arrive()
                slot += 1
                if slot == NS:
                    slot, phase = 0, phase ^ 1
            if tt >= PB:
                while not pfree[pb].passed(pph ^ 1):
                    yield
                assert pfree[pb].phase == tt // PB, "worker answered by the wrong phase of pfree[]"
            assert part[pb][w] is None, f"warp {w} overwrote an unread partial sum in buffer {pb}"
            part[pb][w] = tt
            pbar[pb].arrive()
            pb += 1
            if pb == PB:
                pb, pph = 0, pph ^ 1
            yield

    def epilogue():
        pb, pph = 0, 0
        for tt in range(n_tiles):
            while not pbar[pb].passed(pph):
                yield
            assert pbar[pb].phase == tt // PB + 1, "epilogue answered by the wrong phase of pbar[]"
            assert all(p == tt for p in part[pb]), f"epilogue of tile {tt} read {part[pb]}"
            part[pb] = [None] * NCW
            pfree[pb].arrive()
            pb += 1
            if pb == PB:
                pb, pph = 0, pph ^ 1
            yield

    actors = [producer()] + [worker(w) for w in range(NCW)] + [epilogue()]
    alive = list(range(len(actors)))
    idle = 0
    while alive:
        # the copy engine lands outstanding copies in order, at random times
        if in_flight and rng.random() < 0.5:
            slot, t, k = in_flight.pop(0)
            ring[slot] = (t, k)
            full[slot].arrive()
        i = rng.choice(alive)
        if i == slow and len(alive) > 1 and rng.random() < 0.98:
            continue
        before = (tuple(b.phase for b in full + empty + pbar + pfree), len(in_flight))
        try:
            next(actors[i])
        except StopIteration:
            alive.remove(i)
        after = (tuple(b.phase for b in full + empty + pbar + pfree), len(in_flight))
        idle = idle + 1 if before == after else 0
        assert idle < 200000, "no progress: deadlock"
    assert not in_flight


@pytest.mark.parametrize("NS,kranges,n_tiles", [(5, 2, 13), (4, 7, 3), (2, 1, 40), (8, 1, 30), (6, 2, 1), (3, 4, 11), (5, 1, 7)])
def test_ring_and_partial_sum_protocol_under_random_schedules(NS, kranges, n_tiles):
    for seed in range(4):
        for slow in (None, 0, 3, NCW + 1):
            simulate(NS, kranges, n_tiles, seed, slow)


def test_the_simulation_sees_the_bug_the_pfree_barriers_prevent():
    """Without the pfree wait a starved epilogue warp gets its unread partial sums overwritten: the model must notice."""
    import inspect
    src = inspect.getsource(simulate).replace("if tt >= PB:", "if False:")
    ns = dict(globals())
    exec(src, ns)
    with pytest.raises(AssertionError):
        for seed in range(4):
            ns["simulate"](2, 1, 40, seed, NCW + 1)
