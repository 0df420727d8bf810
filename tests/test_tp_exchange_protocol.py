"""The fused tensor-parallel exchange (matvec_tma.cuh) as a protocol, checked on the CPU.

Every rank runs   push(0), recv(0), push(1), recv(1), ...   where push(i) writes {value, tag = base + i + 1} words into slot i & 1
of EVERY rank's receive area (posted NVLink stores, no flag, no fence) and recv(i) polls its own slot until all P sources carry
the tag of exchange i.  Two slots are enough only because of that program order; this test drives P simulated ranks under random
schedules (including stores from one rank landing out of order with respect to another rank's) and checks that

  * nobody deadlocks,
  * a recv never sees a word of another exchange (tags) nor a torn one (value and tag travel in one 8-byte word),
  * a push never overwrites a word its consumer has not read yet,
  * the sequence numbering of forward (ar_base = token_serial * 1024, idx < 2 * n_layers) stays unique across tokens,
  * and that the drain of HYDRATE tokens is NOT optional: a token whose last exchange nobody receives lets a fast rank lap a
    slow one (this simulation found that hole in the first version of the backend; tp_drain_kernel closes it).
"""
import random

import pytest


class Rank:
    def __init__(self, r, P, plan):
        self.r, self.P = r, P
        self.slots = [[(0.0, 0)] * P for _ in range(2)]      # [slot][source] -> (value, tag); zero-initialised like the device buffer
        self.unread = [[False] * P for _ in range(2)]
        self.plan = plan                                      # list of ("push", seq, slot, value) / ("recv", seq, slot)
        self.pc = 0
        self.inflight = []                                    # stores issued by this rank that have not landed yet: (dst, slot, value, tag)


def _plan(n_tokens, n_layers, undrained_every, rank):
    """What one rank executes: per token 2L pushes, each followed by its receive — the next norm-prologue kernel, or the drain
    kernel for the last exchange of a HYDRATE token.  undrained_every > 0 models the broken variant without the drain."""
    plan = []
    for t in range(n_tokens):
        base = t * 1024
        undrained = undrained_every and t % undrained_every == 0
        for idx in range(2 * n_layers):
            seq = base + idx + 1
            plan.append(("push", seq, idx & 1, float(seq * 100 + rank)))
            if not (undrained and idx == 2 * n_layers - 1):
                plan.append(("recv", seq, idx & 1))
    return plan


def _simulate(P, seed, undrained_every):
    rng = random.Random(seed * 17 + P)
    ranks = [Rank(r, P, _plan(n_tokens=6, n_layers=3, undrained_every=undrained_every, rank=r)) for r in range(P)]
    consumed = [[set() for _ in range(P)] for _ in range(P)]   # consumed[dst][src] = tags dst has read from src
    steps = 0
    while any(rk.pc < len(rk.plan) or rk.inflight for rk in ranks):
        steps += 1
        assert steps < 2_000_000, "deadlock or livelock"
        rk = rng.choice(ranks)
        # stores in flight land in random order, at random times (posted writes: no ordering between destinations)
        if rk.inflight and rng.random() < 0.6:
            dst, slot, value, tag = rk.inflight.pop(rng.randrange(len(rk.inflight)))
            d = ranks[dst]
            old_tag = d.slots[slot][rk.r][1]
            # the word being replaced must have been consumed already, unless it belongs to an exchange nobody consumes (hydrate tail)
            assert not d.unread[slot][rk.r], f"rank {rk.r} overwrote unread tag {old_tag} on rank {dst}"
            d.slots[slot][rk.r] = (value, tag)
            d.unread[slot][rk.r] = True
            continue
        if rk.pc >= len(rk.plan):
            continue
        op = rk.plan[rk.pc]
        if op[0] == "push":
            _, seq, slot, value = op
            for dst in range(P):
                rk.inflight.append((dst, slot, value, seq))
            rk.pc += 1
        else:
            _, seq, slot = op
            words = rk.slots[slot]
            if all(tag == seq for _, tag in words):           # the kernel's poll: every source carries this exchange's tag
                for src, (value, tag) in enumerate(words):
                    assert value == float(seq * 100 + src)
                    assert tag not in consumed[rk.r][src]
                    consumed[rk.r][src].add(tag)
                    rk.unread[slot][src] = False
                rk.pc += 1
            else:
                assert all(tag <= seq for _, tag in words), "a word from a LATER exchange arrived before this one was read"
    for rk in ranks:
        assert rk.pc == len(rk.plan)


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_two_slots_suffice_under_any_schedule(P, seed):
    _simulate(P, seed, undrained_every=0)


def test_an_unreceived_exchange_breaks_the_scheme():
    """Without the drain a HYDRATE token leaves its last exchange unreceived; some schedule then overwrites an unread word
    (or starves a receiver).  This is why enqueue_token launches tp_drain_kernel in HYDRATE mode."""
    broken = 0
    for P in (2, 4, 8):
        for seed in range(4):
            try:
                _simulate(P, seed, undrained_every=2)
            except AssertionError:
                broken += 1
    assert broken > 0


def test_sequence_numbers_are_unique_across_tokens():
    for n_layers in (2, 32, 80, 512):
        assert 2 * n_layers <= 1024                           # xalm_cuda.cu: ar_base = token_serial * 1024
        seen = set()
        for t in range(5):
            for idx in range(2 * n_layers):
                seq = t * 1024 + idx + 1
                assert seq not in seen and seq != 0           # 0 is the tag of the zero-initialised buffer
                seen.add(seq)
