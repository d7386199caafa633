"""A model of the in-band tagged exchange of the tm2 recurrent kernels (csrc/lstm_recurrent_tm2.cu), run on the CPU under random
schedules.  The kernels have no step counter, no fence and no re-arming: a word carries a one-bit step tag, two buffers alternate
(q & 1), tag = (q >> 1) & 1, consumers poll the words themselves.  The model keeps what the hardware guarantees and nothing more:

  * stores of one thread to ONE address become visible in program order (coherence); stores to different addresses become visible in
    any order and after any delay;
  * a load returns the latest visible value of its address; successive loads of one address never go back in time;
  * a CTA computes step q only from the words of step q-1 it has polled successfully, and stores its own words of step q afterwards.

Checked: every word a consumer accepts for step q-1 IS the step q-1 value of its producer (never q-3, q-5, q+1), nobody waits for
ever, and the same holds for the BPTT direction of the protocol (C partial sums per consumer instead of one word per producer).  With
ONE buffer, or with the tag taken from q & 1, the same model finds the failure within a few steps -- the checks are not vacuous."""
import random

import pytest


class Memory:
    """Per-address FIFO of stores in flight; `tick` makes a random prefix of some queues visible."""

    def __init__(self, rng, init):
        self.rng, self.visible, self.inflight, self.init = rng, {}, {}, init

    def store(self, addr, value):
        self.inflight.setdefault(addr, []).append(value)

    def load(self, addr):
        return self.visible.get(addr, self.init)

    def tick(self):
        for addr in self.rng.sample(list(self.inflight), k=min(len(self.inflight), 1 + self.rng.randrange(4))):
            q = self.inflight[addr]
            self.visible[addr] = q.pop(0)
            if not q:
                del self.inflight[addr]


def run_all_gather(n_cta, steps, seed, nbuf=2, tag_of=lambda q: (q >> 1) & 1, max_events=400000):
    """Forward direction: CTA p publishes word (value, tag) for step q at address (q % nbuf, p); every CTA needs the n_cta words of step
    q-1 before it can compute step q.  Values are (producer, step) pairs so a wrong acceptance is caught by identity, not by luck.
    Returns None, or a string describing the first protocol violation."""
    rng = random.Random(seed)
    mem = Memory(rng, init=(None, 1))              # memset pattern: tag 1, the first two steps carry tag 0
    state = [{"q": 0, "got": set(), "to_store": False} for _ in range(n_cta)]
    for _ in range(max_events):
        if all(s["q"] == steps for s in state):
            return None
        if rng.random() < 0.35:
            mem.tick()
            continue
        c = rng.randrange(n_cta)
        s = state[c]
        q = s["q"]
        if q == steps:
            continue
        if s["to_store"] or q == 0:                # step 0 reads nothing: h(-1) = 0
            mem.store((q % nbuf, c), ((c, q), tag_of(q)))
            s.update(q=q + 1, got=set(), to_store=False)
            continue
        # poll ONE still-missing word of step q-1 (the threads of a CTA poll independently)
        p = rng.choice([p for p in range(n_cta) if p not in s["got"]])
        value, tag = mem.load(((q - 1) % nbuf, p))
        if tag == tag_of(q - 1):
            if value != (p, q - 1):
                return "CTA %d accepted %r for step %d of producer %d" % (c, value, q - 1, p)
            s["got"].add(p)
            if len(s["got"]) == n_cta:
                s["to_store"] = True
    return "no progress: steps reached %r" % [s["q"] for s in state]


@pytest.mark.parametrize("n_cta", [1, 2, 3, 8, 16])
def test_two_buffers_and_a_one_bit_tag_suffice(n_cta):
    for seed in range(12):
        assert run_all_gather(n_cta, steps=40, seed=seed) is None


def test_the_model_finds_the_failures_of_weaker_protocols():
    # one buffer: a fast producer overwrites step q-1 with step q before a slow consumer has read it (the consumer then waits for ever
    # or accepts a value two steps ahead)
    assert any(run_all_gather(3, 40, seed, nbuf=1, tag_of=lambda q: q & 1, max_events=60000) for seed in range(12))
    # two buffers but tag = q & 1: within one buffer every step carries the SAME tag, so the stale step q-3 word is accepted as q-1
    assert any(run_all_gather(3, 40, seed, nbuf=2, tag_of=lambda q: q & 1, max_events=60000) for seed in range(12))
    # no tag at all
    assert any(run_all_gather(3, 40, seed, nbuf=2, tag_of=lambda q: 0, max_events=60000) for seed in range(12))


def run_reduce_scatter(n_cta, steps, seed, max_events=600000):
    """BPTT direction: at step q CTA p stores one partial FOR EACH consumer c at address (q & 1, p, c); consumer c adds the n_cta
    partials of step q-1 addressed to it, then produces step q.  Same tag rule."""
    rng = random.Random(seed)
    mem = Memory(rng, init=(None, 1))
    state = [{"q": 0, "got": set(), "to_store": False} for _ in range(n_cta)]
    tag_of = lambda q: (q >> 1) & 1                                                       # noqa: E731
    for _ in range(max_events):
        if all(s["q"] == steps for s in state):
            return None
        if rng.random() < 0.35:
            mem.tick()
            continue
        c = rng.randrange(n_cta)
        s = state[c]
        q = s["q"]
        if q == steps:
            continue
        if s["to_store"] or q == 0:
            for dst in range(n_cta):
                mem.store((q & 1, c, dst), ((c, dst, q), tag_of(q)))
            s.update(q=q + 1, got=set(), to_store=False)
            continue
        p = rng.choice([p for p in range(n_cta) if p not in s["got"]])
        value, tag = mem.load(((q - 1) & 1, p, c))
        if tag == tag_of(q - 1):
            if value != (p, c, q - 1):
                return "CTA %d accepted %r for step %d of producer %d" % (c, value, q - 1, p)
            s["got"].add(p)
            if len(s["got"]) == n_cta:
                s["to_store"] = True
    return "no progress: steps reached %r" % [s["q"] for s in state]


@pytest.mark.parametrize("n_cta", [1, 2, 5, 8])
def test_reduce_scatter_direction(n_cta):
    for seed in range(8):
        assert run_reduce_scatter(n_cta, steps=30, seed=seed) is None
