"""CPU model check of the peer-exchange protocol of csrc/peer.cuh (C3): randomised interleavings of W
ranks that push / wait / read with epoch flags and payload slots.  It pins the design claim that TWO
payload slots are enough for back-to-back steps (a rank can never overwrite a slot a peer still has to
read) and shows that the same model does catch the hazard with ONE slot.  Pure Python, no GPU: the
CUDA implementation of the same protocol is tested in tests/test_gpu_peer_exchange.py (one GPU) and
tools/peer_exchange_check.py (N GPUs, against NCCL)."""
import random

import pytest


class Rank:
    """Program of one rank, as a generator of atomic actions on the shared state.

    push(e): for every destination (any order): payload store, then flag store (release order per dst)
    wait(e): blocked until all local flags are >= e
    read(e): one action per source, any time before this rank's next push (same stream)
    """

    def __init__(self, r, world, steps, slots, state, rng):
        self.r, self.world, self.steps, self.slots, self.state, self.rng = r, world, steps, slots, state, rng
        self.prog = self._run()
        self.blocked_on = None
        self.done = False

    def _run(self):
        st = self.state
        for e in range(1, self.steps + 1):
            dsts = list(range(self.world))
            self.rng.shuffle(dsts)  # one CTA per destination, unordered
            pending = [("pay", d) for d in dsts]
            while pending:
                i = self.rng.randrange(len(pending))
                kind, d = pending.pop(i)
                if kind == "pay":
                    st["payload"][d][e % self.slots][self.r] = (self.r, e)
                    pending.append(("flag", d))  # the flag of a destination follows its payload
                else:
                    st["flags"][d][self.r] = e
                yield
            while min(st["flags"][self.r]) < e:  # wait(e): flags are monotonic
                self.blocked_on = e
                yield
            self.blocked_on = None
            srcs = list(range(self.world))
            self.rng.shuffle(srcs)
            for s in srcs:  # read(e)
                got = st["payload"][self.r][e % self.slots][s]
                if got != (s, e):
                    st["errors"].append((self.r, e, s, got))
                yield
        self.done = True


def simulate(world, steps, slots, seed, bias=None):
    rng = random.Random(seed)
    state = {"payload": [[[None] * world for _ in range(slots)] for _ in range(world)],
             "flags": [[0] * world for _ in range(world)], "errors": []}
    ranks = [Rank(r, world, steps, slots, state, rng) for r in range(world)]
    idle = 0
    while not all(k.done for k in ranks):
        live = [k for k in ranks if not k.done]
        if bias is not None and rng.random() < 0.9:  # adversarial schedule: one rank runs as far ahead as it can
            k = ranks[bias] if not ranks[bias].done and ranks[bias].blocked_on is None else rng.choice(live)
        else:
            k = rng.choice(live)
        before = (k.blocked_on, [row[:] for row in state["flags"]])
        try:
            next(k.prog)
        except StopIteration:
            k.done = True
        idle = idle + 1 if (k.blocked_on is not None and before[0] is not None) else 0
        assert idle < 20000, "deadlock: every scheduled rank is blocked"
    return state["errors"]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_two_slots_are_enough(world):
    for seed in range(60):
        assert simulate(world, steps=6, slots=2, seed=seed) == []
        assert simulate(world, steps=6, slots=2, seed=1000 + seed, bias=seed % world) == []


def test_model_catches_the_hazard_with_one_slot():
    """With a single slot a fast rank overwrites keys a slow peer has not read yet: the model must see it,
    otherwise the test above would prove nothing."""
    bad = 0
    for seed in range(60):
        bad += bool(simulate(3, steps=6, slots=1, seed=seed, bias=seed % 3))
    assert bad > 0
