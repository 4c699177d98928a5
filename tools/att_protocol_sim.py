"""Discrete-event model of the attention kernel's synchronisation protocol (csrc/attention.cu; it also still models the
round-1 experiment variants - CTA pair, S-first, P-alias - that were measured on hardware in round 2 and removed from the
tree, see profiles/attention_r2.md: the model is how their protocols were validated before they ever ran):
mbarriers with phase parity and transaction counts, the TMA producer(s), the single UMMA-issuing thread, the in-order
tensor pipe with tcgen05.commit arrivals (multicast for the CTA pair), and the softmax warpgroups. Every buffer (Q tiles,
K / V ring slots, S_t, P_t, O_t) carries a version tag; each consumer asserts that it sees exactly the version it is
meant to see, every wait asserts that it was released by the phase it was written for (no parity aliasing), and the run
must terminate (no deadlock). Actors are stepped in random order over many seeds.

It does NOT model timing or the hardware - it checks that the waits / arrives / commits, ring indices and parities written
in the kernels form a correct protocol for every interleaving tried. Variants: ctas = 1 | 2 (CTA pair), order = "default" |
"s_first" (-DL4P_ATT_S_FIRST) | "alias" (-DL4P_ATT_P_ALIAS).

    python tools/att_protocol_sim.py            # all variants, 200 seeds each
"""
from __future__ import annotations

import random
import sys
from collections import deque


class Deadlock(Exception):
    pass


class Barrier:
    def __init__(self, name, count):
        self.name, self.count = name, count
        self.pending, self.tx, self.phase = count, 0, 0          # phase = number of completed phases

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, n=1):
        assert self.pending >= n, f"{self.name}: too many arrivals in phase {self.phase}"
        self.pending -= n
        self._maybe_complete()

    def expect_tx(self, nbytes):          # mbarrier.arrive.expect_tx
        self.tx += nbytes
        self.arrive(1)

    def complete_tx(self, nbytes):        # TMA completion
        self.tx -= nbytes
        self._maybe_complete()

    def ready(self, parity):              # mbarrier.try_wait.parity
        return (self.phase & 1) != parity


def wait(bar: Barrier, parity: int, intended_phase: int):
    """Generator: block until try_wait.parity passes; then check it was released by the intended phase (an aliased parity
    would let the waiter through one or more phases too early or too late)."""
    while not bar.ready(parity):
        yield
    assert bar.phase == intended_phase + 1, (f"{bar.name}: wait for phase {intended_phase} (parity {parity}) passed at "
                                             f"completed-phase count {bar.phase}")


class Sim:
    def __init__(self, ctas=1, order="default", nblk=16, kKS=2, kVS=2, seed=0):
        self.n, self.order, self.nblk, self.kKS, self.kVS = ctas, order, nblk, kKS, kVS
        self.rng = random.Random(seed)
        R = range(ctas)
        thr = 128 * ctas                                           # softmax threads of one tile across the CTA group
        L = 0                                                      # leader CTA
        self.bar = {}

        def mk(name, count, ranks):
            for r in ranks:
                self.bar[(name, r)] = Barrier(f"{name}@cta{r}", count)

        mk("q", 1, [L])
        for s in range(kKS):
            mk(f"kfull{s}", 1, [L]); mk(f"kempty{s}", 1, R)
        for s in range(kVS):
            mk(f"vfull{s}", 1, [L]); mk(f"vempty{s}", 1, R)
        for t in range(2):
            mk(f"sfull{t}", 1, R); mk(f"pvdone{t}", 1, R)
            mk(f"sfree{t}", thr, [L]); mk(f"pfull{t}", thr, [L])
        # buffers: version tags
        self.Q = {r: None for r in R}
        self.K = {(r, s): None for r in R for s in range(kKS)}
        self.V = {(r, s): None for r in R for s in range(kVS)}
        self.S = {(r, t): None for r in R for t in range(2)}       # ("S", j) after the S UMMAs, ("P", j) when P aliases S
        self.P = {(r, t): None for r in R for t in range(2)}
        self.O = {(r, t): -1 for r in R for t in range(2)}         # last key block accumulated
        self.kused = {(r, s): 2 for r in R for s in range(kKS)}    # UMMA groups that have read the slot's current block
        self.vused = {(r, s): 2 for r in R for s in range(kVS)}    # (2 = both tiles done = slot free)
        self.mutate = None                                         # self-test: name of a deliberately broken rule
        self.pipe = deque()                                        # in-order tensor pipe of the issuing thread
        self.tma = []                                              # in-flight TMA loads (complete in any order)
        self.done = {}

    # ------------------------------------------------------------------ actors
    def producer(self, r):
        lead = self.bar[("q", 0)]
        if r == 0:
            lead.expect_tx(2 * self.n)                             # 2 tiles per CTA
        for t in range(2):
            self.tma.append(("Q", r, t, lead, 1))
        yield
        for j in range(self.nblk):
            for kind, ring in (("K", self.kKS), ("V", self.kVS)):
                s = j % ring
                emp = self.bar[(f"{kind.lower()}empty{s}", r)]
                yield from wait(emp, ((j // ring) & 1) ^ 1, j // ring - 1)
                full = self.bar[(f"{kind.lower()}full{s}", 0)]
                if r == 0:
                    full.expect_tx(self.n)                         # one unit of bytes per CTA's half
                self.tma.append((kind, r, s, full, 1, j))
                yield
        self.done[("prod", r)] = True

    def tma_engine(self):
        while True:
            if self.tma:
                i = self.rng.randrange(len(self.tma))
                op = self.tma.pop(i)
                if op[0] == "Q":
                    _, r, t, bar, n = op
                    self.Q[r] = "Q"
                else:
                    kind, r, s, bar, n, j = op
                    buf, used = (self.K, self.kused) if kind == "K" else (self.V, self.vused)
                    # the slot must be free: both tiles' UMMAs on its previous block have executed
                    assert used[(r, s)] == 2, f"TMA overwrites {kind} slot {s} of cta{r} (block {buf[(r, s)]}) while in use"
                    buf[(r, s)] = j
                    used[(r, s)] = 0
                bar.complete_tx(n)
            yield

    def issue_s(self, t, s, jn):
        def run():
            for r in range(self.n):
                assert self.Q[r] == "Q", "S UMMA before Q landed"
                assert self.K[(r, s)] == jn, f"S_{t}({jn}): K slot {s} of cta{r} holds block {self.K[(r, s)]}"
                if self.order == "alias":
                    assert self.S[(r, t)] in (None, ("P", jn - 1)), f"S_{t}({jn}) overwrites {self.S[(r, t)]}"
                self.S[(r, t)] = ("S", jn)
                self.kused[(r, s)] += 1
        self.pipe.append(("mma", run))
        self.pipe.append(("commit", [self.bar[(f"sfull{t}", r)] for r in range(self.n)]))

    def issue_pv(self, t, s, j):
        def run():
            for r in range(self.n):
                assert self.V[(r, s)] == j, f"PV_{t}({j}): V slot {s} of cta{r} holds block {self.V[(r, s)]}"
                src = self.S[(r, t)] if self.order == "alias" else self.P[(r, t)]
                assert src == ("P", j), f"PV_{t}({j}) reads {src} in cta{r}"
                assert self.O[(r, t)] == j - 1, f"O_{t} of cta{r} at block {self.O[(r, t)]} when PV({j}) runs"
                self.O[(r, t)] = j
                self.vused[(r, s)] += 1
        self.pipe.append(("mma", run))
        self.pipe.append(("commit", [self.bar[(f"pvdone{t}", r)] for r in range(self.n)]))

    def commit(self, name):
        self.pipe.append(("commit", [self.bar[(name, r)] for r in range(self.n)]))

    def issuer(self):
        B, nblk, kKS, kVS = self.bar, self.nblk, self.kKS, self.kVS
        yield from wait(B[("q", 0)], 0, 0)
        yield from wait(B[("kfull0", 0)], 0, 0)
        self.issue_s(0, 0, 0); self.issue_s(1, 0, 0); self.commit("kempty0")
        yield
        for j in range(nblk):
            jn, sk, sv = j + 1, (j + 1) % kKS, j % kVS
            if self.order == "alias":
                for t in range(2):
                    yield from wait(B[(f"pfull{t}", 0)], j & 1, j)
                    if t == 0:
                        yield from wait(B[(f"vfull{sv}", 0)], (j // kVS) & 1, j // kVS)
                    if jn < nblk and t == 0:
                        yield from wait(B[(f"kfull{sk}", 0)], (jn // kKS) & 1, jn // kKS)
                    self.issue_pv(t, sv, j)
                    if t == 1:
                        self.commit(f"vempty{sv}")
                    if jn < nblk:
                        self.issue_s(t, sk, jn)
                        if t == 1:
                            self.commit(f"kempty{sk}")
                    yield
                if j == nblk - 1:
                    yield from self.drain()
                    self.done["issuer"] = True
                continue
            if self.order == "s_first":
                if jn < nblk:
                    yield from wait(B[(f"kfull{sk}", 0)], (jn // kKS) & 1, jn // kKS)
                    for t in range(2):
                        yield from wait(B[(f"sfree{t}", 0)], j & 1, j)
                        self.issue_s(t, sk, jn)
                        if t == 1:
                            self.commit(f"kempty{sk}")
                        yield
                for t in range(2):
                    yield from wait(B[(f"pfull{t}", 0)], j & 1, j)
                    if t == 0:
                        yield from wait(B[(f"vfull{sv}", 0)], (j // kVS) & 1, j // kVS)
                    self.issue_pv(t, sv, j)
                    if t == 1:
                        self.commit(f"vempty{sv}")
                    yield
                if j == nblk - 1:
                    yield from self.drain()
                    self.done["issuer"] = True
                continue
            for t in range(2):
                if jn < nblk:
                    if t == 0:
                        yield from wait(B[(f"kfull{sk}", 0)], (jn // kKS) & 1, jn // kKS)
                    if self.mutate != "no_sfree_wait":
                        yield from wait(B[(f"sfree{t}", 0)], j & 1, j)
                    self.issue_s(t, sk, jn)
                    if t == 1 or self.mutate == "early_kempty":
                        self.commit(f"kempty{sk}")
                    yield
                yield from wait(B[(f"pfull{t}", 0)], j & 1, j)
                if t == 0:
                    yield from wait(B[(f"vfull{sv}", 0)], (j // kVS) & 1, j // kVS)
                self.issue_pv(t, sv, j)
                if t == 1:
                    self.commit(f"vempty{sv}")
                yield
        yield from self.drain()
        self.done["issuer"] = True

    def drain(self):
        """CTA pair only: the last commit (V slot release, multicast) has no other waiter; the issuer waits for its local
        arrival so that no multicast arrive is in flight towards the peer's shared memory when the cluster exits."""
        if self.n == 2:
            last = self.nblk - 1
            yield from wait(self.bar[(f"vempty{last % self.kVS}", 0)], (last // self.kVS) & 1, last // self.kVS)

    def tensor_pipe(self):
        while True:
            if self.pipe:
                kind, x = self.pipe.popleft()
                if kind == "mma":
                    x()
                else:
                    for b in x:
                        b.arrive(1)
            yield

    def softmax(self, r, t):
        B = self.bar
        for j in range(self.nblk):
            yield from wait(B[(f"sfull{t}", r)], j & 1, j)
            assert self.S[(r, t)] == ("S", j), f"softmax_{t}({j}) of cta{r} reads {self.S[(r, t)]}"
            yield                                                   # tcgen05.ld of S_t into registers
            assert self.S[(r, t)] == ("S", j), f"S_{t}({j}) of cta{r} overwritten while being loaded"
            B[(f"sfree{t}", 0)].arrive(128)
            yield                                                   # max / (rare) rescale / exp
            if j > 0:
                yield from wait(B[(f"pvdone{t}", r)], (j - 1) & 1, j - 1)
                assert self.O[(r, t)] == j - 1
            if self.order == "alias":
                assert self.S[(r, t)] == ("S", j), f"P_{t}({j}) would overwrite {self.S[(r, t)]}"
                self.S[(r, t)] = ("P", j)
            else:
                self.P[(r, t)] = ("P", j)
            B[(f"pfull{t}", 0)].arrive(128)
            yield
        yield from wait(B[(f"pvdone{t}", r)], (self.nblk - 1) & 1, self.nblk - 1)
        assert self.O[(r, t)] == self.nblk - 1
        self.done[("softmax", r, t)] = True

    # ------------------------------------------------------------------ scheduler
    def run(self, max_steps=2_000_000):
        actors = {"issuer": self.issuer()}
        for r in range(self.n):
            actors[("prod", r)] = self.producer(r)
            for t in range(2):
                actors[("softmax", r, t)] = self.softmax(r, t)
        engines = [self.tma_engine(), self.tensor_pipe()]
        want = set(actors)
        idle = 0
        for _ in range(max_steps):
            if want <= set(self.done):
                assert not self.tma
                if self.n == 2:   # the pair kernel's issuer drains its last multicast commit before the cluster exits
                    assert not self.pipe, "CTA pair exits with a multicast commit still in flight"
                while self.pipe:  # single CTA: the trailing V-slot release lands in the CTA's own (still live) smem
                    next(engines[1])
                return True
            live = [k for k in actors if k not in self.done]
            pick = self.rng.random()
            before = self._state()
            if pick < 0.3:
                next(self.rng.choice(engines))
            else:
                k = self.rng.choice(live)
                try:
                    next(actors[k])
                except StopIteration:
                    pass
            idle = idle + 1 if self._state() == before and not self.pipe and not self.tma else 0
            if idle > 20000:
                raise Deadlock({k: (b.phase, b.pending, b.tx) for k, b in self.bar.items()})
        raise Deadlock("step limit")

    def _state(self):
        return (tuple((b.phase, b.pending, b.tx) for b in self.bar.values()), len(self.done), len(self.pipe), len(self.tma))


def self_test(seeds=40):
    """The model must REJECT broken protocols: (a) S_t(j+1) issued without waiting for the softmax to have loaded S_t(j),
    (b) the K slot released after the first tile's S only, (c) a ring one slot too shallow for the parity scheme."""
    caught = {}
    for name in ("no_sfree_wait", "early_kempty"):
        caught[name] = 0
        for seed in range(seeds):
            sim = Sim(1, "default", 16, 2, 2, seed)
            sim.mutate = name
            try:
                sim.run()
            except (AssertionError, Deadlock):
                caught[name] += 1
    assert all(v > 0 for v in caught.values()), caught
    return caught


def check_all(seeds=200, nblk=16, verbose=True):
    for ctas, rings in ((1, (2, 2)), (2, (4, 4))):
        for order in ("default", "s_first", "alias"):
            if ctas == 2 and order == "s_first":
                continue                                            # not built for the pair kernel
            for seed in range(seeds):
                Sim(ctas, order, nblk, rings[0], rings[1], seed).run()
            if verbose:
                print(f"ctas={ctas} order={order:8s} rings K/V={rings}: {seeds} random interleavings ok "
                      f"(no deadlock, no parity aliasing, every buffer consumed at the right version)")


if __name__ == "__main__":
    print("self-test (broken protocols rejected in x of 40 interleavings):", self_test())
    check_all(int(sys.argv[1]) if len(sys.argv) > 1 else 200)
