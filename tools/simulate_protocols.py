#!/usr/bin/env python
"""Randomised discrete-event simulation of the mbarrier protocols of the two fused convolution kernels
(csrc/dce_tc_block2.cuh, csrc/dce_tc_block1.cuh) and of tapgemm_kernel (csrc/dce_tc.cuh: fc.0 with two accumulators in one
TMEM buffer, fc.3 with two issuers on alternating K-stages, the layer-wise convolutions): every warp role is a coroutine that waits on / arrives at mbarriers
exactly as the kernel does, the tensor pipe is an in-order queue whose tcgen05.commit entries arrive when everything
before them has retired, bulk copies land after a random latency.  Durations are drawn at random, so many interleavings
are explored.  Checked on every run:

  * no deadlock: every role finishes all its tiles;
  * no parity aliasing: a wait that passes has seen exactly the completion it was written for (an mbarrier wait only
    knows the parity of the phase: a barrier that runs two completions ahead looks "not yet complete");
  * no operand hazard: nothing writes a buffer (slab, ring slot, accumulator, staging tile) while an MMA that reads it,
    an epilogue that loads it or a copy that drains it is still in flight, and no MMA reads a buffer that is being written.

`mutations` removes one wait at a time: the simulation must then report a violation (tests/test_properties_cpu.py), which
is what shows the checks can fail.  This is test infrastructure; it models the protocol, not the arithmetic.

    python tools/simulate_protocols.py [runs]
"""
import heapq
import random
import sys


class Violation(Exception):
    pass


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completions = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending < 0:
            raise Violation(f"{self.name}: more arrivals than its count in one phase")
        if self.pending == 0:
            self.completions += 1
            self.pending = self.count


class Res:
    """A shared-memory buffer or a TMEM accumulator: who is reading it (MMAs in flight, epilogue loads, outgoing
    copies) and whether someone is writing it."""

    def __init__(self, name):
        self.name, self.readers, self.writers, self.mma_writers = name, 0, 0, 0
        self.version = None            # what the buffer holds: set by its writers, checked by readers that say what they expect

    def stamp(self, tag):
        if tag is not None:
            self.version = tag

    def expect(self, tag, who):
        if tag is not None and self.version != tag:
            raise Violation(f"{who}: reads {self.name} expecting {tag}, it holds {self.version}")

    def begin_read(self):
        if self.writers or self.mma_writers:
            raise Violation(f"{self.name}: read while it is being written")
        self.readers += 1

    def begin_mma_write(self):          # MMAs accumulating into the same TMEM columns are ordered by the pipe itself
        if self.readers or self.writers:
            raise Violation(f"{self.name}: an MMA accumulates into it while it is {'read' if self.readers else 'written'} by a warp")
        self.mma_writers += 1

    def end_mma_write(self):
        self.mma_writers -= 1

    def end_read(self):
        self.readers -= 1

    def begin_write(self):
        if self.readers or self.writers or self.mma_writers:
            raise Violation(f"{self.name}: written while {'read' if self.readers else 'written'} by someone else")
        self.writers += 1

    def end_write(self):
        self.writers -= 1


def _pairs(items):
    """[res | (res, tag)] -> [(res, tag)]"""
    return [it if isinstance(it, tuple) else (it, None) for it in items]


class Sim:
    def __init__(self, rng, pipe_depth=3):
        self.rng, self.now, self.heap, self.seq = rng, 0.0, [], 0
        self.pipe, self.pipe_busy, self.pipe_depth = [], False, pipe_depth
        self.agents, self.blocked, self.done = {}, {}, set()

    # ---- scheduling ------------------------------------------------------------------------------------------
    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.heap, (t, self.seq, fn))

    def spawn(self, name, gen):
        self.agents[name] = gen
        self.at(self.now, lambda n=name: self.step(n, None))

    def step(self, name, value):
        gen = self.agents[name]
        try:
            op = gen.send(value)
        except StopIteration:
            self.done.add(name)
            return
        kind = op[0]
        if kind == "delay":
            self.at(self.now + op[1], lambda: self.step(name, None))
        elif kind == "wait":                     # ("wait", bar, j): the j-th completion of bar (j < 0: passes at once)
            self.blocked[name] = op
            self.poll(name)
        elif kind == "arrive":
            op[1].arrive()
            self.wake()
            self.at(self.now, lambda: self.step(name, None))
        elif kind == "mma":                      # ("mma", duration, reads, writes) | ("commit", bars)
            self.blocked[name] = op
            self.poll(name)
        elif kind == "commit":
            self.pipe.append(("commit", op[1]))
            self.pump()
            self.at(self.now, lambda: self.step(name, None))
        elif kind == "copy":                     # ("copy", latency, dst_res or None, src_res or None, bar or None)
            _, lat, dst, src, bar = op
            (dst, dtag), (src, stag) = _pairs([dst])[0], _pairs([src])[0]
            if dst: dst.begin_write(); dst.stamp(dtag)
            if src: src.begin_read(); src.expect(stag, name)

            def land(dst=dst, src=src, bar=bar):
                if dst: dst.end_write()
                if src: src.end_read()
                if bar: bar.arrive()
                self.wake()
            self.at(self.now + lat, land)
            self.at(self.now, lambda: self.step(name, None))
        elif kind == "copy_wait":                # like copy, but the agent resumes when it has landed (wait_group)
            _, lat, dst, src = op
            (dst, dtag), (src, stag) = _pairs([dst])[0], _pairs([src])[0]
            if dst: dst.begin_write(); dst.stamp(dtag)
            if src: src.begin_read(); src.expect(stag, name)

            def land2(dst=dst, src=src):
                if dst: dst.end_write()
                if src: src.end_read()
                self.step(name, None)
            self.at(self.now + lat, land2)
        elif kind == "rw":                       # ("rw", duration, reads, writes): generic-proxy access by a warp
            _, dur, reads, writes = op
            reads, writes = _pairs(reads), _pairs(writes)
            for r, t in reads: r.begin_read(); r.expect(t, name)
            for w, t in writes: w.begin_write(); w.stamp(t)

            def fin(reads=reads, writes=writes):
                for r, _ in reads: r.end_read()
                for w, _ in writes: w.end_write()
                self.step(name, None)
            self.at(self.now + dur, fin)
        else:
            raise ValueError(op)

    def poll(self, name):
        op = self.blocked.get(name)
        if op is None:
            return
        if op[0] == "wait":
            bar, j = op[1], op[2]
            if j < 0 or (bar.completions & 1) != (j & 1):
                if j >= 0 and bar.completions != j + 1:
                    raise Violation(f"{name}: wait on {bar.name} for completion {j} passed at completion count {bar.completions} (parity aliasing)")
                del self.blocked[name]
                self.at(self.now, lambda: self.step(name, None))
        elif op[0] == "mma":
            if len(self.pipe) < self.pipe_depth:
                del self.blocked[name]
                _, dur, reads, writes = op
                reads, writes = _pairs(reads), _pairs(writes)
                for r, t in reads: r.begin_read(); r.expect(t, name)
                for w, t in writes: w.begin_mma_write(); w.stamp(t)
                self.pipe.append(("mma", dur, reads, writes))
                self.pump()
                self.at(self.now, lambda: self.step(name, None))

    def wake(self):
        for n in list(self.blocked):
            self.poll(n)

    def pump(self):
        if self.pipe_busy or not self.pipe:
            return
        head = self.pipe[0]
        if head[0] == "commit":
            self.pipe.pop(0)
            for b in head[1]: b.arrive()
            self.wake()
            self.pump()
            return
        self.pipe_busy = True

        def retire():
            _, _, reads, writes = self.pipe.pop(0)
            for r, _ in reads: r.end_read()
            for w, _ in writes: w.end_mma_write()
            self.pipe_busy = False
            self.wake()
            self.pump()
        self.at(self.now + head[1], retire)

    def run(self, limit=10_000_000):
        n = 0
        while self.heap:
            t, _, fn = heapq.heappop(self.heap)
            self.now = t
            fn()
            n += 1
            if n > limit:
                raise Violation("event limit")
        missing = set(self.agents) - self.done
        if missing:
            raise Violation("deadlock: " + ", ".join(f"{m} at {self.blocked.get(m, ('?',))[0]} {getattr(self.blocked.get(m, (0, None))[1], 'name', '')}" for m in sorted(missing)))


# ---------------------------------------------------------------------------------------------------------------
# block2_kernel: weight ring (4 slots), slabA, slabB, single-buffered D3 / D4, staging tile
# ---------------------------------------------------------------------------------------------------------------
def simulate_block2(seed, tiles=7, mutate=None, epi_warps=4):
    rng = random.Random(seed)
    sim = Sim(rng)
    R = lambda lo, hi: rng.uniform(lo, hi)
    ring = [Res(f"ring{i}") for i in range(4)]
    slabA, slabB, D3, D4, stage = Res("slabA"), Res("slabB"), Res("D3"), Res("D4"), Res("staging")
    wfull = [Bar(f"wfull{i}", 1) for i in range(4)]
    wempty = [Bar(f"wempty{i}", 1) for i in range(4)]
    a_full, a_empty = Bar("a_full", 1), Bar("a_empty", 1)
    d3_full, d3_empty = Bar("d3_full", 1), Bar("d3_empty", epi_warps)
    x3_full, x3_empty = Bar("x3_full", epi_warps), Bar("x3_empty", 1)
    d4_full, d4_empty = Bar("d4_full", 1), Bar("d4_empty", epi_warps)
    st_full, st_empty = Bar("st_full", epi_warps), Bar("st_empty", 1)
    skip = lambda tag: mutate == tag

    def producer():                      # blocks in the order the issuer consumes them: c3(0); then c3(k+1), c4(k)
        total = tiles * 12
        for it in range(total):
            slot = it % 4
            if not skip("wempty"):
                yield ("wait", wempty[slot], it // 4 - 1)
            yield ("copy", R(50, 400), (ring[slot], it), None, wfull[slot])
            yield ("delay", R(1, 20))

    def loader():
        for k in range(tiles):
            if not skip("a_empty"):
                yield ("wait", a_empty, k - 1)
            yield ("copy", R(50, 600), (slabA, k), None, a_full)

    def issuer():
        it = 0

        def stages(n, slab, acc, k):
            nonlocal it
            for s in range(n):
                slot = it % 4
                yield ("wait", wfull[slot], it // 4)
                for _ in range(6):
                    yield ("mma", R(8, 16), [(slab, k), (ring[slot], it)], [(acc, k)])
                yield ("commit", [wempty[slot]])
                it += 1

        def c3(k):
            yield ("wait", a_full, k)
            if not skip("d3_empty"):
                yield ("wait", d3_empty, k - 1)
            yield from stages(4, slabA, D3, k)
            yield ("commit", [a_empty, d3_full])

        def c4(k):
            yield ("wait", x3_full, k)
            if not skip("d4_empty"):
                yield ("wait", d4_empty, k - 1)
            yield from stages(8, slabB, D4, k)
            yield ("commit", [x3_empty, d4_full])

        if tiles:
            yield from c3(0)
        for k in range(tiles):
            if k + 1 < tiles:
                yield from c3(k + 1)
            yield from c4(k)

    def epilogue(w):
        def e1(k):
            yield ("wait", d3_full, k)
            yield ("rw", R(5, 30), [(D3, k)], [])        # tcgen05.ld of the whole accumulator
            yield ("arrive", d3_empty)
            yield ("delay", R(20, 120))                  # bias / ReLU / split
            if not skip("x3_empty"):
                yield ("wait", x3_empty, k - 1)
            yield ("rw", R(10, 80), [], [] if w else [(slabB, k)])   # the warps write disjoint rows: one of them stands for the buffer
            yield ("arrive", x3_full)

        def e2(k):
            yield ("wait", d4_full, k)
            yield ("rw", R(5, 30), [(D4, k)], [])
            yield ("arrive", d4_empty)
            yield ("delay", R(20, 200))
            if not skip("st_empty"):
                yield ("wait", st_empty, k - 1)
            yield ("rw", R(10, 60), [], [] if w else [(stage, k)])
            yield ("arrive", st_full)

        for k in range(tiles):
            yield from e1(k)
            if k > 0:
                yield from e2(k - 1)
        if tiles:
            yield from e2(tiles - 1)

    def store():
        for k in range(tiles):
            yield ("wait", st_full, k)
            yield ("copy_wait", R(20, 500) if rng.random() < 0.8 else R(500, 6000), None, (stage, k))   # bulk copies + wait_group.read
            yield ("arrive", st_empty)

    sim.spawn("producer", producer()); sim.spawn("loader", loader()); sim.spawn("issuer", issuer()); sim.spawn("store", store())
    for w in range(epi_warps):
        sim.spawn(f"epi{w}", epilogue(w))
    sim.run()
    return sim.now


# ---------------------------------------------------------------------------------------------------------------
# block1_kernel: a CTA PAIR.  Per CTA: slab0[2] (raw rows -> conv1 operand), slab1[2], D1[2], D2[2], staging tile; the leader's
# two issuer warps issue M = 256 MMAs that read both CTAs' slabs and write both CTAs' accumulators; commits are multicast
# ---------------------------------------------------------------------------------------------------------------
def simulate_block1(seed, tiles=7, mutate=None, epi_warps=2):
    rng = random.Random(seed)
    sim = Sim(rng, pipe_depth=4)
    R = lambda lo, hi: rng.uniform(lo, hi)
    skip = lambda tag: mutate == tag

    class Cta:
        def __init__(self, r):
            n = lambda s: f"cta{r}.{s}"
            self.slab0 = [Res(n(f"slab0[{i}]")) for i in range(2)]
            self.slab1 = [Res(n(f"slab1[{i}]")) for i in range(2)]
            self.D1 = [Res(n(f"D1[{i}]")) for i in range(2)]
            self.D2 = [Res(n(f"D2[{i}]")) for i in range(2)]
            self.stage = Res(n("staging"))
            self.raw_full = [Bar(n(f"raw_full{i}"), 1) for i in range(2)]
            self.x0_full = [Bar(n(f"x0_full{i}"), 1) for i in range(2)]
            self.x0_empty = [Bar(n(f"x0_empty{i}"), 1) for i in range(2)]
            self.d1_full = [Bar(n(f"d1_full{i}"), 1) for i in range(2)]
            self.d1_empty = [Bar(n(f"d1_empty{i}"), epi_warps) for i in range(2)]
            self.x1_full = [Bar(n(f"x1_full{i}"), epi_warps) for i in range(2)]
            self.x1_empty = [Bar(n(f"x1_empty{i}"), 1) for i in range(2)]
            self.d2_full = [Bar(n(f"d2_full{i}"), 1) for i in range(2)]
            self.d2_empty = [Bar(n(f"d2_empty{i}"), epi_warps) for i in range(2)]
            self.st_full, self.st_done = Bar(n("st_full"), epi_warps), Bar(n("st_done"), 1)
            self.p1_ready = [Bar(n(f"p1_ready{i}"), 1) for i in range(2)]      # used in the leader only
            self.p2_ready = [Bar(n(f"p2_ready{i}"), 1) for i in range(2)]

    C = [Cta(0), Cta(1)]
    lead = C[0]

    def loader(c):
        for k in range(tiles):
            b = k & 1
            if not skip("x0_empty"):
                yield ("wait", c.x0_empty[b], k // 2 - 1)
            yield ("copy", R(30, 300), (c.slab0[b], ("raw", k)), None, c.raw_full[b])

    def converter(c):
        for k in range(tiles):
            b = k & 1
            yield ("wait", c.raw_full[b], k // 2)
            yield ("rw", R(40, 200), [(c.slab0[b], ("raw", k))], [])     # the raw rows are read ...
            yield ("rw", R(20, 100), [], [(c.slab0[b], k)])              # ... and the operand image written in place
            yield ("arrive", c.x0_full[b])

    def relay(c, conv):                   # the peer's issuer warps
        for k in range(tiles):
            b, j = k & 1, k // 2
            if conv == 1:
                yield ("wait", c.x0_full[b], j); yield ("wait", c.d1_empty[b], j - 1)
                yield ("arrive", lead.p1_ready[b])
            else:
                yield ("wait", c.x1_full[b], j); yield ("wait", c.d2_empty[b], j - 1)
                yield ("arrive", lead.p2_ready[b])

    def issuer(conv):                     # the leader's issuer warps
        for k in range(tiles):
            b, j = k & 1, k // 2
            if conv == 1:
                yield ("wait", lead.x0_full[b], j)
                if not skip("d1_empty"):
                    yield ("wait", lead.d1_empty[b], j - 1)
                if not skip("p1_ready"):
                    yield ("wait", lead.p1_ready[b], j)
                for _ in range(24):
                    yield ("mma", R(4, 9), [(C[0].slab0[b], k), (C[1].slab0[b], k)], [(C[0].D1[b], k), (C[1].D1[b], k)])
                yield ("commit", [C[0].x0_empty[b], C[1].x0_empty[b], C[0].d1_full[b], C[1].d1_full[b]])
            else:
                yield ("wait", lead.x1_full[b], j)
                if not skip("d2_empty"):
                    yield ("wait", lead.d2_empty[b], j - 1)
                if not skip("p2_ready"):
                    yield ("wait", lead.p2_ready[b], j)
                for _ in range(24):
                    yield ("mma", R(4, 9), [(C[0].slab1[b], k), (C[1].slab1[b], k)], [(C[0].D2[b], k), (C[1].D2[b], k)])
                yield ("commit", [C[0].x1_empty[b], C[1].x1_empty[b], C[0].d2_full[b], C[1].d2_full[b]])

    def epilogue(c, w):
        def e1(k):
            b, j = k & 1, k // 2
            yield ("wait", c.d1_full[b], j)
            yield ("rw", R(5, 25), [(c.D1[b], k)], [])
            yield ("arrive", c.d1_empty[b])
            yield ("delay", R(20, 100))
            if not skip("x1_empty"):
                yield ("wait", c.x1_empty[b], j - 1)
            yield ("rw", R(10, 60), [], [] if w else [(c.slab1[b], k)])
            yield ("arrive", c.x1_full[b])

        def e2(k):
            b, j = k & 1, k // 2
            yield ("wait", c.d2_full[b], j)
            yield ("rw", R(5, 25), [(c.D2[b], k)], [])
            yield ("arrive", c.d2_empty[b])
            yield ("delay", R(20, 150))
            if not skip("st_done"):
                yield ("wait", c.st_done, k - 1)
            yield ("rw", R(5, 40), [], [] if w else [(c.stage, k)])
            yield ("arrive", c.st_full)

        for k in range(tiles):
            yield from e1(k)
            if k > 0:
                yield from e2(k - 1)
        if tiles:
            yield from e2(tiles - 1)

    def store(c):
        for k in range(tiles):
            yield ("wait", c.st_full, k)
            yield ("copy_wait", R(20, 400) if rng.random() < 0.8 else R(400, 3000), None, (c.stage, k))
            yield ("arrive", c.st_done)

    for r, c in enumerate(C):
        sim.spawn(f"loader{r}", loader(c)); sim.spawn(f"converter{r}", converter(c)); sim.spawn(f"store{r}", store(c))
        for w in range(epi_warps):
            sim.spawn(f"epi{r}.{w}", epilogue(c, w))
    sim.spawn("issuer1", issuer(1)); sim.spawn("issuer2", issuer(2))
    sim.spawn("relay1", relay(C[1], 1)); sim.spawn("relay2", relay(C[1], 2))
    sim.run()
    return sim.now


# ---------------------------------------------------------------------------------------------------------------
# tapgemm_kernel (fc.0: MT = 2 issuers, ONE accumulator buffer; fc.3: KS = 2 issuers on alternating K-stages, two buffers):
# 4 producer warps fill a ring of NSTAGE slots, the issuers probe the NEXT stage's barrier (and, with two buffers, the next
# tile's accumulator) in the middle of the current stage, i.e. before they commit it; with one buffer the accumulator is
# awaited at the top of the tile (the order that does not deadlock: DESIGN.md, round 1)
# ---------------------------------------------------------------------------------------------------------------
def simulate_tapgemm(seed, tiles=3, MT=1, KS=1, NBUF=2, NSTAGE=4, stages=6, mutate=None, epi_warps=4, prod_warps=2):
    rng = random.Random(seed)
    sim = Sim(rng, pipe_depth=3)
    R = lambda lo, hi: rng.uniform(lo, hi)
    skip = lambda tag: mutate == tag
    ACCS = MT * KS
    ring = [[Res(f"ring{i}.{p}") for p in range(prod_warps)] for i in range(NSTAGE)]
    acc = [[Res(f"acc{b}.{i}") for i in range(ACCS)] for b in range(2)]
    full = [Bar(f"full{i}", prod_warps) for i in range(NSTAGE)]
    empty = [Bar(f"empty{i}", MT) for i in range(NSTAGE)]
    tfull = [Bar(f"tfull{b}", ACCS) for b in range(2)]
    tempty = [Bar(f"tempty{b}", epi_warps) for b in range(2)]
    total = tiles * stages

    def producer(pw):
        for it in range(total):
            slot = it % NSTAGE
            if not skip("empty"):
                yield ("wait", empty[slot], it // NSTAGE - 1)
            yield ("copy", R(30, 400), (ring[slot][pw], it), None, full[slot])
            yield ("delay", R(1, 10))

    def issuer(iw):
        mt, ks = (0, iw) if KS > 1 else (iw, 0)
        it = ks
        if tiles:
            yield ("wait", tempty[0], -1)
            yield ("wait", full[ks % NSTAGE], 0)
        for t in range(tiles):
            buf = t % NBUF
            if NBUF == 1 and t > 0 and not skip("tempty"):
                yield ("wait", tempty[0], t - 1)
            for s_ in range(ks, stages, KS):
                slot = it % NSTAGE
                last = s_ + KS >= stages
                reads = [(r, it) for r in ring[slot]]
                yield ("mma", R(6, 14), reads, [(acc[buf][iw], t)])
                if it + KS < total:
                    if NBUF > 1 and last and not skip("tempty"):
                        nt = t + 1
                        yield ("wait", tempty[nt % NBUF], nt // NBUF - 1)
                    if not skip("full"):
                        yield ("wait", full[(it + KS) % NSTAGE], (it + KS) // NSTAGE)
                yield ("mma", R(6, 14), reads, [(acc[buf][iw], t)])
                yield ("commit", [empty[slot]] + ([tfull[buf]] if last else []))
                it += KS

    def epilogue(w):
        for t in range(tiles):
            buf = t % NBUF
            yield ("wait", tfull[buf], t // NBUF)
            yield ("rw", R(20, 300) if rng.random() < 0.7 else R(300, 4000), [(a, t) for a in acc[buf]], [])
            yield ("arrive", tempty[buf])

    for pw in range(prod_warps):
        sim.spawn(f"producer{pw}", producer(pw))
    for iw in range(ACCS):
        sim.spawn(f"issuer{iw}", issuer(iw))
    for w in range(epi_warps):
        sim.spawn(f"epi{w}", epilogue(w))
    sim.run()
    return sim.now


TAPGEMM_CONFIGS = [dict(MT=2, KS=1, NBUF=1, NSTAGE=3, stages=6),      # fc.0: two accumulators fill the TMEM, one buffer
                   dict(MT=1, KS=2, NBUF=2, NSTAGE=6, stages=8),      # fc.3: two issuers on alternating K-stages
                   dict(MT=1, KS=1, NBUF=2, NSTAGE=4, stages=5)]      # the layer-wise convolutions
TAPGEMM_MUTATIONS = ["empty", "tempty", "full"]

BLOCK2_MUTATIONS = ["wempty", "a_empty", "d3_empty", "d4_empty", "x3_empty", "st_empty"]
# (block1's waits on d2_empty and x1_empty are implied by the order in which the eight epilogue warps work — epilogue 1 of
# tile k comes after epilogue 2 of tile k-2, which waited for conv2 of tile k-2 — so removing them changes nothing
# observable; they stay in the kernel because that order is an implementation detail of another role)
BLOCK1_MUTATIONS = ["x0_empty", "d1_empty", "st_done", "p1_ready", "p2_ready"]


def check(runs=200, tiles=(1, 2, 3, 7)):
    """-> (runs executed, mutations that were NOT detected)"""
    n = 0
    for seed in range(runs):
        for t in tiles:
            simulate_block2(seed * 7 + t, tiles=t)
            simulate_block1(seed * 11 + t, tiles=t)
            n += 2
    missed = []
    for ci, cfg in enumerate(TAPGEMM_CONFIGS):
        for seed in range(runs):
            for t in tiles[:3]:
                simulate_tapgemm(seed * 13 + t, tiles=t, **cfg)
                n += 1
        for m in TAPGEMM_MUTATIONS:
            caught = False
            for seed in range(max(60, runs // 2)):
                try:
                    simulate_tapgemm(seed, tiles=4, mutate=m, **cfg)
                except Violation:
                    caught = True
                    break
            if not caught:
                missed.append(f"tapgemm{ci}:{m}")
    for name, fn, muts in (("block2", simulate_block2, BLOCK2_MUTATIONS), ("block1", simulate_block1, BLOCK1_MUTATIONS)):
        for m in muts:
            caught = False
            for seed in range(max(60, runs // 2)):
                try:
                    fn(seed, tiles=7, mutate=m)
                except Violation:
                    caught = True
                    break
            if not caught:
                missed.append(f"{name}:{m}")
    return n, missed


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    n, missed = check(runs)
    print(f"{n} randomised runs of the shipped protocols: no deadlock, no parity aliasing, no operand hazard")
    print("every removed wait is detected" if not missed else f"NOT detected when removed: {missed}")
    sys.exit(1 if missed else 0)
