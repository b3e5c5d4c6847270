#!/usr/bin/env python
"""Randomised simulation of block2_kernel's mbarrier protocol (csrc/dce_tc_block2.cuh), with and without clusters.

No GPU.  Every warp role of every CTA of one cluster is a coroutine that executes exactly the waits, arrives, bulk
copies and tcgen05.commits the kernel executes, in the kernel's order; a scheduler interleaves them at random, bulk
copies and MMA completions land at random later times (MMAs of one CTA retire in issue order).  Checked on every run:
  * nobody deadlocks (all roles of all CTAs run to completion),
  * every parity wait is unambiguous: when a role waits for phase n of a barrier, the barrier is in phase n or n + 1
    (a barrier that could run two phases ahead of a waiter would alias under mbarrier.try_wait.parity),
  * a ring slot / slabA / slabB is never overwritten while an issued MMA has not retired, and never read by an MMA
    before the data it expects has landed,
  * with clusters, a weight block is fetched exactly once per cluster and lands in every CTA.
This pins the synchronisation design (in particular the multicast variant, `block2_cluster`), not the CUDA text.
    python tools/simulate_block2_protocol.py [runs]"""
import random
import sys

RING = 4


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.tx, self.phase = name, count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, n=1, tx=0):
        assert self.pending >= n, f"{self.name}: more arrivals than the barrier's count in phase {self.phase}"
        self.pending -= n
        self.tx += tx
        self._check()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def ready(self, n, parity):
        """A wait for phase n, expressed in the kernel as parity `parity`: True = passes now."""
        assert (n & 1) == parity, f"{self.name}: the kernel's parity expression does not match phase {n}"
        assert self.phase <= n + 1, f"{self.name}: barrier ran to phase {self.phase} while a waiter still waits for {n}"
        if n < 0:
            return True
        assert self.phase >= n, f"{self.name}: wait for phase {n} would pass early (barrier only in phase {self.phase})"
        return self.phase > n


class Cta:
    def __init__(self, rank, cl):
        self.rank = rank
        b = lambda name, c: Barrier(f"cta{rank}.{name}", c)
        self.wfull = [b(f"wfull{i}", 1) for i in range(RING)]
        self.wempty = [b(f"wempty{i}", cl) for i in range(RING)]
        self.a_full, self.a_empty = b("a_full", 1), b("a_empty", 1)
        self.d3_full = [b(f"d3_full{i}", 1) for i in range(2)]
        self.d3_empty = [b(f"d3_empty{i}", 8) for i in range(2)]
        self.d4_full = [b(f"d4_full{i}", 1) for i in range(2)]
        self.d4_empty = [b(f"d4_empty{i}", 8) for i in range(2)]
        self.x3_full, self.x3_empty = b("x3_full", 256), b("x3_empty", 1)
        self.slot_block = [None] * RING          # which block `it` a ring slot holds
        self.slot_readers = [0] * RING           # issued-but-not-retired MMA groups reading the slot
        self.slabA_tile, self.slabA_readers = None, 0
        self.slabB_tile, self.slabB_readers = None, 0
        self.pipe = []                           # in-order tensor pipe: callbacks run when the MMAs before them retire


def simulate(cl, rounds, real_tiles, seed):
    """cl CTAs of one cluster (cl = 1: the cluster-less kernel); CTA r has real_tiles[r] <= rounds real tiles."""
    rng = random.Random(seed)
    ctas = [Cta(r, cl) for r in range(cl)]
    inflight = []                                # asynchronous events: bulk copies that have not landed yet
    fetched = {}                                 # block -> number of L2 fetches

    def producer(c):
        it = 0
        for k in range(-1, rounds):              # c3(0); then for every k: c3(k + 1) if any, c4(k)
            seq = ([("w3", 4)] if k + 1 < rounds else []) + ([("w4", 8)] if k >= 0 else [])
            for _, nblk in seq:
                for _s in range(nblk):
                    slot, use = it % RING, it // RING
                    while not c.wempty[slot].ready(use - 1, ((use & 1) ^ 1)):
                        yield
                    c.wfull[slot].arrive(1, tx=24576)
                    if it % cl == c.rank:
                        fetched[it] = fetched.get(it, 0) + 1
                        for dst in ctas:
                            def land(dst=dst, slot=slot, it=it):
                                assert dst.slot_readers[slot] == 0, f"block {it} lands in a slot an MMA still reads (cta{dst.rank})"
                                dst.slot_block[slot] = it
                                dst.wfull[slot].complete_tx(24576)
                            inflight.append(land)
                    it += 1
                    yield

    def loader(c):
        for k in range(rounds):
            while not c.a_empty.ready(k - 1, (k & 1) ^ 1):
                yield
            if cl > 1 and k >= real_tiles[c.rank]:
                c.a_full.arrive(1)
            else:
                c.a_full.arrive(1, tx=33280)
                def land(c=c, k=k):
                    assert c.slabA_readers == 0, "slabA overwritten while conv3 still reads it"
                    c.slabA_tile = k
                    c.a_full.complete_tx(33280)
                inflight.append(land)
            yield

    def issuer(c):
        state = {"it": 0}
        total = rounds * 12

        def stage(slab):
            it = state["it"]
            slot = it % RING
            assert c.slot_block[slot] == it, f"cta{c.rank}: MMAs of block {it} issued but slot holds {c.slot_block[slot]}"
            c.slot_readers[slot] += 1
            if slab == "A":
                c.slabA_readers += 1
            else:
                c.slabB_readers += 1
            yield                                        # taps 0, 1 issued; now the mid-block probe of the next block
            if it + 1 < total:
                nslot, nuse = (it + 1) % RING, (it + 1) // RING
                while not c.wfull[nslot].ready(nuse, nuse & 1):
                    yield
            def retire(slot=slot, slab=slab):            # tcgen05.commit(wempty[slot]) [multicast with clusters]
                c.slot_readers[slot] -= 1
                if slab == "A":
                    c.slabA_readers -= 1
                else:
                    c.slabB_readers -= 1
                for dst in (ctas if cl > 1 else [c]):
                    dst.wempty[slot].arrive(1)
            c.pipe.append(retire)
            state["it"] += 1

        def issue_c3(k):
            buf, ph = k & 1, (k >> 1) & 1
            while not c.a_full.ready(k, k & 1):
                yield
            while not c.d3_empty[buf].ready((k >> 1) - 1, ph ^ 1):
                yield
            if not (cl > 1 and k >= real_tiles[c.rank]):
                assert c.slabA_tile == k, f"conv3({k}) reads slabA holding tile {c.slabA_tile}"
            for _s in range(4):
                yield from stage("A")
            c.pipe.append(lambda: c.a_empty.arrive(1))
            c.pipe.append(lambda buf=buf: c.d3_full[buf].arrive(1))

        def issue_c4(k):
            buf, ph = k & 1, (k >> 1) & 1
            while not c.x3_full.ready(k, k & 1):
                yield
            while not c.d4_empty[buf].ready((k >> 1) - 1, ph ^ 1):
                yield
            assert c.slabB_tile == k, f"conv4({k}) reads slabB holding tile {c.slabB_tile}"
            for _s in range(8):
                yield from stage("B")
            c.pipe.append(lambda: c.x3_empty.arrive(1))
            c.pipe.append(lambda buf=buf: c.d4_full[buf].arrive(1))

        if rounds > 0:
            while not c.wfull[0].ready(0, 0):
                yield
            yield from issue_c3(0)
        for k in range(rounds):
            if k + 1 < rounds:
                yield from issue_c3(k + 1)
            yield from issue_c4(k)

    def epilogue(c):
        def epi1(k):
            buf, ph = k & 1, (k >> 1) & 1
            while not c.d3_full[buf].ready(k >> 1, ph):
                yield
            c.d3_empty[buf].arrive(8)
            while not c.x3_empty.ready(k - 1, (k & 1) ^ 1):
                yield
            assert c.slabB_readers == 0, "slabB overwritten while conv4 still reads it"
            c.slabB_tile = k
            c.x3_full.arrive(256)
            yield

        def epi2(k):
            buf, ph = k & 1, (k >> 1) & 1
            while not c.d4_full[buf].ready(k >> 1, ph):
                yield
            c.d4_empty[buf].arrive(8)
            yield

        for k in range(rounds):
            yield from epi1(k)
            if k > 0:
                yield from epi2(k - 1)
        if rounds > 0:
            yield from epi2(rounds - 1)

    roles = []
    for c in ctas:
        roles += [producer(c), loader(c), issuer(c), epilogue(c)]
    alive = list(roles)
    idle_steps = 0
    while alive:
        progressed = False
        # asynchronous completions first or later, at random
        if inflight and rng.random() < 0.5:
            inflight.pop(rng.randrange(len(inflight)))()
            progressed = True
        for c in ctas:
            if c.pipe and rng.random() < 0.5:
                c.pipe.pop(0)()                          # MMAs of a CTA retire in issue order
                progressed = True
        r = rng.choice(alive)
        before = snapshot(ctas)
        try:
            next(r)
        except StopIteration:
            alive.remove(r)
            progressed = True
        if snapshot(ctas) != before:
            progressed = True
        idle_steps = 0 if progressed else idle_steps + 1
        if idle_steps > 20000 and not inflight and not any(c.pipe for c in ctas):
            raise AssertionError(f"deadlock: {len(alive)} roles still waiting (cl={cl}, rounds={rounds}, real={real_tiles}, seed={seed})")
    while inflight:
        inflight.pop()()
    for c in ctas:
        while c.pipe:
            c.pipe.pop(0)()
    total_blocks = rounds * 12
    assert all(fetched.get(i, 0) == 1 for i in range(total_blocks)), "a block was not fetched exactly once per cluster"
    return True


def simulate_tapgemm(cl, mt, nstage, stages, tiles, seed, ksa=4, nbuf=2, top_of_tile_wait=False):
    """tapgemm_kernel (csrc/dce_tc.cuh): 4 producer warps, MT issuers, 8 epilogue warps per CTA; cl = 2: a CTA pair with
    one tile each that shares the activation slabs by multicast (option "fc_cluster"), cl = 1: the plain kernel with
    `tiles` tiles per CTA.  nbuf = accumulator buffers in TMEM (1 for fc.0, whose two 256-column accumulators fill it).
    Same checks as `simulate`.

    Found with this simulation: with nbuf = 1 and MORE THAN ONE tile per CTA the kernel deadlocks — the issuer probes the
    next tile's `tempty` in the middle of the current tile's last stage, i.e. before it has committed `tfull` for the
    buffer the epilogue must drain first.  fc.0 is the only nbuf = 1 layer and always runs one tile per CTA on a B200
    (128 tiles, 148 SMs); launch_layer() now refuses the combination instead of hanging on a smaller device.
    top_of_tile_wait = True is the repaired order the F8 issue loop uses: with one buffer, `tempty` is awaited at the
    top of the next tile."""
    rng = random.Random(seed)
    ncta = cl
    NA, SLABB, BB = mt * 2 * ksa, 2080, 32768

    class T:
        pass
    ctas = []
    for r in range(ncta):
        c = T()
        c.rank = r
        c.full = [Barrier(f"cta{r}.full{i}", 4) for i in range(nstage)]
        c.empty = [Barrier(f"cta{r}.empty{i}", mt * cl) for i in range(nstage)]
        c.tfull = [Barrier(f"cta{r}.tfull{i}", mt) for i in range(2)]
        c.tempty = [Barrier(f"cta{r}.tempty{i}", 8) for i in range(2)]
        c.slot_stage = [[None] * (NA + 1) for _ in range(nstage)]       # which stage `it` each piece of a slot holds
        c.slot_readers = [0] * nstage
        c.pipes = [[] for _ in range(mt)]                               # MMAs retire in order per issuer (conservative)
        ctas.append(c)
    inflight = []
    fetched = {}

    def producer(c, pw):
        my = [cc for cc in range(pw, NA + 1, 4)]
        my_bytes = sum(SLABB if cc < NA else BB for cc in my)
        it = 0
        for _t in range(tiles):
            for _s in range(stages):
                slot, use = it % nstage, it // nstage
                while not c.empty[slot].ready(use - 1, (use & 1) ^ 1):
                    yield
                c.full[slot].arrive(1, tx=my_bytes)
                for cc in my:
                    if cc < NA and cl > 1:
                        if (cc & 1) != c.rank:
                            continue
                        dsts = ctas
                    else:
                        dsts = [c]
                    if cc < NA:
                        fetched[(it, cc)] = fetched.get((it, cc), 0) + 1
                    for dst in dsts:
                        def land(dst=dst, slot=slot, cc=cc, it=it):
                            assert dst.slot_readers[slot] == 0, f"stage {it} piece {cc} lands in a slot an MMA still reads (cta{dst.rank})"
                            dst.slot_stage[slot][cc] = it
                            dst.full[slot].complete_tx(SLABB if cc < NA else BB)
                        inflight.append(land)
                it += 1
                yield

    def issuer(c, m):
        total = tiles * stages
        it = 0
        if tiles > 0:
            while not c.tempty[0].ready(-1, 1):
                yield
            while not c.full[0].ready(0, 0):
                yield
        for tcount in range(tiles):
            buf = tcount % nbuf
            for s in range(stages):
                slot = it % nstage
                if top_of_tile_wait and nbuf == 1 and s == 0 and tcount > 0:
                    while not c.tempty[0].ready(tcount - 1, (tcount & 1) ^ 1):
                        yield
                assert all(v == it for v in c.slot_stage[slot]), f"cta{c.rank}: MMAs of stage {it} issued on {c.slot_stage[slot]}"
                c.slot_readers[slot] += 1
                yield
                if it + 1 < total:
                    if s == stages - 1 and not (top_of_tile_wait and nbuf == 1):
                        nt = tcount + 1
                        while not c.tempty[nt % nbuf].ready(nt // nbuf - 1, ((nt // nbuf) & 1) ^ 1):
                            yield
                    nuse = (it + 1) // nstage
                    while not c.full[(it + 1) % nstage].ready(nuse, nuse & 1):
                        yield
                def retire(slot=slot):
                    c.slot_readers[slot] -= 1
                    for dst in (ctas if cl > 1 else [c]):
                        dst.empty[slot].arrive(1)
                c.pipes[m].append(retire)
                if s == stages - 1:
                    c.pipes[m].append(lambda buf=buf: c.tfull[buf].arrive(1))
                it += 1

    def epilogue(c):
        for tcount in range(tiles):
            buf, tph = tcount % nbuf, (tcount // nbuf) & 1
            while not c.tfull[buf].ready(tcount // nbuf, tph):
                yield
            yield
            c.tempty[buf].arrive(8)

    alive = []
    for c in ctas:
        alive += [producer(c, pw) for pw in range(4)] + [issuer(c, m) for m in range(mt)] + [epilogue(c)]
    idle = 0
    while alive:
        progressed = False
        if inflight and rng.random() < 0.5:
            inflight.pop(rng.randrange(len(inflight)))()
            progressed = True
        for c in ctas:
            for pipe in c.pipes:
                if pipe and rng.random() < 0.5:
                    pipe.pop(0)()
                    progressed = True
        r = rng.choice(alive)
        before = [(b.phase, b.pending, b.tx) for c in ctas for b in c.full + c.empty + c.tfull + c.tempty]
        try:
            next(r)
        except StopIteration:
            alive.remove(r)
            progressed = True
        if before != [(b.phase, b.pending, b.tx) for c in ctas for b in c.full + c.empty + c.tfull + c.tempty]:
            progressed = True
        idle = 0 if progressed else idle + 1
        if idle > 20000 and not inflight and not any(p for c in ctas for p in c.pipes):
            raise AssertionError(f"deadlock: {len(alive)} roles waiting (cl={cl}, mt={mt}, nstage={nstage}, stages={stages}, tiles={tiles}, seed={seed})")
    assert all(v == 1 for v in fetched.values()), "an activation slab was not fetched exactly once per pair"
    return True


def check_tapgemm(runs=40, seed=0):
    rng = random.Random(seed)
    for _ in range(runs):
        cl = rng.choice([1, 2])
        mt, nstage, nbuf = rng.choice([(2, 3, 1), (1, 6, 2), (2, 3, 2), (2, 4, 2)])     # fc.0, fc.3, conv3, conv4 configurations
        fixed = nbuf == 1 and cl == 1 and rng.random() < 0.5           # the F8 loop's order: any number of tiles with one buffer
        tiles = rng.choice([1, 2, 3, 4]) if (fixed or (cl == 1 and nbuf > 1)) else 1
        simulate_tapgemm(cl, mt, nstage, rng.choice([4, 7, 12]), tiles, rng.randrange(1 << 30), nbuf=nbuf, top_of_tile_wait=fixed)
    return runs


def snapshot(ctas):
    out = []
    for c in ctas:
        for b in c.wfull + c.wempty + [c.a_full, c.a_empty, c.x3_full, c.x3_empty] + c.d3_full + c.d3_empty + c.d4_full + c.d4_empty:
            out.append((b.phase, b.pending, b.tx))
        out.append((tuple(c.slot_block), tuple(c.slot_readers), len(c.pipe)))
    return out


def check(runs=60, seed=0):
    rng = random.Random(seed)
    n = 0
    for _ in range(runs):
        cl = rng.choice([1, 2, 4])
        rounds = rng.choice([1, 2, 3, 5])
        # with clusters, some CTAs may have one real tile less than `rounds` (dummy last round)
        real = [rounds - (1 if cl > 1 and rng.random() < 0.4 else 0) for _ in range(cl)]
        simulate(cl, rounds, real, rng.randrange(1 << 30))
        n += 1
    return n


if __name__ == "__main__":
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    print(f"tapgemm (fc.0 / fc.3, plain and as CTA pairs): {check_tapgemm(runs // 2)} randomised runs passed")
    print(f"{check(runs)} randomised runs (clusters of 1 / 2 / 4, 1-5 rounds, dummy last rounds): no deadlock, no phase aliasing, "
          f"no operand overwritten under a pending MMA, every weight block fetched once per cluster")
