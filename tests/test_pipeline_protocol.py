"""CPU model check of gemm_tn's intra-CTA pipeline protocol (srfrd_b200/csrc/gemm.cu, gemm_tn_kernel).

The kernel's warps talk through mbarriers whose waits are PARITY waits: a wait for "the phase with parity p" passes as soon
as the barrier's most recently completed phase has parity p -- also when that is the phase BEFORE the one the waiter meant
(two phases behind) or two phases after it.  DESIGN.md section 5 states the rule the kernel is built on ("no waiter ever
skips a phase"); this test transcribes every warp role's loop (producer + tile-id ring, one or two MMA issuers, aux
producer, two epilogue sets with 2 or 4 tile buffers and the deferred store-read wait) into coroutines over simulated
barriers and runs them under random interleavings and random completion delays of the asynchronous operations
(TMA loads, tcgen05.commit arrivals, TMA store reads).  It asserts
  * every wait passes on exactly the phase it was written for (no aliasing, no skipped phase),
  * a tile-id ring entry is never overwritten before every reader has consumed it,
  * a tile buffer is never reloaded / rewritten while a TMA store may still read it,
  * all roles terminate (no deadlock) and every tile is processed exactly once, in order, by each role.
No GPU, no product code is executed: this is a model of the protocol, kept next to the kernel it mirrors.
"""
import random

import pytest

RING = 16


class MBar:
    """mbarrier with an arrival count and a transaction count; `phase` = number of completed phases."""

    def __init__(self, count):
        self.count, self.arrived, self.tx, self.phase = count, 0, 0, 0

    def _maybe_complete(self):
        if self.arrived == self.count and self.tx == 0:
            self.arrived = 0
            self.phase += 1

    def arrive(self, n=1):
        self.arrived += n
        assert self.arrived <= self.count, "more arrivals than the barrier expects"
        self._maybe_complete()

    def arrive_expect_tx(self, ntx):
        self.tx += ntx
        self.arrive()

    def complete_tx(self):
        self.tx -= 1
        self._maybe_complete()

    def passes(self, parity):
        # try_wait.parity: true iff the most recently completed phase has this parity (fresh barrier: parity 1 passes)
        return ((self.phase - 1) & 1) == parity


class Sim:
    def __init__(self, seed, T, kblocks, stages, kgroup, nacc, nbuf, aux, lnf, max_delay=6):
        self.rng = random.Random(seed)
        self.T, self.kblocks, self.stages, self.kgroup, self.nacc, self.nbuf = T, kblocks, stages, kgroup, nacc, nbuf
        self.aux, self.lnf, self.max_delay = aux, lnf, max_delay
        self.merged = kgroup >= kblocks
        self.two = self.merged and stages % 2 == 0
        self.full = [MBar(1) for _ in range(stages)]
        self.empty = [MBar(1) for _ in range(stages)]
        self.tfull = [MBar(1) for _ in range(4)]
        self.tempty = [MBar(8) for _ in range(4)]
        self.xfull = [MBar(1) for _ in range(4)]
        self.bfree = [MBar(1) for _ in range(4)]
        self.sfull = [MBar(1) for _ in range(RING)]
        self.ring = [None] * RING
        self.events = []            # (due_step, seq, fn)
        self.step_no = 0
        self.seq = 0
        self.progress = {}          # reader name -> number of ring entries consumed
        self.buf_busy = [0] * 4     # outstanding TMA store reads per tile buffer
        self.done_tiles = {"mma": [], "epi": [], "aux": []}

    # ---- asynchronous completions ------------------------------------------------------------------
    def later(self, fn, after=None):
        due = self.step_no + (after if after is not None else self.rng.randint(1, self.max_delay))
        self.seq += 1
        self.events.append((due, self.seq, fn))

    def wait(self, bar, parity, intended):
        """generator: spin until the parity wait passes, then check it passed on the intended phase"""
        while not bar.passes(parity):
            yield
        assert bar.phase == intended + 1, f"wait meant phase {intended}, barrier has completed {bar.phase} phases"

    def read_ring(self, who, idx):
        yield from self.wait(self.sfull[idx % RING], (idx // RING) & 1, idx // RING)
        t = self.ring[idx % RING]
        self.progress[who] = idx + 1
        return t

    # ---- warp 16: producer + tile ids ----------------------------------------------------------------
    def producer(self):
        stage, phase, uses = 0, 0, [0] * self.stages
        for tl in range(self.T + 1):
            t = tl if tl < self.T else -1
            for j in ([tl, tl + 1] if t < 0 else [tl]):
                if j >= RING:                     # about to overwrite entry j - RING: every reader must be past it
                    for who, n in self.progress.items():
                        if who in self.readers_of(j - RING):
                            assert n > j - RING, f"ring entry {j - RING} overwritten before {who} read it"
                self.ring[j % RING] = t
                self.sfull[j % RING].arrive()
            if t < 0:
                return
            for kb in range(0, self.kblocks, self.kgroup):
                yield from self.wait(self.empty[stage], phase ^ 1, uses[stage] - 1)
                n = min(self.kgroup, self.kblocks - kb)
                bar = self.full[stage]
                bar.arrive_expect_tx(n)
                for _ in range(n):
                    self.later(bar.complete_tx)
                uses[stage] += 1
                stage += 1
                if stage == self.stages:
                    stage, phase = 0, phase ^ 1
            yield

    def readers_of(self, idx):
        r = {"mma%d" % (idx & 1 if self.two else 0), "epi%d" % (idx & 1)}
        if self.aux:
            r.add("aux")
        return r

    # ---- warps 17 / 19: MMA issuers -------------------------------------------------------------------
    def issuer(self, p):
        who = "mma%d" % p
        step = 2 if self.two else 1
        commits = []                              # in-order completion of this thread's commits

        def commit(bar):
            commits.append(bar)
            delay = self.rng.randint(1, self.max_delay)

            def fire():
                # commits of one thread complete in issue order
                while commits and commits[0] is not bar:
                    return self.later(fire, 1)
                commits.pop(0)
                bar.arrive()
            self.later(fire, delay)

        t = yield from self.read_ring(who, p)
        stage, phase = (p if self.two else 0), 0
        uses = [0] * self.stages
        tl = p
        while t >= 0:
            a, aph = tl % self.nacc, (tl // self.nacc) & 1
            nx = tl + step
            if self.merged:
                t_next = yield from self.read_ring(who, nx)
            yield from self.wait(self.tempty[a], aph ^ 1, tl // self.nacc - 1)
            for kb in range(0, self.kblocks, self.kgroup):
                yield from self.wait(self.full[stage], phase, uses[stage])
                uses[stage] += 1
                yield                                 # MMA issue
                commit(self.empty[stage])
                if kb + self.kgroup >= self.kblocks:
                    commit(self.tfull[a])
                if self.two:
                    stage += 2
                    if stage >= self.stages:
                        stage, phase = stage - self.stages, phase ^ 1
                else:
                    stage += 1
                    if stage == self.stages:
                        stage, phase = 0, phase ^ 1
            self.done_tiles["mma"].append(t)
            if not self.merged:
                t_next = yield from self.read_ring(who, nx)
            t = t_next
            tl += step

    # ---- warp 18: aux producer ------------------------------------------------------------------------
    def aux_warp(self):
        n = 0
        while True:
            t = yield from self.read_ring("aux", n)
            if t < 0:
                return
            b = n % self.nbuf
            yield from self.wait(self.bfree[b], ((n // self.nbuf) & 1) ^ 1, n // self.nbuf - 1)
            assert self.buf_busy[b] == 0, "aux tile loaded into a buffer a TMA store may still be reading"
            bar = self.xfull[b]
            bar.arrive_expect_tx(2)
            self.later(bar.complete_tx)
            self.later(bar.complete_tx)
            self.done_tiles["aux"].append(t)
            n += 1
            yield

    # ---- warps 0..15: one coroutine per epilogue set ----------------------------------------------------
    def epilogue(self, s):
        who = "epi%d" % s
        reads = []                                # this issuer thread's outstanding bulk groups, oldest first: (buffers)
        pending_b = -1

        def store(bufs):
            grp = {"bufs": bufs, "done": False}
            reads.append(grp)
            for b in bufs:
                self.buf_busy[b] += 1

            def fire():
                if reads and reads[0] is not grp and not reads[0]["done"] and grp in reads and reads.index(grp) > 0 \
                        and not all(g["done"] for g in reads[:reads.index(grp)]):
                    return self.later(fire, 1)    # bulk groups complete in order
                grp["done"] = True
                for b in grp["bufs"]:
                    self.buf_busy[b] -= 1
            self.later(fire)
            return grp

        def wait_read(keep):
            """cp.async.bulk.wait_group.read keep: all but the `keep` newest groups have been read"""
            while True:
                while reads and reads[0]["done"]:
                    reads.pop(0)
                if len(reads) <= keep:
                    return
                yield

        n = s
        while True:
            t = yield from self.read_ring(who, n)
            if t < 0:
                break
            a, aph = n % self.nacc, (n // self.nacc) & 1
            b, tph = n % self.nbuf, (n // self.nbuf) & 1
            yield from self.wait(self.tfull[a], aph, n // self.nacc)
            if self.aux:
                yield from self.wait(self.xfull[b], tph, n // self.nbuf)
            else:
                yield from self.wait(self.bfree[b], tph ^ 1, n // self.nbuf - 1)
            assert self.buf_busy[b] == 0, "epilogue writes a tile buffer a TMA store may still be reading"
            yield                                     # TMEM -> registers -> smem
            self.tempty[a].arrive(8)
            yield                                     # named barrier of the set
            if self.lnf:
                store([b])                            # x
                yield                                 # LayerNorm into the second staging buffer, second named barrier
                store([])                             # y (its own buffer, not modelled)
                yield from wait_read(0)
                self.bfree[b].arrive()
            elif self.nbuf == 4:
                store([b])
                if pending_b >= 0:
                    yield from wait_read(1)
                    self.bfree[pending_b].arrive()
                pending_b = b
            else:
                store([b])
                yield from wait_read(0)
                self.bfree[b].arrive()
            self.done_tiles["epi"].append(t)
            n += 2
        if pending_b >= 0:
            yield from wait_read(0)
            self.bfree[pending_b].arrive()

    # ---- scheduler --------------------------------------------------------------------------------------
    def run(self, max_steps=200000):
        procs = {"producer": self.producer(), "mma0": self.issuer(0), "epi0": self.epilogue(0), "epi1": self.epilogue(1)}
        if self.two:
            procs["mma1"] = self.issuer(1)
        if self.aux:
            procs["aux"] = self.aux_warp()
        for k in procs:
            if k != "producer":
                self.progress[k] = 0
        while procs:
            self.step_no += 1
            assert self.step_no < max_steps, f"deadlock / livelock: {sorted(procs)} still running"
            due = [e for e in self.events if e[0] <= self.step_no]
            self.events = [e for e in self.events if e[0] > self.step_no]
            for _, _, fn in sorted(due, key=lambda e: (e[0], e[1])):
                fn()
            name = self.rng.choice(sorted(procs))
            try:
                next(procs[name])
            except StopIteration:
                del procs[name]
        # drain outstanding asynchronous work
        while self.events:
            self.step_no += 1
            due = [e for e in self.events if e[0] <= self.step_no]
            self.events = [e for e in self.events if e[0] > self.step_no]
            for _, _, fn in sorted(due, key=lambda e: (e[0], e[1])):
                fn()
        tiles = list(range(self.T))
        assert sorted(self.done_tiles["mma"]) == tiles and sorted(self.done_tiles["epi"]) == tiles
        if self.aux:
            assert self.done_tiles["aux"] == tiles
        assert all(v == 0 for v in self.buf_busy)


def host_config(kblocks, aux, lnf, narrow=True):
    """the combinations srfrd_gemm_tn's host code can produce (gemm.cu, srfrd_gemm_tn)"""
    out = []
    for stages in (2, 3, 4, 5, 6):
        for kgroup in {1, kblocks}:
            for nacc in ((4,) if narrow else (2,)):
                for nbuf in ((2, 4) if (aux and not lnf) else (2,)):
                    out.append(dict(stages=stages, kgroup=kgroup, nacc=nacc, nbuf=nbuf))
    return out


@pytest.mark.parametrize("kblocks", [1, 2, 3, 5])
@pytest.mark.parametrize("aux,lnf", [(False, False), (True, False), (True, True)])
def test_gemm_tn_pipeline_protocol(kblocks, aux, lnf):
    n = 0
    for narrow in (True, False):
        for cfg in host_config(kblocks, aux, lnf, narrow):
            for T in (1, 2, 3, 5, 11, 23, 40):
                for seed in range(3):
                    Sim(seed * 7919 + T, T, kblocks, aux=aux, lnf=lnf, max_delay=(2 if seed == 0 else 9), **cfg).run()
                    n += 1
    assert n > 100


def test_model_detects_a_skipped_phase():
    """the checker itself: a consumer that falls two phases behind a count-1 barrier is reported"""
    sim = Sim(0, 1, 1, 2, 1, 4, 2, False, False)
    bar = MBar(1)
    bar.arrive()
    bar.arrive()
    bar.arrive()                                   # three phases completed; the waiter meant phase 0 (parity 0)
    g = sim.wait(bar, 0, 0)
    with pytest.raises(AssertionError, match="meant phase 0"):
        next(g)


def _flags(make, runs=8):
    n = 0
    for seed in range(runs):
        try:
            make(seed).run(max_steps=20000)
        except AssertionError:
            n += 1
    return n


def test_model_flags_a_ring_that_is_too_small(monkeypatch):
    """producer up to `stages` tiles ahead of the MMAs + `nacc` accumulators ahead of the epilogues needs > 4 entries"""
    import sys
    monkeypatch.setattr(sys.modules[__name__], "RING", 4)
    assert _flags(lambda seed: Sim(seed, 23, 1, stages=6, kgroup=1, nacc=4, nbuf=2, aux=False, lnf=False)) == 8


def test_model_flags_two_issuers_on_an_odd_stage_count():
    """the kernel only enables the second issuer when stages % 2 == 0: with 3 stages the issuers would share barriers"""
    def make(seed):
        sim = Sim(seed, 23, 2, stages=3, kgroup=2, nacc=4, nbuf=2, aux=True, lnf=False)
        sim.two = True
        return sim
    assert _flags(make) == 8


def test_model_flags_the_lookahead_deadlock():
    """the bug found on the GPU this round: reading the NEXT ring entry before issuing a tile whose K blocks outnumber the
    pipeline stages deadlocks (the producer needs the tile's first stages back before it can publish the next id)"""
    def make(seed):
        sim = Sim(seed, 5, 5, stages=3, kgroup=1, nacc=2, nbuf=2, aux=False, lnf=False)
        sim.merged = True                          # forces the early ring read of the merged path
        return sim
    assert _flags(make) == 8


# =========================================================================================================
# catalogue_tilemax_kernel (srfrd_b200/csrc/topk.cu): NACC issuer warps, two epilogue sets, pieces of item tiles
# =========================================================================================================
class TopkSim(Sim):
    """Same machinery for the catalogue kernel's protocol.  The item-tile ring is consumed by issuer (tile % NACC); an
    issuer SKIPS the other issuers' stages without touching their barriers, which is only sound when the ring length is
    a multiple of NACC * kblocks (host: make_plan).  tfull is indexed [accumulator][set], tempty[accumulator] collects the
    arrivals of whichever set read it, afull / aempty hand the user tiles in TMEM from set 0 to the issuers per piece."""

    def __init__(self, seed, pieces, kblocks, stages, nacc, max_delay=6):
        Sim.__init__(self, seed, sum(pieces), kblocks, stages, 1, nacc, 2, False, False, max_delay)
        self.pieces = pieces
        self.tfull2 = [MBar(1) for _ in range(2 * nacc)]
        self.tempty = [MBar(1) for _ in range(nacc)]       # (4 * UBS warp arrivals modelled as one)
        self.afull, self.aempty = MBar(1), MBar(nacc)
        self.meet = [0, 0]                                  # per-piece rendezvous of the two sets (named barrier)
        self.seen = {"mma": [], "epi": []}

    def t_producer(self):
        stage, phase, uses = 0, 0, [0] * self.stages
        for ntiles in self.pieces:
            for _ in range(ntiles):
                for _ in range(self.kblocks):
                    yield from self.wait(self.empty[stage], phase ^ 1, uses[stage] - 1)
                    bar = self.full[stage]
                    bar.arrive_expect_tx(1)
                    self.later(bar.complete_tx)
                    uses[stage] += 1
                    stage += 1
                    if stage == self.stages:
                        stage, phase = 0, phase ^ 1
                yield

    def t_issuer(self, w):
        commits = []

        def commit(bar):
            commits.append(bar)

            def fire():
                if commits[0] is not bar:
                    return self.later(fire, 1)
                commits.pop(0)
                bar.arrive()
            self.later(fire)

        stage, phase, uphase, aphase = 0, 0, 0, 0
        n, own, passes = 0, 0, 0
        for pi, ntiles in enumerate(self.pieces):
            yield from self.wait(self.afull, uphase, pi)
            uphase ^= 1
            for _ in range(ntiles):
                if n % self.nacc != w:
                    stage += self.kblocks
                    if stage >= self.stages:
                        stage, phase, passes = stage - self.stages, phase ^ 1, passes + 1
                    n += 1
                    continue
                yield from self.wait(self.tempty[w], aphase ^ 1, own - 1)
                aphase ^= 1
                for kb in range(self.kblocks):
                    yield from self.wait(self.full[stage], phase, passes)
                    yield
                    if kb == self.kblocks - 1:
                        commit(self.tfull2[w * 2 + (n & 1)])
                    commit(self.empty[stage])
                    stage += 1
                    if stage == self.stages:
                        stage, phase, passes = 0, phase ^ 1, passes + 1
                self.seen["mma"].append(n)
                own += 1
                n += 1
            commit(self.aempty)

    def t_epilogue(self, s):
        uphase, cnt, n = 0, [0] * self.nacc, 0
        for pi, ntiles in enumerate(self.pieces):
            if s == 0:
                yield from self.wait(self.aempty, uphase ^ 1, pi - 1)
                uphase ^= 1
                yield                                     # stage the user tiles into TMEM
                self.afull.arrive()
            self.meet[s] += 1                             # named barrier: both threads of a row start the piece together
            while self.meet[1 - s] < self.meet[s]:
                yield
            for _ in range(ntiles):
                if (n & 1) != s:
                    n += 1
                    continue
                a = n % self.nacc
                par, use = cnt[a] & 1, cnt[a]
                cnt[a] += 1
                yield from self.wait(self.tfull2[a * 2 + s], par, use)
                yield                                     # tcgen05.ld
                self.tempty[a].arrive()
                self.seen["epi"].append(n)
                n += 1

    def run(self, max_steps=400000):
        procs = {"producer": self.t_producer(), "epi0": self.t_epilogue(0), "epi1": self.t_epilogue(1)}
        for w in range(self.nacc):
            procs["mma%d" % w] = self.t_issuer(w)
        while procs:
            self.step_no += 1
            assert self.step_no < max_steps, f"deadlock / livelock: {sorted(procs)} still running"
            due = [e for e in self.events if e[0] <= self.step_no]
            self.events = [e for e in self.events if e[0] > self.step_no]
            for _, _, fn in sorted(due, key=lambda e: (e[0], e[1])):
                fn()
            name = self.rng.choice(sorted(procs))
            try:
                next(procs[name])
            except StopIteration:
                del procs[name]
        tiles = list(range(sum(self.pieces)))
        assert sorted(self.seen["mma"]) == tiles and sorted(self.seen["epi"]) == tiles


@pytest.mark.parametrize("nacc", [2, 3])
@pytest.mark.parametrize("kblocks", [1, 2, 4])
def test_catalogue_kernel_protocol(nacc, kblocks):
    unit = nacc * kblocks
    for mult in (1, 2, 5):
        for pieces in ([1], [2], [7], [3, 1, 4], [1, 1, 1, 1], [11, 6], [40]):
            for seed in range(3):
                TopkSim(seed + 13 * mult, pieces, kblocks, stages=unit * mult, nacc=nacc, max_delay=(2 if seed == 0 else 9)).run()


def test_catalogue_model_flags_a_ring_that_is_not_a_multiple_of_the_issuer_count():
    """make_plan rounds the ring down to a multiple of NACC * kblocks; without that a stage is waited for by different
    issuers in turn, and when TMA completions arrive far out of order one of them passes on the PREVIOUS phase"""
    flagged = 0
    for seed in range(16):
        try:
            TopkSim(seed, [23], 1, stages=4, nacc=3, max_delay=1000).run(max_steps=400000)
        except AssertionError as e:
            assert "wait meant phase" in str(e)
            flagged += 1
    assert flagged >= 4


def test_protocols_hold_under_extreme_reordering():
    """completion delays two orders of magnitude longer than a tile's issue time (loads, commits, store reads far out of order)"""
    for seed in range(4):
        for nacc, kb, mult in ((3, 1, 2), (2, 2, 1), (3, 2, 3)):
            TopkSim(seed, [11, 6, 9], kb, stages=nacc * kb * mult, nacc=nacc, max_delay=400).run(max_steps=2000000)
        Sim(seed, 23, 2, stages=4, kgroup=2, nacc=4, nbuf=4, aux=True, lnf=False, max_delay=400).run(max_steps=2000000)
        Sim(seed, 23, 2, stages=2, kgroup=2, nacc=4, nbuf=2, aux=True, lnf=True, max_delay=400).run(max_steps=2000000)
        Sim(seed, 23, 5, stages=3, kgroup=1, nacc=2, nbuf=2, aux=True, lnf=False, max_delay=400).run(max_steps=2000000)


# =========================================================================================================
# attn_long_fwd_kernel (srfrd_b200/csrc/attention_long.cu): one ring of [128 x 64] tiles feeds S = Q K^T and O = P V
# =========================================================================================================
class LongFwdSim(Sim):
    """Producer and MMA issuer walk the same tile sequence (per K block: Q tile, K tiles 0..i; then per 64-column output
    block: V tiles 0..i) through one ring; the Q slot is only released with the LAST key tile of its K block, so the ring
    must hold i + 2 tiles at least.  Output blocks go through 4 accumulator slots (o_full / o_empty, use count -> parity),
    softmax warps of `part` p read the slots with slot == p; s_full / p_full / pv_done are per-item barriers."""

    def __init__(self, seed, items, kblocks, ring, max_delay=6):
        Sim.__init__(self, seed, len(items), kblocks, ring, 1, 2, 2, False, False, max_delay)
        self.items = items                                  # query-tile index i of every item this CTA handles
        self.s_full, self.p_full, self.pv_done = MBar(1), MBar(4), MBar(1)     # p_full: 16 warps = 4 per modelled part
        self.o_full = [MBar(1) for _ in range(4)]
        self.o_empty = [MBar(4) for _ in range(4)]          # the four lane quarters of one part
        self.items_done = {"mma": 0, "sm": [0, 0, 0, 0]}

    def l_producer(self):
        slot, ph, uses = 0, 0, [0] * self.stages

        def push():
            nonlocal slot, ph
            yield from self.wait(self.empty[slot], ph ^ 1, uses[slot] - 1)
            bar = self.full[slot]
            bar.arrive_expect_tx(1)
            self.later(bar.complete_tx)
            uses[slot] += 1
            slot += 1
            if slot == self.stages:
                slot, ph = 0, ph ^ 1
        for i in self.items:
            for _ in range(self.kblocks):
                yield from push()                           # Q
                for _ in range(i + 1):
                    yield from push()                       # K_j
            for _ in range(self.kblocks):
                for _ in range(i + 1):
                    yield from push()                       # V_j
            yield

    def l_mma(self):
        commits = []

        def commit(bar):
            commits.append(bar)

            def fire():
                if commits[0] is not bar:
                    return self.later(fire, 1)
                commits.pop(0)
                bar.arrive()
            self.later(fire)

        slot, ph, iph, ocnt = 0, 0, 0, 0
        uses = [0] * self.stages

        def take():
            nonlocal slot, ph
            yield from self.wait(self.full[slot], ph, uses[slot])
            uses[slot] += 1
            s = slot
            slot += 1
            if slot == self.stages:
                slot, ph = 0, ph ^ 1
            return s
        for n, i in enumerate(self.items):
            for _ in range(self.kblocks):
                q = yield from take()
                for j in range(i + 1):
                    k = yield from take()
                    yield
                    commit(self.empty[k])
                    if j == i:
                        commit(self.empty[q])
            commit(self.s_full)
            yield from self.wait(self.p_full, iph, n)
            for cb in range(self.kblocks):
                o = ocnt & 3
                yield from self.wait(self.o_empty[o], ((ocnt >> 2) & 1) ^ 1, (ocnt >> 2) - 1)
                for _ in range(i + 1):
                    v = yield from take()
                    yield
                    commit(self.empty[v])
                commit(self.o_full[o])
                if cb == self.kblocks - 1:
                    commit(self.pv_done)
                ocnt += 1
            iph ^= 1
            self.items_done["mma"] += 1

    def l_softmax(self, part):
        """the four warps (lane quarters) of one `part`"""
        iph, ocnt = 0, 0
        for n, _ in enumerate(self.items):
            yield from self.wait(self.s_full, iph, n)
            yield                                           # S -> registers, exponentials
            if n:
                yield from self.wait(self.pv_done, iph ^ 1, n - 1)
            yield                                           # P -> shared memory
            self.p_full.arrive()
            for cb in range(self.kblocks):
                oc = ocnt + cb
                if (oc & 3) != part:
                    continue
                yield from self.wait(self.o_full[oc & 3], (oc >> 2) & 1, oc >> 2)
                yield
                self.o_empty[oc & 3].arrive(4)
            ocnt += self.kblocks
            iph ^= 1
            self.items_done["sm"][part] += 1

    def run(self, max_steps=400000):
        procs = {"producer": self.l_producer(), "mma": self.l_mma()}
        for p in range(4):
            procs["sm%d" % p] = self.l_softmax(p)
        while procs:
            self.step_no += 1
            assert self.step_no < max_steps, f"deadlock / livelock: {sorted(procs)} still running"
            due = [e for e in self.events if e[0] <= self.step_no]
            self.events = [e for e in self.events if e[0] > self.step_no]
            for _, _, fn in sorted(due, key=lambda e: (e[0], e[1])):
                fn()
            name = self.rng.choice(sorted(procs))
            try:
                next(procs[name])
            except StopIteration:
                del procs[name]
        assert self.items_done["mma"] == len(self.items) and self.items_done["sm"] == [len(self.items)] * 4


@pytest.mark.parametrize("kblocks", [1, 2, 5, 8])            # head widths 64 ... 512
def test_long_attention_forward_protocol(kblocks):
    for items in ([0], [1], [0, 1], [1, 1, 0, 1, 0, 0, 1], [0] * 9, [1] * 9):
        for seed in range(3):
            LongFwdSim(seed, items, kblocks, ring=9, max_delay=(2 if seed == 0 else 12)).run()
    LongFwdSim(5, [1, 0, 1, 1, 0], kblocks, ring=9, max_delay=300).run(max_steps=3000000)


def test_long_attention_model_flags_a_ring_shorter_than_a_query_tiles_keys():
    """the Q slot is held until its last key tile has been multiplied: with i = 1 a ring of 2 tiles can never advance"""
    with pytest.raises(AssertionError, match="deadlock"):
        LongFwdSim(0, [1, 1], 2, ring=2).run(max_steps=20000)


# =========================================================================================================
# attn_long_bwd_kernel<MODE> (attention_long.cu): FlashAttention-2 style passes over (query tile, key tile) pairs
# =========================================================================================================
class LongBwdSim(Sim):
    """Per pair: S (and, except in the dV pass, dP into the SAME 128 TMEM columns once S is in registers), then one output
    MMA batch per 64-column block from the P / dS tile the workers wrote to shared memory.  Barriers: s_full, s_read(8),
    dp_full, sd_free(8), x_full(8), x_done per pair; acc_full, acc_free(8) per item.  The eight worker warps are modelled as
    two processes (the two `part`s) arriving with four each, so that they may drift apart."""

    def __init__(self, seed, items, kblocks, ring, dv, max_delay=6):
        Sim.__init__(self, seed, len(items), kblocks, ring, 1, 2, 2, False, False, max_delay)
        self.items, self.dv = items, dv                       # pairs per item
        self.s_full, self.dp_full, self.x_done, self.acc_full = MBar(1), MBar(1), MBar(1), MBar(1)
        self.s_read, self.sd_free, self.x_full, self.acc_free = MBar(8), MBar(8), MBar(8), MBar(8)
        self.done = {"mma": 0, "w": [0, 0]}

    def b_producer(self):
        slot, ph, uses = 0, 0, [0] * self.stages

        def push():
            nonlocal slot, ph
            yield from self.wait(self.empty[slot], ph ^ 1, uses[slot] - 1)
            bar = self.full[slot]
            bar.arrive_expect_tx(1)
            self.later(bar.complete_tx)
            uses[slot] += 1
            slot += 1
            if slot == self.stages:
                slot, ph = 0, ph ^ 1
        for npairs in self.items:
            for _ in range(npairs):
                for _ in range(self.kblocks * (2 if self.dv else 4)):
                    yield from push()
                for _ in range(self.kblocks):
                    yield from push()
            yield

    def b_mma(self):
        commits = []

        def commit(bar):
            commits.append(bar)

            def fire():
                if commits[0] is not bar:
                    return self.later(fire, 1)
                commits.pop(0)
                bar.arrive()
            self.later(fire)

        slot, ph, np_, uses = 0, 0, 0, [0] * self.stages

        def take():
            nonlocal slot, ph
            yield from self.wait(self.full[slot], ph, uses[slot])
            uses[slot] += 1
            s = slot
            slot += 1
            if slot == self.stages:
                slot, ph = 0, ph ^ 1
            return s
        for ni, npairs in enumerate(self.items):
            for pr in range(npairs):
                pp = np_ & 1
                yield from self.wait(self.sd_free, pp ^ 1, np_ - 1)
                for ph2 in range(1 if self.dv else 2):
                    if ph2 == 1:
                        yield from self.wait(self.s_read, pp, np_)
                    for _ in range(self.kblocks):
                        a = yield from take()
                        b = yield from take()
                        yield
                        commit(self.empty[a])
                        commit(self.empty[b])
                    commit(self.s_full if ph2 == 0 else self.dp_full)
                yield from self.wait(self.x_full, pp, np_)
                if pr == 0:
                    yield from self.wait(self.acc_free, (ni & 1) ^ 1, ni - 1)
                for cb in range(self.kblocks):
                    c = yield from take()
                    yield
                    commit(self.empty[c])
                    if cb == self.kblocks - 1:
                        commit(self.x_done)
                        if pr == npairs - 1:
                            commit(self.acc_full)
                np_ += 1
            self.done["mma"] += 1

    def b_worker(self, part):
        np_ = 0
        for ni, npairs in enumerate(self.items):
            for _ in range(npairs):
                pp = np_ & 1
                yield from self.wait(self.s_full, pp, np_)
                yield                                         # S -> registers
                (self.sd_free if self.dv else self.s_read).arrive(4)
                if not self.dv:
                    yield from self.wait(self.dp_full, pp, np_)
                    yield                                     # dP -> registers
                    self.sd_free.arrive(4)
                yield from self.wait(self.x_done, pp ^ 1, np_ - 1)
                yield                                         # P / dS -> shared memory
                self.x_full.arrive(4)
                np_ += 1
            yield from self.wait(self.acc_full, ni & 1, ni)
            yield                                             # output rows
            self.acc_free.arrive(4)
            self.done["w"][part] += 1

    def run(self, max_steps=600000):
        procs = {"producer": self.b_producer(), "mma": self.b_mma(), "w0": self.b_worker(0), "w1": self.b_worker(1)}
        while procs:
            self.step_no += 1
            assert self.step_no < max_steps, f"deadlock / livelock: {sorted(procs)} still running"
            due = [e for e in self.events if e[0] <= self.step_no]
            self.events = [e for e in self.events if e[0] > self.step_no]
            for _, _, fn in sorted(due, key=lambda e: (e[0], e[1])):
                fn()
            name = self.rng.choice(sorted(procs))
            try:
                next(procs[name])
            except StopIteration:
                del procs[name]
        assert self.done["mma"] == len(self.items) and self.done["w"] == [len(self.items)] * 2


@pytest.mark.parametrize("dv", [False, True])
@pytest.mark.parametrize("kblocks", [1, 2, 5, 8])
def test_long_attention_backward_protocol(kblocks, dv):
    for items in ([1], [2], [1, 2], [2, 1, 2, 2, 1, 1], [1] * 7, [2] * 7):
        for seed in range(3):
            LongBwdSim(seed, items, kblocks, ring=11, dv=dv, max_delay=(2 if seed == 0 else 12)).run()
    LongBwdSim(9, [2, 1, 2, 1], kblocks, ring=11, dv=dv, max_delay=300).run(max_steps=4000000)


def test_long_backward_model_flags_a_miscounted_barrier():
    """s_read must collect all eight worker warps before dP may overwrite S in the shared TMEM columns: initialised with a
    count of 4 it completes a phase per `part`, and its single waiter (the MMA issuer) is found off by a phase"""
    sim = LongBwdSim(0, [2, 2, 2], 1, ring=11, dv=False)
    sim.s_read = MBar(4)
    with pytest.raises(AssertionError):
        sim.run(max_steps=50000)
