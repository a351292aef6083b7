"""A host-side model of the ordered-reassembly protocol (csrc/t2_gather.cu) driven the way bench.py drives it: streams
are FIFO queues of operations, counters are the only coupling between ranks, a host `sync` blocks until the rank's
streams have drained, a collective blocks until every rank has reached it.  It shows the property bench.py's multi-rank
warm-up relies on -- with the stop decision made collectively no schedule deadlocks -- and that the time-based loop per
rank it replaced does deadlock once the ranks' clocks are skewed by more than the ring allows (the hang seen at N = 4)."""
import random

N_SLOTS = 2


class Model(object):
    def __init__(self, world):
        self.world = world
        self.arrived = [0] * world                 # root's control block
        self.released = [0] * world                # every rank's released counter
        self.events = set()                        # recorded CUDA events
        # streams: producer / side per rank, consumer on the root; each a list of pending ops (callables -> bool done)
        self.streams = {("p", r): [] for r in range(world)}
        self.streams.update({("s", r): [] for r in range(world)})
        self.streams[("c", 0)] = []

    def enqueue_step(self, r, k):
        w = self.world
        if r == 0:
            if k >= N_SLOTS:
                self.streams[("p", 0)].append(lambda k=k: ("rel", k - N_SLOTS) in self.events)
            self.streams[("p", 0)].append(lambda k=k: self._set_arrived(0, k + 1))
            for q in range(w):
                self.streams[("c", 0)].append(lambda k=k, q=q: self.arrived[q] >= k + 1)
            self.streams[("c", 0)].append(lambda k=k: self._release(k))
        else:
            if k >= 2:
                self.streams[("p", r)].append(lambda k=k, r=r: ("pushed", r, k - 2) in self.events)
            self.streams[("p", r)].append(lambda k=k, r=r: self._record(("produced", r, k)))
            self.streams[("s", r)].append(lambda k=k, r=r: ("produced", r, k) in self.events)
            if k >= N_SLOTS:
                self.streams[("s", r)].append(lambda k=k, r=r: self.released[r] >= k - N_SLOTS + 1)
            self.streams[("s", r)].append(lambda k=k, r=r: self._set_arrived(r, k + 1))
            self.streams[("s", r)].append(lambda k=k, r=r: self._record(("pushed", r, k)))

    def _record(self, e):
        self.events.add(e)
        return True

    def _set_arrived(self, r, v):
        self.arrived[r] = max(self.arrived[r], v)
        return True

    def _release(self, k):
        for r in range(1, self.world):
            self.released[r] = max(self.released[r], k + 1)
        self.events.add(("rel", k))
        return True

    def run_devices(self):
        progress = True
        while progress:
            progress = False
            for q in self.streams.values():
                while q and q[0]():
                    q.pop(0)
                    progress = True

    def drained(self, r):
        return not self.streams[("p", r)] and not self.streams[("s", r)] and (r != 0 or not self.streams[("c", 0)])


def simulate(world, programs, max_ticks=200000):
    """programs[r] = generator yielding host actions: ("step",), ("sync",), ("coll", value) -> receives the OR of the values.
    Returns True when every rank's program ends, False on deadlock."""
    m = Model(world)
    gens = [p(r) for r, p in enumerate(programs)]
    state = [None] * world            # pending blocking action per rank
    steps = [0] * world
    done = [False] * world
    send = [None] * world
    for _ in range(max_ticks):
        moved = False
        for r in range(world):
            if done[r]:
                continue
            if state[r] is None:
                try:
                    state[r] = gens[r].send(send[r])
                    send[r] = None
                except StopIteration:
                    done[r] = True
                    moved = True
                    continue
            a = state[r]
            if a[0] == "step":
                m.enqueue_step(r, steps[r])
                steps[r] += 1
                state[r] = None
                moved = True
            elif a[0] == "sync":
                m.run_devices()
                if m.drained(r):
                    state[r] = None
                    moved = True
        m.run_devices()
        waiting = [r for r in range(world) if not done[r] and state[r] is not None and state[r][0] == "coll"]
        if len(waiting) == world - sum(done) and waiting and len(waiting) == world:
            res = any(state[r][1] for r in waiting)
            for r in waiting:
                send[r] = res
                state[r] = None
            moved = True
        if all(done):
            return True
        if not moved:
            return False
    return False


def collective_program(budget):
    """bench.py / shard.collective_warmup: chunks of 8 steps, sync, collective stop decision; `budget[r]` = how many chunks
    rank r's own clock would like to run (skewed clocks)."""
    def prog(r):
        n = 0
        while True:
            for _ in range(8):
                yield ("step",)
            n += 1
            yield ("sync",)
            stop = yield ("coll", n >= budget[r])
            if stop:
                break
        for _ in range(20):           # the timed region: the same count everywhere, no host sync in between
            yield ("step",)
        yield ("sync",)
        yield ("coll", True)
    return prog


def time_based_program(count):
    """the replaced loop: rank r runs `count[r]` steps by its own clock, synchronising every 8, THEN equalises."""
    def prog(r):
        for i in range(count[r]):
            yield ("step",)
            if (i + 1) % 8 == 0:
                yield ("sync",)
        yield ("coll", True)
    return prog


def test_collective_warmup_never_deadlocks():
    rng = random.Random(7)
    for world in (2, 4, 8):
        for _ in range(40):
            budget = [rng.randint(1, 12) for _ in range(world)]
            assert simulate(world, [collective_program(budget)] * world), (world, budget)


def test_time_based_warmup_deadlocks_under_skew():
    # equal counts are fine ...
    assert simulate(4, [time_based_program([40, 40, 40, 40])] * 4)
    # ... a rank that runs 16 steps past the root's last step synchronises on releases the root never issues
    assert not simulate(4, [time_based_program([40, 56, 40, 40])] * 4)
    # ... and so does a root that runs past the others (it waits for arrivals that never come)
    assert not simulate(4, [time_based_program([56, 40, 40, 40])] * 4)
