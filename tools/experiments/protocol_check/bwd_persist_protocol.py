"""Randomised schedule exploration of the mbarrier protocol of attn_tc_bwd_persist.cuh (CPU only, no GPU needed).

The three roles of the kernel (TMA producer, MMA issuer, elementwise warps — one representative thread, arrival counts
scaled to 1) are written as generators that yield at every mbarrier wait; barriers are modelled as phase counters with
parity waits exactly like `mbarrier.try_wait.parity` (a waiter that falls two phases behind would alias — that is
detected).  Asynchronous completions (TMA loads landing, tcgen05.commit arrivals) are separate events that the scheduler
fires in random order, but commits of one thread complete in issue order.  Resources carry an owner / state and every
access asserts the hazard rules (e.g. the tensor core may not overwrite S / dP before the elementwise warps have them in
registers, P / dS may not be overwritten while the previous tile's MMAs still read them).  Thousands of random schedules
over random item lists (including empty items) must all terminate without deadlock or assertion.
Mutation check (done when this was written): removing the `mma_done` wait or the `kv_empty` wait is caught (P/dS resp. K/V
overwritten while in use); removing the `acc_empty` wait or the `sdp_free` wait at an item start is NOT — both are implied
by the `pds_full` wait of the neighbouring iteration, i.e. they are redundant (harmless) in the kernel.
Run: python tools/experiments/protocol_check/bwd_persist_protocol.py"""
import random
import sys

QS = 3


class Bar:
    def __init__(self, name):
        self.name, self.phase = name, 0          # number of completed phases

    def complete(self):
        self.phase += 1

    def passed(self, parity):
        """try_wait.parity(p): true iff the phase with parity p has completed = current phase parity != p."""
        return (self.phase & 1) != parity


class Deadlock(Exception):
    pass


def run(items, seed):
    rnd = random.Random(seed)
    bars = {n: Bar(n) for n in ["kv_full", "kv_empty", "sdp_full", "sdp_free", "pds_full", "mma_done", "acc_full", "acc_empty",
                                "dq_full0", "dq_full1"] + [f"qdo_full{i}" for i in range(QS)] + [f"qdo_empty{i}" for i in range(QS)]}
    pending = []                 # async events: (thread, seq, fn) — same-thread commits fire in order; TMA loads any order
    seqs = {"mma": 0}
    # resource state
    st = {"sdp": "free", "pds": "free", "kv": None, "acc": "free", "dq": ["free", "free"], "qdo": [None] * QS}
    log = []

    def commit(bar, also=None):
        """tcgen05.commit by the MMA thread: arrives when all its earlier MMAs are done → in issue order."""
        seqs["mma"] += 1
        pending.append(("mma", seqs["mma"], (bar, also)))

    def tma(bar, also):
        pending.append(("tma", rnd.random(), (bar, also)))

    def wait(bar, parity, expect_phase):
        """yield until passed; expect_phase = index of the completion we mean — detects parity aliasing."""
        while not bars[bar].passed(parity):
            yield
        assert bars[bar].phase == expect_phase + 1 or bars[bar].phase == expect_phase + 2 and False, \
            f"aliasing on {bar}: waited for completion #{expect_phase}, barrier already at {bars[bar].phase}"

    def producer():
        g = kc = 0
        for n_it in items:
            if n_it == 0:
                continue
            if kc > 0:
                yield from wait("kv_empty", (kc - 1) & 1, kc - 1)
            assert st["kv"] is None, "K/V overwritten while in use"
            item_id = kc
            tma("kv_full", lambda item_id=item_id: st.__setitem__("kv", item_id))
            for it in range(n_it):
                s, u = g % QS, g // QS
                if u > 0:
                    yield from wait(f"qdo_empty{s}", ((u & 1) ^ 1), u - 1)
                assert st["qdo"][s] is None, "Q/dO stage overwritten while in use"
                tma(f"qdo_full{s}", lambda s=s, g=g: st["qdo"].__setitem__(s, g))
                g += 1
            kc += 1

    def mma():
        g = kc = 0

        def issue_sdp(gi):
            s = gi % QS
            yield from wait(f"qdo_full{s}", (gi // QS) & 1, gi // QS)
            assert st["qdo"][s] == gi and st["kv"] is not None
            assert st["sdp"] == "free", f"S/dP overwritten before consumed (g={gi})"
            st["sdp"] = "mma"
            commit("sdp_full", lambda gi=gi: st.__setitem__("sdp", ("ready", gi)))

        for n_it in items:
            if n_it == 0:
                continue
            yield from wait("kv_full", kc & 1, kc)
            assert st["kv"] == kc
            if g > 0:
                yield from wait("sdp_free", (g - 1) & 1, g - 1)
            yield from issue_sdp(g)
            for it in range(n_it):
                s, sq, last = g & 1, g % QS, it + 1 == n_it
                if not last:
                    yield from wait("sdp_free", g & 1, g)
                    yield from issue_sdp(g + 1)
                yield from wait("pds_full", g & 1, g)
                assert st["pds"] == ("ready", g), f"P/dS not those of iteration {g}: {st['pds']}"
                assert st["dq"][s] == "free", f"dQ buffer {s} overwritten before drained (g={g})"
                st["dq"][s] = "mma"
                commit(f"dq_full{s}", lambda s=s, g=g: st["dq"].__setitem__(s, ("ready", g)))
                if it == 0 and kc > 0:
                    yield from wait("acc_empty", (kc - 1) & 1, kc - 1)
                if it == 0:
                    assert st["acc"] == "free", "dK/dV accumulators overwritten before stored"
                    st["acc"] = "mma"
                assert st["qdo"][sq] == g

                def done(sq=sq, g=g, last=last, kc=kc):
                    st["qdo"][sq] = None
                    st["pds"] = "free"
                commit(f"qdo_empty{sq}", done)
                commit("mma_done")
                if last:
                    commit("kv_empty", lambda: st.__setitem__("kv", None))
                    commit("acc_full", lambda kc=kc: st.__setitem__("acc", ("ready", kc)))
                g += 1
            kc += 1

    def elementwise():
        g = kc = 0

        def drain(gi):
            yield from wait(f"dq_full{gi & 1}", (gi >> 1) & 1, gi >> 1)
            assert st["dq"][gi & 1] == ("ready", gi), f"drain of {gi} sees {st['dq']}"
            st["dq"][gi & 1] = "free"

        for n_it in items:
            for it in range(n_it):
                yield from wait("sdp_full", g & 1, g)
                assert st["sdp"] == ("ready", g), f"S/dP of {g} expected, {st['sdp']}"
                st["sdp"] = "free"                       # in registers
                bars["sdp_free"].complete()
                if g > 0:
                    yield from wait("mma_done", (g - 1) & 1, g - 1)
                assert st["pds"] == "free", f"P/dS overwritten while MMAs read them (g={g})"
                st["pds"] = ("ready", g)
                bars["pds_full"].complete()
                if it > 0:
                    yield from drain(g - 1)
                g += 1
            if n_it > 0:
                yield from drain(g - 1)
                yield from wait("acc_full", kc & 1, kc)
                assert st["acc"] == ("ready", kc)
                st["acc"] = "free"
                bars["acc_empty"].complete()
                kc += 1

    roles = {"producer": producer(), "mma": mma(), "elementwise": elementwise()}
    alive = dict(roles)
    steps = 0
    idle = 0
    while alive or pending:
        steps += 1
        choices = list(alive)
        fire = None
        if pending:
            # eligible async events: any TMA; the OLDEST outstanding commit only (in-order completion)
            commits = [e for e in pending if e[0] == "mma"]
            tmas = [e for e in pending if e[0] == "tma"]
            elig = tmas + ([min(commits, key=lambda e: e[1])] if commits else [])
            if elig and (not choices or rnd.random() < 0.4):
                fire = rnd.choice(elig)
        if fire is not None:
            pending.remove(fire)
            bar, also = fire[2]
            if also:
                also()
            bars[bar].complete()
            idle = 0
            continue
        if not choices:
            continue
        name = rnd.choice(choices)
        before = (tuple(b.phase for b in bars.values()), len(pending))
        try:
            next(alive[name])
        except StopIteration:
            del alive[name]
        after = (tuple(b.phase for b in bars.values()), len(pending))
        idle = 0 if (before != after or name not in alive) else idle + 1
        if idle > 2000 and not pending:
            raise Deadlock(f"items={items} seed={seed}: stuck with {list(alive)}; bars={ {k: v.phase for k, v in bars.items()} }")
    return steps


def main():
    rnd = random.Random(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    for trial in range(n):
        items = [rnd.choice([0, 1, 1, 2, 3, 5, 8]) for _ in range(rnd.randint(1, 6))]
        run(items, trial)
    print(f"{n} random schedules over random item lists: no deadlock, no hazard, no parity aliasing")


if __name__ == "__main__":
    main()
