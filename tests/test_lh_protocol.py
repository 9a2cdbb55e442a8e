"""Model check of the hand-off protocol of the lead / helper latency kernel (csrc/poseidon_lh.cuh), on the CPU.

The kernel's warps exchange operands through shared memory under NAMED barriers: the producer `bar.arrive`s and goes on, the
consumer `bar.sync`s. Such a protocol can be right under the usual timing and still wrong: a lab build dead-locked when the lead
re-armed the x^4 barrier for round k + 1 before a helper (stalled on an L2 miss) had waited at round k — two phases of one barrier
merged (profiles/r02_latency_lab.md section 6). This test explores EVERY interleaving of the lead warp and the helper warps over
the kernel's sequence of barrier and shared-memory operations (a few rounds, two permutations — the structure repeats) and checks

  * no barrier ever receives a second arrival of a warp before its current phase completed (phases cannot merge),
  * nobody dead-locks, every warp reaches its end,
  * every shared-memory read sees exactly the version it expects, and no buffer is overwritten before all its readers have read.

It also shows the check has teeth: the protocol as first written (one barrier id and one buffer for x^4) is rejected.
The operation sequences below mirror k_hash_lh line by line; keep them in step with the kernel."""
import sys

import pytest

sys.setrecursionlimit(100000)

ROUNDS = 3   # odd, like the 57 partial rounds (round 56 of the first permutation and round 0 of the second share a parity)
PERMS = 2


def programs(n_helpers, x4_parity):
    """(lead program, [helper programs]); ops: ('sync', bar) ('arrive', bar) ('write', buf, version) ('read', buf, version)"""
    def x4(k):
        return k & 1 if x4_parity else 0

    lead = []
    for perm in range(PERMS):
        lead.append(("sync", "init"))
        lead += [("read", f"X{i}", ("init", perm)) for i in range(n_helpers)]
        for k in range(ROUNDS):
            lead += [("write", f"X{i}", (perm, k)) for i in range(n_helpers)]
            lead.append(("arrive", "x"))
            lead += [("write", f"X4{x4(k)}_{i}", (perm, k)) for i in range(n_helpers)]
            lead.append(("arrive", f"x4{x4(k)}"))
            lead.append(("sync", "yk"))
            lead += [("read", f"YK{i}", (perm, k)) for i in range(n_helpers)]
        lead += [("write", f"X{i}", (perm, ROUNDS)) for i in range(n_helpers)]
        lead.append(("arrive", "x"))
    helpers = []
    for i in range(n_helpers):
        h = []
        for perm in range(PERMS):
            h.append(("write", f"X{i}", ("init", perm)))
            h.append(("arrive", "init"))
            for k in range(ROUNDS):
                h.append(("sync", "x"))
                h.append(("read", f"X{i}", (perm, k)))
                h.append(("write", f"YK{i}", (perm, k)))
                h.append(("arrive", "yk"))
                h.append(("sync", f"x4{x4(k)}"))
                h.append(("read", f"X4{x4(k)}_{i}", (perm, k)))
            h.append(("sync", "x"))
            h.append(("read", f"X{i}", (perm, ROUNDS)))
        helpers.append(h)
    return lead, helpers


def explore(n_helpers, x4_parity):
    """DFS over all interleavings; returns None or the first violation found (a string)"""
    lead, helpers = programs(n_helpers, x4_parity)
    progs = [lead] + helpers
    nthreads = len(progs)
    need = nthreads  # every barrier phase: all warps take part (arrive or sync)

    def reader_set(buf, version):
        # X_i: written by the lead, read by helper i — or written by helper i ("init"), read by the lead; X4*_i: lead -> helper i; YK_i: helper i -> lead
        if buf.startswith("YK"):
            return frozenset([0])
        if buf.startswith("X4"):
            return frozenset([1 + int(buf.split("_")[1])])
        i = int(buf[1:])
        return frozenset([0]) if version[0] == "init" else frozenset([1 + i])

    seen = set()
    # state: (pcs, barriers, buffers)  barriers: tuple of (name, arrived frozenset, waiting frozenset); buffers: tuple of (name, version, unread frozenset)
    init = (tuple(0 for _ in progs), (), ())
    stack = [init]
    while stack:
        state = stack.pop()
        if state in seen:
            continue
        seen.add(state)
        pcs, bars, bufs = state
        bard = {b[0]: (b[1], b[2]) for b in bars}
        bufd = {b[0]: (b[1], b[2]) for b in bufs}
        blocked = set()
        for _, (arr, wait) in bard.items():
            blocked |= wait
        progressed = False
        done = all(pcs[t] == len(progs[t]) for t in range(nthreads))
        for t in range(nthreads):
            if pcs[t] == len(progs[t]) or t in blocked:
                continue
            op = progs[t][pcs[t]]
            nb, nf = dict(bard), dict(bufd)
            npcs = list(pcs)
            if op[0] in ("arrive", "sync"):
                arr, wait = nb.get(op[1], (frozenset(), frozenset()))
                if t in arr:
                    return f"warp {t} arrives twice at barrier {op[1]} within one phase (op {pcs[t]}): two phases merge"
                arr = arr | {t}
                if op[0] == "sync":
                    wait = wait | {t}
                if len(arr) == need:
                    arr, wait = frozenset(), frozenset()   # complete: everybody released, barrier re-initialised
                nb[op[1]] = (arr, wait)
                npcs[t] += 1
            elif op[0] == "write":
                old = nf.get(op[1])
                if old is not None and old[1]:
                    return f"warp {t} overwrites {op[1]} version {old[0]} before warps {sorted(old[1])} have read it"
                nf[op[1]] = (op[2], reader_set(op[1], op[2]))
                npcs[t] += 1
            else:  # read
                cur = nf.get(op[1])
                if cur is None or cur[0] != op[2]:
                    return f"warp {t} reads {op[1]} expecting version {op[2]} but finds {None if cur is None else cur[0]}"
                nf[op[1]] = (cur[0], cur[1] - {t})
                npcs[t] += 1
            progressed = True
            stack.append((tuple(npcs), tuple(sorted((k, v[0], v[1]) for k, v in nb.items() if v[0] or v[1])),
                          tuple(sorted((k, v[0], v[1]) for k, v in nf.items()))))
        if not progressed and not done:
            return f"dead-lock at program counters {pcs}"
    return None


@pytest.mark.parametrize("n_helpers", [1, 2, 3])   # the kernel runs 1 lead + 3 helper warps per block
def test_shipped_protocol_is_safe_under_every_interleaving(n_helpers):
    assert explore(n_helpers, x4_parity=True) is None


def test_single_x4_barrier_is_rejected():
    """the protocol as first written: one barrier id and one buffer for x^4 — the lead can re-arm it before a late helper waited"""
    why = explore(2, x4_parity=False)
    assert why is not None and ("twice" in why or "overwrites" in why or "dead-lock" in why), why
