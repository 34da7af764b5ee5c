"""Exhaustive interleaving check of the peer-memory exchange protocol of csrc/p2p_update.cu (host-side model, no GPU).

The kernels implement, per rank r and step e = 1, 2, ... on one in-order stream:
    publish(e)   export_r[e & 1] <- gradient of step e
    signal(e)    flags_p[r] <- e              for every rank p (one store each, any order between ranks)
    wait(e)      block until flags_r[p] >= e  for every rank p
    adam(e)      read export_p[e & 1] for every rank p       (must see step e's gradient of every rank)
DESIGN.md §6 claims the double buffer makes a second barrier unnecessary: nobody overwrites a slot a peer may still read.
This test explores EVERY interleaving of the ranks' operations (each store / each read its own atomic event) for 2 ranks x 8
steps, 3 ranks x 3 steps and 4 ranks x 3 steps and asserts that every read returns the epoch it expects — and that the same protocol with a
single export buffer is caught violating it, so the check has teeth.
"""
from collections import deque


def _program(rank, world, steps):
    ops = []
    for e in range(1, steps + 1):
        ops.append(("publish", e))
        for p in range(world):
            ops.append(("signal", e, p))
        ops.append(("wait", e))
        for p in range(world):
            ops.append(("read", e, p))
    return ops


def _explore(world, steps, slots):
    progs = [_program(r, world, steps) for r in range(world)]
    # state: (pc per rank, export[r][slot] epoch, flags[p][r])
    init = (tuple(0 for _ in range(world)),
            tuple(tuple(0 for _ in range(slots)) for _ in range(world)),
            tuple(tuple(0 for _ in range(world)) for _ in range(world)))
    seen, todo = {init}, deque([init])
    violations, finals = 0, 0
    while todo:
        pcs, export, flags = todo.popleft()
        done = True
        for r in range(world):
            if pcs[r] >= len(progs[r]):
                continue
            done = False
            op = progs[r][pcs[r]]
            ex, fl = export, flags
            if op[0] == "publish":
                e = op[1]
                ex = tuple(tuple(e if (i == r and s == e % slots) else v for s, v in enumerate(row)) for i, row in enumerate(export))
            elif op[0] == "signal":
                e, p = op[1], op[2]
                fl = tuple(tuple(e if (i == p and j == r) else v for j, v in enumerate(row)) for i, row in enumerate(flags))
            elif op[0] == "wait":
                if any(flags[r][p] < op[1] for p in range(world)):
                    continue  # blocked: not enabled in this state
            elif op[0] == "read":
                e, p = op[1], op[2]
                if export[p][e % slots] != e:
                    violations += 1
                    continue  # do not explore past a violation
            nxt = (tuple(pc + 1 if i == r else pc for i, pc in enumerate(pcs)), ex, fl)
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
        if done:
            finals += 1
    return violations, finals, len(seen)


def test_double_buffered_exchange_is_safe_under_every_interleaving():
    for world, steps in ((2, 8), (3, 3), (4, 3)):
        violations, finals, states = _explore(world, steps, slots=2)
        assert violations == 0 and finals >= 1, (world, steps, violations, finals, states)


def test_single_buffer_would_race():
    violations, _, _ = _explore(2, 3, slots=1)
    assert violations > 0
