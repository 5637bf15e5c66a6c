#!/usr/bin/env python
"""Randomised simulation of the mbarrier protocol between the auxiliary-tile producer (warp 3), the
MMA issuer and the two epilogue groups of `conv_gemm_kernel` (csrc/conv_gemm.cu).

The kernel's barriers are phase-parity mbarriers: `try_wait(parity)` cannot tell "this phase is
done" from "I am a whole phase ahead".  With ONE consumer per barrier program order rules the second
case out; with the two independent epilogue groups it does not, and the first version of the
two-group epilogue read an auxiliary stage that had not been filled yet (device trap on B200) when a
stage was shared by both groups.  This model replays the protocol under random interleavings and
checks, at every consume, that the stage really holds the tile the consumer expects, that the
producer never overwrites a stage still being read, and that nobody deadlocks.

    python tools/sim_epilogue_protocol.py [--runs 2000]

`simulate_inplace` models the lean inference epilogue that stages its result in the residual
landing slot (4 slots, handed back per thread after its bulk store has read them).

Variants: "main" = the shipped protocol (a consumer first waits for the RELEASE of the stage's
previous use, then for its fill); "no_prewait" = the earlier, broken one (fill wait only).

Findings (tests/test_protocol_sim.py pins them):
  * >= 2 auxiliary stages: both variants are hazard-free under every interleaving tried, even with
    the threads of a group running completely unsynchronised (the model has no group barriers);
  * 1 stage (the configuration that trapped on hardware): "no_prewait" fails at once; "main" is
    correct when a group behaves as one thread, but a thread that runs two store blocks ahead of a
    sibling's release can still alias phases -> the kernel must never be launched with a single
    auxiliary stage.  run_conv guarantees that: store blocks with two auxiliary tiles run on
    128-wide tiles (4 slots = 2 stages), everything else has >= 3 stages.
"""
import argparse
import random


class MBar:
    """mbarrier with an arrival count; phase = parity of the number of completed phases."""

    def __init__(self, count):
        self.count = count
        self.pending = count
        self.phase = 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0
        if self.pending == 0:
            self.pending = self.count
            self.phase ^= 1

    def done(self, parity):
        """try_wait.parity: true once the phase with this parity has completed."""
        return self.phase != parity


class ProtocolError(Exception):
    pass


def simulate(stages, blocks_per_tile, tiles, variant, rng, threads_per_group=2):
    """One random interleaving.  Returns the number of scheduler steps; raises ProtocolError."""
    n_blocks = tiles * blocks_per_tile
    rfull = [MBar(1) for _ in range(stages)]
    rempty = [MBar(threads_per_group) for _ in range(stages)]
    tfull = [MBar(1) for _ in range(2)]
    tempty = [MBar(2 * threads_per_group) for _ in range(2)]
    slot = [None] * stages          # block whose tiles the stage holds
    readers = [0] * stages          # consumer threads currently reading the stage

    def producer():
        for a in range(n_blocks):
            rs, phase = a % stages, (a // stages) & 1
            while not rempty[rs].done(phase ^ 1):
                yield
            if readers[rs]:
                raise ProtocolError(f"fill of block {a} overwrites stage {rs} while it is being read")
            slot[rs] = a
            yield                    # the TMA load is in flight for a while
            rfull[rs].arrive()       # complete_tx
            yield

    def mma():
        for t in range(tiles):
            acc, phase = t & 1, (t >> 1) & 1
            while not tempty[acc].done(phase ^ 1):
                yield
            yield
            tfull[acc].arrive()

    def consumer(group):
        for t in range(tiles):
            acc, phase = t & 1, (t >> 1) & 1
            while not tfull[acc].done(phase):
                yield
            for sb in range(blocks_per_tile):
                a = t * blocks_per_tile + sb
                if (a & 1) != group:
                    continue
                rs, use = a % stages, a // stages
                if variant == "main" and use > 0:
                    while not rempty[rs].done((use - 1) & 1):
                        yield
                while not rfull[rs].done(use & 1):
                    yield
                if slot[rs] != a:
                    raise ProtocolError(f"group {group} consumed stage {rs} holding block {slot[rs]} "
                                        f"instead of {a}")
                readers[rs] += 1
                yield                # residual add, pack, store
                if slot[rs] != a:
                    raise ProtocolError(f"stage {rs} changed under the reader of block {a}")
                readers[rs] -= 1
                rempty[rs].arrive()
                yield
            tempty[acc].arrive()
            yield

    actors = [producer(), mma()]
    for g in range(2):
        actors += [consumer(g) for _ in range(threads_per_group)]
    live = list(actors)
    steps, idle = 0, 0
    while live:
        a = rng.choice(live)
        before = (tuple(b.phase for b in rfull + rempty + tfull + tempty),
                  tuple(b.pending for b in rfull + rempty + tfull + tempty), tuple(slot), tuple(readers))
        try:
            next(a)
        except StopIteration:
            live.remove(a)
            idle = 0
            continue
        after = (tuple(b.phase for b in rfull + rempty + tfull + tempty),
                 tuple(b.pending for b in rfull + rempty + tfull + tempty), tuple(slot), tuple(readers))
        idle = idle + 1 if before == after else 0
        steps += 1
        if idle > 200 * len(live):
            raise ProtocolError("deadlock: no actor can make progress")
    return steps


def simulate_inplace(stages, blocks_per_tile, tiles, rng, threads_per_group=2):
    """The lean inference epilogue with a residual (conv_gemm.cu, LEAN && RES): store block a uses
    landing slot a % stages -- with an even number of slots always the same group's --, every
    thread writes its result over the residual bytes it has consumed, issues its own bulk store
    from the slot, and hands the slot back (one arrival per thread) right before it waits for the
    tile of its NEXT block, once that store has finished reading.  Checks: the producer never
    refills a slot that is being read by a thread or by a store in flight, every consume finds the
    expected tile, nobody deadlocks."""
    assert stages % 2 == 0, "a slot must always belong to the same epilogue group"
    n_blocks = tiles * blocks_per_tile
    rfull = [MBar(1) for _ in range(stages)]
    rempty = [MBar(threads_per_group) for _ in range(stages)]
    tfull = [MBar(1) for _ in range(2)]
    tempty = [MBar(2 * threads_per_group) for _ in range(2)]
    slot = [None] * stages
    readers = [0] * stages          # threads between their first read and their last write
    storing = [0] * stages          # bulk stores still reading the slot

    def producer():
        for a in range(n_blocks):
            rs, phase = a % stages, (a // stages) & 1
            while not rempty[rs].done(phase ^ 1):
                yield
            if readers[rs] or storing[rs]:
                raise ProtocolError(f"fill of block {a} overwrites slot {rs} while it is in use")
            slot[rs] = a
            yield
            rfull[rs].arrive()
            yield

    def mma():
        for t in range(tiles):
            acc, phase = t & 1, (t >> 1) & 1
            while not tempty[acc].done(phase ^ 1):
                yield
            yield
            tfull[acc].arrive()

    def consumer(group):
        held = None                  # slot this thread's last bulk store was issued from
        for t in range(tiles):
            acc, phase = t & 1, (t >> 1) & 1
            while not tfull[acc].done(phase):
                yield
            for sb in range(blocks_per_tile):
                a = t * blocks_per_tile + sb
                if (a & 1) != group:
                    continue
                rs, use = a % stages, a // stages
                if held is not None:
                    for _ in range(rng.randint(0, 3)):   # cp.async.bulk.wait_group.read
                        yield
                    storing[held] -= 1
                    rempty[held].arrive()
                held = rs
                while not rfull[rs].done(use & 1):
                    yield
                if slot[rs] != a:
                    raise ProtocolError(f"group {group} consumed slot {rs} holding block {slot[rs]} "
                                        f"instead of {a}")
                readers[rs] += 1
                yield                # residual read, math, in-place write
                if slot[rs] != a:
                    raise ProtocolError(f"slot {rs} changed under the reader of block {a}")
                readers[rs] -= 1
                storing[rs] += 1     # the bulk store reads the slot from now on
                yield
            tempty[acc].arrive()
            yield
        # (the last store of a thread is drained by cp.async.bulk.wait_group 0 at kernel end)

    actors = [producer(), mma()]
    for g in range(2):
        actors += [consumer(g) for _ in range(threads_per_group)]
    live = list(actors)
    steps, idle = 0, 0
    bars = rfull + rempty + tfull + tempty
    while live:
        a = rng.choice(live)
        before = (tuple(b.phase for b in bars), tuple(b.pending for b in bars), tuple(slot),
                  tuple(readers), tuple(storing))
        try:
            next(a)
        except StopIteration:
            live.remove(a)
            idle = 0
            continue
        after = (tuple(b.phase for b in bars), tuple(b.pending for b in bars), tuple(slot),
                 tuple(readers), tuple(storing))
        idle = idle + 1 if before == after else 0
        steps += 1
        if idle > 400 * len(live):
            raise ProtocolError("deadlock: no actor can make progress")
    return steps


def check(variant, runs, seed=0):
    """Returns {(stages, blocks_per_tile): first error or None}."""
    rng = random.Random(seed)
    out = {}
    for stages in (1, 2, 3, 4):
        for bpt in (1, 2, 4):
            err = None
            for _ in range(runs):
                try:
                    simulate(stages, bpt, tiles=rng.randint(3, 9), variant=variant, rng=rng)
                except ProtocolError as e:
                    err = str(e)
                    break
            out[(stages, bpt)] = err
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=2000)
    args = ap.parse_args()
    for variant in ("main", "no_prewait"):
        print(variant)
        for key, err in check(variant, args.runs).items():
            print(f"  stages {key[0]} blocks/tile {key[1]}: {'ok' if err is None else err}")
    print("inplace (lean residual epilogue)")
    rng = random.Random(1)
    for stages in (2, 4):
        for bpt in (1, 2, 4):
            err = None
            for _ in range(args.runs):
                try:
                    simulate_inplace(stages, bpt, rng.randint(3, 9), rng)
                except ProtocolError as e:
                    err = str(e)
                    break
            print(f"  slots {stages} blocks/tile {bpt}: {'ok' if err is None else err}")
