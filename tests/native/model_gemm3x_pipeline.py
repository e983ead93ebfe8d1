"""Discrete model of the barrier protocol of gemm3x_kernel / wgrad3x_kernel (tzk_gemm3x.cu) — round-2 groundwork.

The kernels are warp-specialised: TMA producer, MMA issuer, four transform/epilogue warps, tied together by mbarriers
(full / ready / empty per shared-memory stage, acc_full / acc_empty per TMEM accumulator).  A wrong initial parity, a
missing arrival or a count that is off by one shows up on hardware as a hang or as silently stale data.  This model
restates each role as a coroutine that follows the kernel's control flow line by line (same loop nests, same
stage / phase / accumulator bookkeeping, same arrival counts), runs them under a random scheduler with an asynchronous,
in-order tensor-core queue (an MMA reads shared memory when it EXECUTES, a commit arrives when everything issued before
it has executed), and checks that

  * nobody deadlocks,
  * every MMA finds in its stage exactly the chunk (tile, k-block) it was issued for, already split into hi / lo,
  * every epilogue finds in its accumulator exactly the k-blocks of its own tile, each once,
  * the TMA never overwrites a stage that still has unexecuted MMAs or untransformed data pending.

mbarrier semantics modelled: `pending` arrivals per phase (+ outstanding transaction bytes); the phase completes when
both reach zero, flips the parity bit and re-arms; try_wait.parity(P) succeeds iff the current parity != P (so waiting
on parity 1 of a fresh barrier falls through — the producers' first pass over the empty barriers relies on that).

    python tests/native/model_gemm3x_pipeline.py        # a few thousand random schedules
"""
import random
from collections import deque


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.parity = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.parity ^= 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier was initialised for"
        self.pending -= 1
        self._maybe_flip()

    def expect_tx(self, nbytes):          # arrive.expect_tx: one arrival + outstanding bytes
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        assert self.tx >= 0
        self._maybe_flip()

    def done(self, parity):               # try_wait.parity
        return self.parity != parity


class Deadlock(Exception):
    pass


def simulate(num_tiles_for_cta, num_k, stages, n_acc, seed, n_epi=4):
    """One CTA processing `num_tiles_for_cta` tiles of `num_k` k-blocks (gemm3x: n_acc = 2; wgrad3x: one tile, n_acc = 1,
    no acc_empty)."""
    rng = random.Random(seed)
    full = [MBar(1) for _ in range(stages)]
    ready = [MBar(n_epi) for _ in range(stages)]
    empty = [MBar(1) for _ in range(stages)]
    acc_full = [MBar(1) for _ in range(n_acc)]
    acc_empty = [MBar(n_epi) for _ in range(n_acc)]
    slot = [None] * stages                # ("raw" | "split", tile, kb)
    slot_pending_mma = [0] * stages       # issued but unexecuted MMAs reading the stage
    acc = [[] for _ in range(n_acc)]      # k-blocks accumulated (tile, kb)
    tc_queue = deque()                    # in-order asynchronous tensor-core work: ("mma", stage, tile, kb, acc, first) | ("commit", bar)
    tma_queue = []                        # outstanding TMA loads: (stage, tile, kb) — complete in any order
    X = 100                               # bytes per chunk (any positive number)
    log = {"epilogues": 0}

    def wait(bar, parity):
        while not bar.done(parity):
            yield

    def tma_producer():
        stage, phase = 0, 0
        for t in range(num_tiles_for_cta):
            for kb in range(num_k):
                yield from wait(empty[stage], phase ^ 1)
                assert slot_pending_mma[stage] == 0, "TMA overwrites a stage with MMAs still to execute"
                assert slot[stage] is None or slot[stage][0] == "consumed", f"TMA overwrites live data {slot[stage]}"
                full[stage].expect_tx(X)
                tma_queue.append((stage, t, kb))
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield

    def mma_issuer():
        stage, phase, a, acc_phase = 0, 0, 0, 0
        for t in range(num_tiles_for_cta):
            if n_acc > 1:
                yield from wait(acc_empty[a], acc_phase ^ 1)
            for kb in range(num_k):
                yield from wait(ready[stage], phase)
                tc_queue.append(("mma", stage, t, kb, a, kb == 0))
                slot_pending_mma[stage] += 1
                tc_queue.append(("commit", empty[stage], stage))
                if kb == num_k - 1:
                    tc_queue.append(("commit", acc_full[a], None))
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield
            if n_acc > 1:
                a ^= 1
                if a == 0:
                    acc_phase ^= 1

    transformed = {}                       # (stage, fill#) -> warps done, to mark the slot split when all four passed

    def transform_epilogue(w):
        stage, phase, a, acc_phase = 0, 0, 0, 0
        fills = [0] * stages
        for t in range(num_tiles_for_cta):
            for kb in range(num_k):
                yield from wait(full[stage], phase)
                assert slot[stage] is not None and slot[stage][1:] == (t, kb), f"transform sees {slot[stage]}, wants {(t, kb)}"
                key = (stage, fills[stage])
                transformed[key] = transformed.get(key, 0) + 1
                if transformed[key] == n_epi:
                    slot[stage] = ("split", t, kb)
                fills[stage] += 1
                ready[stage].arrive()
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield
            yield from wait(acc_full[a], acc_phase)
            got = sorted(acc[a])
            assert got == [(t, kb) for kb in range(num_k)], f"epilogue of tile {t} finds {got}"
            yield                           # tcgen05.ld ... stores
            if n_acc > 1:
                done_key = ("acc", a, t)
                transformed[done_key] = transformed.get(done_key, 0) + 1
                if transformed[done_key] == n_epi:
                    acc[a] = []             # (the next tile's first MMA overwrites: accumulate flag = 0)
                    log["epilogues"] += 1
                acc_empty[a].arrive()
                a ^= 1
                if a == 0:
                    acc_phase ^= 1
            else:
                log["epilogues"] += 1 if w == 0 else 0

    def tensor_core_step():
        op = tc_queue.popleft()
        if op[0] == "mma":
            _, stage, t, kb, a, first = op
            assert slot[stage] == ("split", t, kb), f"MMA({t},{kb}) reads stage {stage} holding {slot[stage]}"
            if first:
                assert n_acc == 1 or acc[a] == [], f"tile {t} starts on an accumulator still holding {acc[a]}"
                acc[a] = []
            acc[a].append((t, kb))
            slot_pending_mma[stage] -= 1
            slot[stage] = ("consumed", t, kb)
        else:
            op[1].arrive()

    def tma_step():
        stage, t, kb = tma_queue.pop(rng.randrange(len(tma_queue)))
        slot[stage] = ("raw", t, kb)
        full[stage].complete_tx(X)

    threads = [tma_producer(), mma_issuer()] + [transform_epilogue(w) for w in range(n_epi)]
    alive = list(range(len(threads)))
    idle_rounds = 0
    while alive or tc_queue or tma_queue:
        choices = [("t", i) for i in alive]
        if tc_queue:
            choices.append(("tc", None))
        if tma_queue:
            choices.append(("tma", None))
        kind, i = rng.choice(choices)
        before = (tuple((b.pending, b.tx, b.parity) for b in full + ready + empty + acc_full + acc_empty), len(tc_queue),
                  len(tma_queue), len(alive))
        if kind == "tc":
            tensor_core_step()
        elif kind == "tma":
            tma_step()
        else:
            try:
                next(threads[i])
            except StopIteration:
                alive.remove(i)
        after = (tuple((b.pending, b.tx, b.parity) for b in full + ready + empty + acc_full + acc_empty), len(tc_queue),
                 len(tma_queue), len(alive))
        idle_rounds = idle_rounds + 1 if before == after else 0
        if idle_rounds > 20000:
            raise Deadlock(f"no progress: alive roles {alive}, tc queue {len(tc_queue)}, tma queue {len(tma_queue)}")
    assert log["epilogues"] == num_tiles_for_cta, (log, num_tiles_for_cta)
    return True


CASES = [
    # (tiles per CTA, k-blocks, stages, accumulators)       which kernel
    (4, 25, 4, 2),    # forward 784 -> 64: 512 tiles over 148 CTAs = 3-4 per CTA, 25 chunks of 32, 4 stages
    (1, 25, 4, 2),
    (7, 2, 3, 2),     # dgrad: K = 64 -> 2 chunks per tile, 3 stages (BN = 112), many tiles per CTA
    (5, 1, 3, 2),     # degenerate: one chunk per tile
    (3, 3, 4, 2),     # fewer chunks than stages
    (1, 98, 4, 1),    # wgrad: one work item per CTA, 98 chunks of 32 rows, single accumulator
    (1, 1, 4, 1),
]


def main(schedules=300):
    for case in CASES:
        for seed in range(schedules):
            simulate(*case, seed=seed)
        print(f"tiles/CTA {case[0]:>2}, k-blocks {case[1]:>2}, stages {case[2]}, accumulators {case[3]}: "
              f"{schedules} random schedules, no deadlock, every MMA / epilogue saw its own data")


if __name__ == "__main__":
    main()
