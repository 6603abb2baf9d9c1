"""Model check of the in-kernel group barrier protocol (DTFFTB_FUSED_SYNC, dtfft_b200/csrc/kernels.cu:
fused_sync_enter / fused_sync_leave; stand-alone form: peer.cu: peer_barrier_kernel) with Python threads
standing in for CTAs: same flags / epochs / tickets, same order of operations, random interleavings.
Checks what the CUDA code must guarantee and what a desk check can get wrong:
  * nobody stores before every member has announced "my destination is free" for this launch;
  * a launch returns (its last CTA leaves) only after every member's every CTA has finished storing;
  * tickets and epochs are ready for the next launch; a member that uses the two stand-alone barrier
    kernels instead (nothing to store) pairs up with folded members, launch after launch.
This is a model of the protocol, not of the device code (which needs a GPU: tools/r02_n2.sh step 5)."""
import random
import threading
import time

P, CTAS, LAUNCHES = 3, 4, 12
ROW_FREE, ROW_LANDED = 0, 1


class Rank:
    def __init__(self, me):
        self.me = me
        self.flags = [[0] * P for _ in range(2)]   # [row][writer]: written by peers (st.release.sys)
        self.epoch = [0, 0]                          # group epoch counters of the two channels
        self.tickets = [0, 0]
        self.lock = threading.Lock()                 # atomicAdd
        self.started = [0] * (LAUNCHES + 1)          # model state for the checks
        self.stored = [0] * (LAUNCHES + 1)


def jitter(rng):
    if rng.random() < 0.3:
        time.sleep(rng.random() * 0.002)


def wait(flag_row, writer, epoch, deadline):
    while flag_row[writer] < epoch:
        assert time.time() < deadline, "protocol deadlock"
        time.sleep(0)


def folded_cta(ranks, r, launch, rng, errors, deadline):
    me = ranks[r]
    try:
        # ---- fused_sync_enter
        with me.lock:
            first = me.tickets[0] == 0
            me.tickets[0] += 1
        epoch = me.epoch[ROW_FREE] + 1
        jitter(rng)
        if first:
            me.started[launch] = 1
            for peer in ranks:
                peer.flags[ROW_FREE][r] = epoch
        for w in range(P):
            wait(me.flags[ROW_FREE], w, epoch, deadline)
        # ---- the stores: every member must have started this launch (its destination is free)
        assert all(x.started[launch] for x in ranks), "stored before every member's destination was free"
        jitter(rng)
        with me.lock:
            me.stored[launch] += 1
        # ---- fused_sync_leave
        with me.lock:
            last = me.tickets[1] == CTAS - 1
            me.tickets[1] += 1
        if not last:
            return
        epoch_l = me.epoch[ROW_LANDED] + 1
        for peer in ranks:
            peer.flags[ROW_LANDED][r] = epoch_l
        for w in range(P):
            wait(me.flags[ROW_LANDED], w, epoch_l, deadline)
        # everybody's blocks have landed: all CTAs of all storing members are done with this launch
        for x in ranks:
            assert x.stored[launch] == (CTAS if x.folded else 0), "left before every block had landed"
        me.epoch[ROW_FREE] += 1
        me.epoch[ROW_LANDED] = epoch_l
        me.tickets = [0, 0]
    except AssertionError as e:
        errors.append(str(e))


def barrier_kernel(ranks, r, row, deadline):
    """peer_barrier_kernel: one block, thread t handles member t."""
    me = ranks[r]
    me.epoch[row] += 1
    epoch = me.epoch[row]
    for peer in ranks:
        peer.flags[row][r] = epoch
    for w in range(P):
        wait(me.flags[row], w, epoch, deadline)


def run_rank(ranks, r, seed, errors, deadline):
    rng = random.Random(seed)
    me = ranks[r]
    for launch in range(1, LAUNCHES + 1):
        jitter(rng)
        if me.folded:
            ctas = [threading.Thread(target=folded_cta, args=(ranks, r, launch, random.Random(rng.random()), errors, deadline))
                    for _ in range(CTAS)]
            for t in ctas:
                t.start()
            for t in ctas:
                t.join()   # kernel boundary: the next launch on the stream starts after this one
        else:              # nothing to store: barrier kernel, (no kernel), barrier kernel
            try:
                me.started[launch] = 1
                barrier_kernel(ranks, r, ROW_FREE, deadline)
                barrier_kernel(ranks, r, ROW_LANDED, deadline)
                for x in ranks:
                    assert x.stored[launch] == (CTAS if x.folded else 0), "barrier member left before every block had landed"
            except AssertionError as e:
                errors.append(str(e))
        if errors:
            return


def test_folded_barriers_pair_up_with_stand_alone_ones():
    for trial, folded in enumerate(([True, True, True], [True, False, True], [False, True, True])):
        ranks = [Rank(i) for i in range(P)]
        for x, f in zip(ranks, folded):
            x.folded = f
        errors = []
        deadline = time.time() + 60
        ts = [threading.Thread(target=run_rank, args=(ranks, r, 100 * trial + r, errors, deadline)) for r in range(P)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errors, errors[:3]
        for x in ranks:
            assert x.epoch == [LAUNCHES, LAUNCHES] and (x.tickets == [0, 0])
