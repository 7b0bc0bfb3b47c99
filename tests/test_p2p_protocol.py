"""Model check (CPU) of the fence-free peer-memory ghost exchange (k_halo_p2p, exsaddle_b200/csrc/xsb_comm.cu).

Protocol: every rank owns a window with `nslots` slots per direction, kept filled with a sentinel.  Exchange s of a rank is one
kernel: part A stores the rank's boundary planes into slot s % nslots of both neighbours' windows; part B reads slot s % nslots
of its OWN window from each neighbour, waiting per element until it is no longer the sentinel, then puts the sentinel back.
Kernels of one rank run in stream order; ranks are not synchronised otherwise.  Claims behind the implementation:
  (1) with TWO slots no store ever lands on data that has not been consumed (no acknowledgement needed);
  (2) every read returns the value of the matching exchange of the matching neighbour;
  (3) with ONE slot claim (1) fails -- the double buffer is necessary, and this test can see the hazard.
The simulation interleaves the ranks' elementary actions (single 8-byte stores / loads, as on the device) under many random schedules."""
import random

import pytest

SENT = None


def simulate(nranks, nexch, nelem, nslots, seed):
    rng = random.Random(seed)
    # window[r][side][slot][e]; side 0 = written by the neighbour below (r-1), side 1 = by the neighbour above (r+1)
    win = [[[[SENT] * nelem for _ in range(nslots)] for _ in range(2)] for _ in range(nranks)]
    overwrites = []; wrong = []

    def program(r):
        """generator of elementary actions of rank r; yields after every action ('ok') or when it must wait ('wait')"""
        for s in range(nexch):
            slot = s % nslots
            # part A: stores, element by element, in a random order over both neighbours (threads of the kernel)
            stores = [(nb, e) for nb in (r - 1, r + 1) if 0 <= nb < nranks for e in range(nelem)]
            rng.shuffle(stores)
            for nb, e in stores:
                side = 1 if nb == r - 1 else 0          # I am the neighbour ABOVE rank r-1, BELOW rank r+1
                if win[nb][side][slot][e] is not SENT:
                    overwrites.append((r, nb, s, e))
                win[nb][side][slot][e] = (r, s, e)
                yield "ok"
            # part B: loads with re-read, then the sentinel goes back
            loads = [(side, e) for side, nb in ((0, r - 1), (1, r + 1)) if 0 <= nb < nranks for e in range(nelem)]
            rng.shuffle(loads)
            for side, e in loads:
                while win[r][side][slot][e] is SENT:
                    yield "wait"
                got = win[r][side][slot][e]
                want = (r - 1 if side == 0 else r + 1, s, e)
                if got != want:
                    wrong.append((r, s, side, e, got, want))
                win[r][side][slot][e] = SENT
                yield "ok"

    progs = [program(r) for r in range(nranks)]
    alive = set(range(nranks)); idle_rounds = 0
    while alive:
        r = rng.choice(sorted(alive))
        burst = rng.choice((1, 1, 2, 5, 50))           # let one rank run ahead now and then
        progressed = False
        for _ in range(burst):
            try:
                if next(progs[r]) == "ok":
                    progressed = True
                else:
                    break
            except StopIteration:
                alive.discard(r); progressed = True
                break
        idle_rounds = 0 if progressed else idle_rounds + 1
        if idle_rounds > 20000:      # nobody can move: a consumed-too-early value is being waited for (only after an overwrite)
            return overwrites, wrong, True
    return overwrites, wrong, False


@pytest.mark.parametrize("nranks", [2, 3, 5])
def test_double_buffered_sentinel_exchange_never_overwrites_unread_data(nranks):
    for seed in range(40):
        overwrites, wrong, stuck = simulate(nranks, nexch=7, nelem=3, nslots=2, seed=seed)
        assert not overwrites and not wrong and not stuck, (seed, overwrites[:3], wrong[:3], stuck)


def test_single_buffer_would_be_overwritten():
    """the hazard the second slot removes: with one slot some schedule stores exchange s+1 onto unread data of exchange s"""
    hit = False
    for seed in range(200):
        overwrites, wrong, stuck = simulate(3, nexch=7, nelem=3, nslots=1, seed=seed)
        if overwrites or wrong or stuck:
            hit = True
            break
    assert hit
