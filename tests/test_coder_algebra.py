"""CPU check of the algebra the fused coder kernel rests on (llcomp_b200/csrc/coder.cu), independent of any GPU.

The kernel does not run RangeEncoder::put + renorm_encoder (/root/reference/llcomp.hpp:33-89) decision by decision.
It splits them into
  * the range recurrence carried as y = x + 0xFF0000 with x = range*M + A (chain_step), and
  * the low / carry / byte side done per 256-decision block from the x values alone: lane-serial low increments,
    renormalisation events compacted, low at event j from the prefix sums E_j, bytes emitted a round of 32 events
    at a time unless a 0xFF byte has to be deferred (byte_side_lanes, shift_low, renorm_slow).
This file restates both forms in plain Python -- the reference one straight from the header, the kernel one
following coder.cu statement by statement -- and requires identical bytes on random and adversarial decision
sequences.  It pins the derivation; the GPU parity tests pin the implementation.
"""
import random

import pytest

M32 = 0xFFFFFFFF
BIAS = 0xFF0000
BLK, PER_LANE = 256, 8
HP_EMPTY = 0x100


# ---- the reference, llcomp.hpp:33-89 ---------------------------------------------------------------------
def reference_encode(bits, probs):
    out = bytearray()
    low, rng, held, pending = 0, 0xFF00, -1, 0

    def renorm():
        nonlocal low, rng, held, pending
        while rng < 0x100:
            if held < 0:
                held = low >> 8
            elif low <= 0xFF00:
                out.append(held)
                out.extend(b"\xff" * pending)
                pending = 0
                held = low >> 8
            elif low >= 0x10000:
                out.append((held + 1) & 0xFF)
                out.extend(b"\x00" * pending)
                pending = 0
                held = (low >> 8) & 0xFF
            else:
                pending += 1
            low = (low & 0xFF) << 8
            rng <<= 8

    for bit, p in zip(bits, probs):
        r1 = (rng * p) >> 8
        if not bit:
            rng -= r1
        else:
            low += rng - r1
            rng = r1
        renorm()
    rng = 0xFF; low += 0xFF; renorm()            # finish(), :75-81
    rng = 0xFF; renorm()
    return bytes(out)


# ---- the kernel's form -----------------------------------------------------------------------------------
def chain_x_values(bits, probs):
    """chain_step: y' = ((y >> 8) - 0xFF00) * (nz * (-255 M) + 256 M) + (A + bias), nz = y >> 24; returns x = y - bias."""
    y = (0xFF00 << 8) + BIAS
    xs = []
    for bit, p in zip(bits, probs):
        m, a = (p, 0) if bit else (256 - p, 255)            # queue entry: (M, A), coder.cu header comment
        nz = y >> 24
        assert nz in (0, 1)
        av = ((y >> 8) - (BIAS >> 8)) & M32
        mf = (nz * ((-255 * m) & M32) + (m << 8)) & M32
        y = (av * mf + a + BIAS) & M32
        xs.append((y - BIAS) & M32)
    return xs


class Tail:                                          # ByteTail of coder.cu
    def __init__(self):
        self.low, self.hp, self.out = 0, HP_EMPTY, bytearray()


def renorm_slow(t):
    held, pending = t.hp & 0xFF, t.hp >> 9
    if t.hp & HP_EMPTY:
        held = t.low >> 8
    elif t.low <= 0xFF00:
        t.out.append(held)
        t.out.extend(b"\xff" * pending)
        pending = 0
        held = t.low >> 8
    elif t.low >= 0x10000:
        t.out.append((held + 1) & 0xFF)
        t.out.extend(b"\x00" * pending)
        pending = 0
        held = (t.low >> 8) & 0xFF
    else:
        pending += 1
    t.hp = held | (pending << 9)


def shift_low(t):
    if t.hp < HP_EMPTY and ((t.low - 0xFF01) & M32) >= 0xFF:
        t.out.append((t.hp + (t.low >> 16)) & 0xFF)
        t.hp = (t.low >> 8) & 0xFF
    else:
        renorm_slow(t)
    t.low = (t.low & 0xFF) << 8


def byte_side_block(t, x_carry, xs, nodelta, cnt):
    """byte_side_lanes for one block: xs / nodelta hold cnt live decisions; returns the new x_carry."""
    x = list(xs) + [0x01000000] * (BLK - cnt)        # the rest of a last block is inert
    nd = list(nodelta) + [True] * (BLK - cnt)
    lane_sum, lane_events = [], []
    for lane in range(32):
        xp = x[PER_LANE * lane - 1] if lane else x_carry
        r = (xp & 0xFFFFFF00) if xp < 0x10000 else (xp >> 8)
        s, ev = 0, []
        for i in range(PER_LANE):
            xi = x[PER_LANE * lane + i]
            sh = xi >> 8
            if not nd[PER_LANE * lane + i]:
                s = (s + r - sh) & M32
            is_ev = xi < 0x10000
            r = (xi & 0xFFFFFF00) if is_ev else sh
            if is_ev:
                ev.append(s)
        lane_sum.append(s)
        lane_events.append(ev)
    new_carry = x[BLK - 1]
    e_tot = (t.low + sum(lane_sum)) & M32
    n_ev = sum(len(e) for e in lane_events)
    if n_ev == 0:
        t.low = e_tot
        return new_carry
    evl, before = [0, 0, 0], 0                       # E_-3..E_-1 = 0, then E_j in event order
    for lane in range(32):
        off = (t.low + before) & M32
        evl.extend((v + off) & M32 for v in lane_events[lane])
        before = (before + lane_sum[lane]) & M32
    low_last = e_last = 0
    for j0 in range(0, n_ev, 32):
        nv = min(32, n_ev - j0)
        low_j, low_p, e0s = [], [], []
        for lane in range(nv):
            e3, e2, e1, e0 = evl[j0 + lane:j0 + lane + 4]
            low_j.append(((((e1 - e2) & 0xFF) << 8) + ((e0 - e1) & M32)) & M32)
            low_p.append(((((e2 - e3) & 0xFF) << 8) + ((e1 - e2) & M32)) & M32)
            e0s.append(e0)
        defers = any(((v - 0xFF01) & M32) < 0xFF for v in low_j)
        if t.hp < HP_EMPTY and not defers:
            for lane in range(nv):
                held = t.hp if lane == 0 else (low_p[lane] >> 8) & 0xFF
                t.out.append((held + (low_j[lane] >> 16)) & 0xFF)
            t.hp = (low_j[nv - 1] >> 8) & 0xFF
        else:
            for v in low_j:
                t.low = v
                shift_low(t)
        low_last, e_last = low_j[nv - 1], e0s[nv - 1]
    t.low = (((low_last & 0xFF) << 8) + ((e_tot - e_last) & M32)) & M32
    return new_carry


def kernel_form_encode(bits, probs):
    xs = chain_x_values(bits, probs)
    t = Tail()
    x_carry = 0xFF00 << 8                            # pseudo-x whose successor range is the initial 0xFF00
    for b0 in range(0, len(bits), BLK):
        cnt = min(BLK, len(bits) - b0)
        x_carry = byte_side_block(t, x_carry, xs[b0:b0 + cnt], [not b for b in bits[b0:b0 + cnt]], cnt)
    t.low = (t.low + 0xFF) & M32                     # finish(): range = 0xFF both times -> exactly one shift each
    shift_low(t)
    shift_low(t)
    return bytes(t.out)


# ---- tests -----------------------------------------------------------------------------------------------
def test_chain_carry_reproduces_the_range_recurrence():
    rnd = random.Random(5)
    for _ in range(50):
        n = rnd.randrange(1, 600)
        bits = [rnd.randrange(2) for _ in range(n)]
        probs = [rnd.randrange(7, 248) for _ in range(n)]
        rng = 0xFF00
        for x, bit, p in zip(chain_x_values(bits, probs), bits, probs):
            r1 = (rng * p) >> 8                                       # llcomp.hpp:62-69
            after = r1 if bit else rng - r1
            m, a = (p, 0) if bit else (256 - p, 255)
            assert x == rng * m + a and (x >> 8) == after
            rng = after << 8 if after < 0x100 else after              # one renormalisation step (:55-56)
            assert 0x100 <= rng <= 0xFFFF


@pytest.mark.parametrize("seed", range(6))
def test_random_sequences(seed):
    rnd = random.Random(seed)
    for _ in range(40):
        n = rnd.choice([1, 2, 7, 8, 9, 255, 256, 257, 511, 513, rnd.randrange(1, 3000)])
        skew = rnd.choice([0.5, 0.9, 0.1, 0.99])
        bits = [1 if rnd.random() < skew else 0 for _ in range(n)]
        lo, hi = rnd.choice([(7, 247), (7, 20), (230, 247), (100, 160)])
        probs = [rnd.randrange(lo, hi + 1) for _ in range(n)]
        assert kernel_form_encode(bits, probs) == reference_encode(bits, probs), (seed, n, skew, lo, hi)


def test_adversarial_runs():
    cases = []
    for n in (1, 300, 2000):
        cases += [([1] * n, [7] * n), ([0] * n, [247] * n), ([1] * n, [247] * n), ([0] * n, [7] * n),
                  ([i & 1 for i in range(n)], [7 if i % 3 else 247 for i in range(n)])]
    for bits, probs in cases:
        assert kernel_form_encode(bits, probs) == reference_encode(bits, probs)


def test_deferred_ff_bytes_and_carries_are_exercised():
    """The byte side has two regimes; make sure the random search above is not only hitting the plain one: look for
    sequences whose reference run defers at least one byte (outstanding_count > 0) and resolves it both ways."""
    rnd = random.Random(99)
    seen_ff = seen_carry = 0
    for _ in range(400):
        n = rnd.randrange(200, 1200)
        bits = [rnd.randrange(2) for _ in range(n)]
        probs = [rnd.randrange(7, 248) for _ in range(n)]
        ref = reference_encode(bits, probs)
        assert kernel_form_encode(bits, probs) == ref
        seen_ff += b"\xff" in ref
        seen_carry += b"\x00" in ref
    assert seen_ff and seen_carry
