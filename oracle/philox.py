"""CPU oracle: Philox4x32-10 and the draw layout of the device stretch move.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Philox4x32-10 is the
published counter-based generator of Salmon et al. (SC'11); the known-answer
vectors of the Random123 distribution are checked in
``tests/test_oracle_golden.py``.  The draw layout below is this repo's own
(documented in DESIGN.md, "K5") and is restated here independently of the
CUDA source so that a whole device chain can be replayed on the CPU.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32; counters broadcast, returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & MASK
                                           for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return [c.astype(np.uint32) for c in (c0, c1, c2, c3)]


def u53(hi, lo):
    """Two uint32 words -> double in [0, 1) with 53 random bits
    ((hi >> 5) * 2^26 + (lo >> 6)) / 2^53 — the classic genrand_res53 map."""
    a = (np.asarray(hi, dtype=np.uint64) >> np.uint64(5)).astype(np.float64)
    b = (np.asarray(lo, dtype=np.uint64) >> np.uint64(6)).astype(np.float64)
    return (a * 67108864.0 + b) / 9007199254740992.0


STREAM_SPLIT = 0      # c2 value: coin of walker pair (2i, 2i+1), counter c0 = id of walker 2i
STREAM_PARTNER = 1    # c2 value: partner draw of a walker
STREAM_MOVE = 2       # c2 value: z-uniform (words 0,1) and accept-uniform (words 2,3)


def pair_bits(seed, nwalkers, step, randomize_split=True, walker_offset=0):
    """Coin of every walker pair (2i, 2i+1); an unpaired last walker gets 0."""
    npair = (nwalkers + 1) // 2
    bits = np.zeros(npair, dtype=np.int64)
    if randomize_split:
        i = np.arange(npair, dtype=np.uint64)
        r = philox4x32(np.uint64(walker_offset) + np.uint64(2) * i, step, STREAM_SPLIT, 0,
                       seed & 0xFFFFFFFF, seed >> 32)
        bits = (r[0] & np.uint32(1)).astype(np.int64)
        if nwalkers % 2:
            bits[-1] = 0
    return bits


def split_sets(seed, nwalkers, step, randomize_split=True, walker_offset=0):
    """Set id (0/1) of every walker at ``step``.

    randomize_split=False: emcee's ``arange(nwalkers) % 2``.
    randomize_split=True : one fair coin per walker pair decides which of
    (2i, 2i+1) is red — a balanced, position-independent random partition
    (emcee shuffles a balanced label vector instead)."""
    bits = pair_bits(seed, nwalkers, step, randomize_split, walker_offset)
    w = np.arange(nwalkers)
    return (bits[w // 2] ^ (w & 1)).astype(np.int64)


def move_draws(seed, nwalkers, step, randomize_split=True, walker_offset=0):
    """All random draws of one ensemble step, exactly as the device makes them.

    Returns dict(sets, partner, u_z, u_acc): ``partner[w]`` is the LOCAL index
    of the complementary walker drawn for walker w, ``u_z`` the uniform behind
    z, ``u_acc`` the accept uniform."""
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    bits = pair_bits(seed, nwalkers, step, randomize_split, walker_offset)
    wl = np.arange(nwalkers)
    sets = (bits[wl // 2] ^ (wl & 1)).astype(np.int64)
    w = wl.astype(np.uint64) + np.uint64(walker_offset)
    r = philox4x32(w, step, STREAM_PARTNER, 0, k0, k1)
    n_other = np.where(sets == 0, nwalkers // 2, (nwalkers + 1) // 2)
    jj = ((r[0].astype(np.uint64) * n_other.astype(np.uint64)) >> np.uint64(32)).astype(np.int64)
    partner = 2 * jj + (bits[np.minimum(jj, len(bits) - 1)] ^ (1 - sets))
    partner = np.where(n_other > 0, partner, -1)
    r = philox4x32(w, step, STREAM_MOVE, 0, k0, k1)
    return dict(sets=sets, partner=partner, u_z=u53(r[0], r[1]), u_acc=u53(r[2], r[3]))
