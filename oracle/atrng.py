"""ATRNG - the counter-based random stream shared by the oracle, the reference
harness and the CUDA kernels.  TEST INFRASTRUCTURE (numpy restatement): only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

The reference draws from CPython's MT19937 (``random``) and numpy's legacy
MT19937 (``np.random``) without ever seeding them, so it
has no reproducible stream of its own.  Parity is therefore defined on a
stream that all three implementations can produce exactly:

    word(seed, env, stream, ctr) = Philox4x32-10(counter=(ctr, stream, env_lo, env_hi),
                                                 key=(seed_lo, seed_hi))[0]

* stream 0  : env dynamics (reset draws, weakness coin, random/dodge bots);
              ``ctr`` is a per-env draw counter kept in the env state.
* stream 1+a: synthetic action stream of agent ``a``; ``ctr`` = tick index,
              words 0..2 give (rot, mov, shoot).

Draw mappings (the harness serves these in place of the reference's
``random.*`` / ``np.random.*`` calls, see tests/refharness.py):

    below(n)        = (word * n) >> 32                          one word
    randint(a, b)   = a + below(b - a + 1)                      random.randint  (env/sprite.py:220)
    choice(n)       = below(n)                                  np.random.choice(range(n)) (env/gaming_env.py:512)
    sample2(n)      = i = below(n); j = below(n-1); j += j >= i random.sample(pop, 2) (env/gaming_env.py:59-60)
    unit53()        = ((w0 >> 5) * 2**26 + (w1 >> 6)) / 2**53   np.random.random() (env/gaming_env.py:144), two words
    uniform(a, b)   = a + (b - a) * unit53()                    random.uniform (env/bots/simple_bots.py:378)
    shuffle(x)      = for i = len-1..1: j = below(i+1); swap    random.shuffle (env/bots/simple_bots.py:373)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-R.  All args broadcastable uint32-valued arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)]
    k0 = np.asarray(k0, dtype=np.uint64) & MASK
    k1 = np.asarray(k1, dtype=np.uint64) & MASK
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    for r in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def word(seed, env, stream, ctr):
    seed = int(seed)
    env = np.asarray(env, dtype=np.uint64)
    w = philox4x32(ctr, stream, env & MASK, env >> np.uint64(32),
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return w[0]


def below_from_word(w, n):
    return (np.asarray(w, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)


def synthetic_actions(seed, env_ids, tick, n_agents):
    """int32[len(env_ids), n_agents, 3] uniform actions on {0,1,2}x{0,1,2}x{0,1}
    for one tick (stream 1+a, ctr = tick).  Same function as the CUDA
    at_synthetic_actions kernel."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    out = np.empty((env_ids.shape[0], n_agents, 3), dtype=np.int32)
    seed = int(seed)
    for a in range(n_agents):
        w = philox4x32(tick, 1 + a, env_ids & MASK, env_ids >> np.uint64(32),
                       seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        out[:, a, 0] = below_from_word(w[0], 3)
        out[:, a, 1] = below_from_word(w[1], 3)
        out[:, a, 2] = below_from_word(w[2], 2)
    return out


class Stream:
    """Scalar stream-0 generator for one env (the draw counter is the state)."""

    def __init__(self, seed, env_id, ctr=0):
        self.seed, self.env_id, self.ctr = int(seed), int(env_id), int(ctr)

    def u32(self):
        w = int(word(self.seed, self.env_id, 0, self.ctr))
        self.ctr = (self.ctr + 1) & 0xFFFFFFFF
        return w

    def below(self, n):
        return (self.u32() * int(n)) >> 32

    def randint(self, a, b):
        return a + self.below(b - a + 1)

    def unit53(self):
        a = self.u32() >> 5
        b = self.u32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0

    def uniform(self, a, b):
        return a + (b - a) * self.unit53()

    def sample2(self, n):
        i = self.below(n)
        j = self.below(n - 1)
        if j >= i:
            j += 1
        return i, j
