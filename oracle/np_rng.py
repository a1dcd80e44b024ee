"""Integer restatement of numpy's Generator(PCG64(SeedSequence(seed))) — the stream behind
gymnasium.utils.seeding.np_random(seed), which the reference's reset samplers draw from
(hole_reacher.py:79-112, viapoint_reacher.py:55-77, simple_reacher.py:85-96, base_reacher.py:73-93).

TEST INFRASTRUCTURE (see oracle/__init__.py): the specification of the device-side sampler
(fancy_gym_b200/csrc/fg_reset.cuh), itself pinned against numpy in tests/test_np_rng.py.

Algorithms (numpy/random: bit_generator.pyx SeedSequence, src/pcg64, src/distributions):
  SeedSequence: 4-word uint32 pool, hashmix / mix with the INIT_A/MULT_A, INIT_B/MULT_B, MIX_MULT_L/R constants
  PCG64      : 128-bit LCG (multiplier 0x2360ED051FC65DA44385DF649FCCF645), output XSL-RR 128/64, step THEN output
  next_double: (next64 >> 11) * 2**-53;   uniform(lo, hi) = lo + (hi - lo) * next_double
  next_uint32: low half of a 64-bit draw, the high half is buffered for the following call
  integers(0, 2) (behind choice([-1, 1])): Lemire on one uint32 -> its top bit
"""
M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
M128 = (1 << 128) - 1
INIT_A, MULT_A, INIT_B, MULT_B = 0x43B0D7E5, 0x931E8875, 0x8B51F9DD, 0x58F38DED
MIX_L, MIX_R, XSHIFT = 0xCA01F9DD, 0x4973F715, 16
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645


def seed_sequence_state(seed: int):
    """SeedSequence(seed).generate_state(4, uint64) for a non-negative int seed"""
    ent = []
    s = int(seed)
    while True:
        ent.append(s & M32)
        s >>= 32
        if not s:
            break
    hc = [INIT_A]

    def hashmix(v):
        v = (v ^ hc[0]) & M32
        hc[0] = (hc[0] * MULT_A) & M32
        v = (v * hc[0]) & M32
        return v ^ (v >> XSHIFT)

    def mix(x, y):
        r = (MIX_L * x - MIX_R * y) & M32
        return r ^ (r >> XSHIFT)

    pool = [hashmix(ent[i] if i < len(ent) else 0) for i in range(4)]
    for i_src in range(4):
        for i_dst in range(4):
            if i_src != i_dst:
                pool[i_dst] = mix(pool[i_dst], hashmix(pool[i_src]))
    for i_src in range(4, len(ent)):
        for i_dst in range(4):
            pool[i_dst] = mix(pool[i_dst], hashmix(ent[i_src]))
    hb = INIT_B
    out32 = []
    for i in range(8):
        v = pool[i % 4] ^ hb
        hb = (hb * MULT_B) & M32
        v = (v * hb) & M32
        out32.append(v ^ (v >> XSHIFT))
    return [out32[2 * j] | (out32[2 * j + 1] << 32) for j in range(4)]


class PCG64:
    def __init__(self, seed: int):
        s = seed_sequence_state(seed)
        initstate = (s[0] << 64) | s[1]
        initseq = (s[2] << 64) | s[3]
        self.inc = ((initseq << 1) | 1) & M128
        self.state = 0
        self._step()
        self.state = (self.state + initstate) & M128
        self._step()
        self.has_uint32, self.uinteger = 0, 0

    def _step(self):
        self.state = (self.state * PCG_MULT + self.inc) & M128

    def next64(self):
        self._step()
        hi, lo = self.state >> 64, self.state & M64
        x, rot = hi ^ lo, self.state >> 122
        return ((x >> rot) | (x << ((-rot) & 63))) & M64

    def next32(self):
        if self.has_uint32:
            self.has_uint32 = 0
            return self.uinteger
        n = self.next64()
        self.has_uint32, self.uinteger = 1, n >> 32
        return n & M32

    def next_double(self):
        return (self.next64() >> 11) * (1.0 / 9007199254740992.0)

    def uniform(self, low, high):
        return low + (high - low) * self.next_double()

    def choice2(self):
        """index drawn by Generator.choice over a 2-element population"""
        return self.next32() >> 31
