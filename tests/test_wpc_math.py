"""The arithmetic behind hash_chain_kernel_wpc, restated lane by lane in Python and checked against plain FNV-1a:
  (1) FNV-1a is a T-function, so the low byte of the running state resolves in 8 prefix-XOR rounds over the stream
      (one warp vote + popcount per round, 3 consecutive stream positions per lane);
  (2) h ^ b = h + e with e = z - l, so  h_m = P^m * (h_0 + sum_i e_i * Q^i),  Q = P^-1 mod 2^64, summed over the
      warp in four 16-bit limbs.
CPU-only: a change to the kernel's math can be tried here before it costs GPU time."""
import random

M = (1 << 64) - 1
P = 0x100000001B3
Q = pow(P, -1, 1 << 64)
H0 = 0xCBF29CE484222325


def fnv(stream):
    h = H0
    for b in stream:
        h = ((h ^ b) * P) & M
    return h


def wpc(stream):
    m = len(stream)
    assert m <= 96
    b = [[stream[3 * L + j] if 3 * L + j < m else 0 for j in range(3)] for L in range(32)]
    z = [row[:] for row in b]                      # z starts as b: bit k of z*0xb3 is then b_k ^ c_k
    for k in range(8):
        mask = 1 << k
        p = [[(z[L][j] * 0xB3) & 0xFFFFFFFF for j in range(3)] for L in range(32)]
        votes = sum((((p[L][0] ^ p[L][1] ^ p[L][2]) >> k) & 1) << L for L in range(32))      # __ballot_sync
        for L in range(32):
            before = bin(votes & ((1 << L) - 1)).count("1")                                   # popc(votes & lanemask_lt)
            m0 = mask if (H0 & 0xFF) & mask else 0
            zc = [z[L][0] ^ m0, z[L][1] ^ m0 ^ (p[L][0] & mask), z[L][2] ^ m0 ^ ((p[L][0] ^ p[L][1]) & mask)]
            sh = before << k
            z[L] = [zc[j] ^ (sh & mask) for j in range(3)]
    limbs = [0, 0, 0, 0]
    for L in range(32):
        acc, hi = 0, 0
        for j in range(3):
            q = pow(Q, 3 * L + j, 1 << 64)
            qlo = q & 0xFFFFFFFF
            qlo_s = qlo - (1 << 32) if qlo >= 1 << 31 else qlo          # q = (qhi + [qlo < 0]) * 2^32 + signed qlo
            qhi = ((q >> 32) + (1 if qlo_s < 0 else 0)) & 0xFFFFFFFF
            e = z[L][j] - (z[L][j] ^ b[L][j])                           # e = z - l, l = z ^ b
            acc += e * qlo_s
            hi = (hi + e * qhi) & 0xFFFFFFFF
        t = (acc + (hi << 32) + (H0 if L == 0 else 0)) & M             # lane 0 carries the offset basis
        for c in range(4):
            limbs[c] += (t >> (16 * c)) & 0xFFFF                        # redux.sync.add of 16-bit limbs
    total = sum(limbs[c] << (16 * c) for c in range(4)) & M
    return (pow(P, m, 1 << 64) * total) & M


def test_vote_round_formulation_equals_fnv1a():
    rng = random.Random(3)
    for _ in range(400):
        stream = [rng.randrange(256) for _ in range(rng.randrange(0, 97))]
        assert wpc(stream) == fnv(stream)


def test_block_streams_of_the_kernel_shape():
    """83 1b P7..P0 90 | 16 tokens in 1/2/3/5-byte CBOR forms | f6  — the streams the kernel folds (<= 92 bytes)."""
    rng = random.Random(4)
    for _ in range(200):
        parent = rng.getrandbits(64) | (1 << 32)
        stream = [0x83, 0x1B] + list(parent.to_bytes(8, "big")) + [0x90]
        for _ in range(16):
            t = rng.choice([rng.randrange(24), rng.randrange(24, 256), rng.randrange(256, 65536), rng.randrange(65536, 1 << 32)])
            if t < 24:
                stream += [t]
            elif t < 256:
                stream += [0x18, t]
            elif t < 65536:
                stream += [0x19] + list(t.to_bytes(2, "big"))
            else:
                stream += [0x1A] + list(t.to_bytes(4, "big"))
        stream += [0xF6]
        assert len(stream) <= 92 and wpc(stream) == fnv(stream)


def test_three_redux_limb_split_is_exact():
    """Variant tried in round 1 (correct, not faster): sum the 32 lanes' 64-bit terms with THREE redux.sync.add.u32 instead of
    four — 27-bit low limbs of both halves, and the two 5-bit tops packed into one word (10-bit fields: 32 * 31 < 1024)."""
    rng = random.Random(5)
    m27 = (1 << 27) - 1
    for _ in range(300):
        terms = [rng.getrandbits(64) if rng.random() < 0.9 else M for _ in range(32)]
        lo = [t & 0xFFFFFFFF for t in terms]
        hi = [t >> 32 for t in terms]
        a = sum(x & m27 for x in lo)                                   # redux 1 (< 2^32: no wrap)
        b = sum((x >> 27) | ((y >> 27) << 10) for x, y in zip(lo, hi))  # redux 2, two 10-bit fields
        c = sum(y & m27 for y in hi)                                   # redux 3
        assert a < 1 << 32 and b < 1 << 32 and c < 1 << 32
        total = (a + ((b & 0x3FF) << 27) + ((c + ((b >> 10) << 27)) << 32)) & M
        assert total == sum(terms) & M


def test_round_without_the_shift():
    """Variant tried in round 1 (correct, not faster): keep C_j = (resolved low bits of z_j) * 0xb3 as an accumulator and add
    zeta * (0xb3 << k) per round, zeta = the newly resolved bit taken from the popcount's bit 0 — the parity then needs no
    shift on the vote -> popc -> vote chain.  Must resolve the same z bytes as the shipped formulation."""
    rng = random.Random(6)
    for _ in range(300):
        m = rng.randrange(0, 97)
        stream = [rng.randrange(256) for _ in range(m)]
        b = [[stream[3 * L + j] if 3 * L + j < m else 0 for j in range(3)] for L in range(32)]
        C = [[0, 0, 0] for _ in range(32)]
        z = [[0, 0, 0] for _ in range(32)]
        for k in range(8):
            t = [[((C[L][j] >> k) ^ (b[L][j] >> k)) & 1 for j in range(3)] for L in range(32)]
            votes = sum((t[L][0] ^ t[L][1] ^ t[L][2]) << L for L in range(32))
            l0k = ((H0 & 0xFF) >> k) & 1
            for L in range(32):
                par = bin(votes & ((1 << L) - 1)).count("1")            # only bit 0 is used
                local = [0, t[L][0], t[L][0] ^ t[L][1]]
                for j in range(3):
                    pre = l0k ^ local[j] ^ ((b[L][j] >> k) & 1)         # prepared while the vote is in flight
                    zeta = (pre ^ par) & 1
                    C[L][j] += zeta * (0xB3 << k)
                    z[L][j] |= zeta << k
        # z must be the xor of the running low byte and the stream byte at every position
        h = H0
        for i, byte in enumerate(stream):
            assert z[i // 3][i % 3] == (h & 0xFF) ^ byte
            h = ((h ^ byte) * P) & M
