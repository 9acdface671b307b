"""Philox4x32-10: the oracle's C implementation against the Random123 known-answer vectors
(Random123 `kat_vectors`, Salmon et al. SC'11) and against an independent pure-Python
restatement; plus the word/decode contract of DESIGN.md section 4.  CPU only."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox_py(ctr, key):
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return tuple(c)


def test_random123_known_answers(oracle):
    for ctr, key, want in KAT:
        assert philox_py(ctr, key) == want
        assert tuple(int(x) for x in oracle.philox4x32_10(ctr, key)) == want


def test_word_and_decode_contract(oracle):
    rs = np.random.RandomState(0)
    for _ in range(200):
        seed, env, step = (int(rs.randint(0, 2 ** 62)) for _ in range(3))
        blk = step >> 2
        words = philox_py((env & 0xFFFFFFFF, env >> 32, blk & 0xFFFFFFFF, blk >> 32), (seed & 0xFFFFFFFF, seed >> 32))
        w = oracle.philox_word(seed, env, step)
        assert w == words[step & 3]
        jr = (w * 100) >> 32
        assert oracle.philox_decode(w) == ((jr >> 2) // 5, (jr >> 2) % 5, jr & 3, w & 3)


def test_decode_is_uniform_over_joint_action_and_draw(oracle):
    # every (joint action, step draw) cell owns 2^32/100 +- 1 of the 2^32 words, and within a cell
    # the reset draw (w & 3) is uniform to within one word
    edges = [-(-(c << 32) // 100) for c in range(101)]          # first w with mulhi(w, 100) == c
    sizes = np.diff(np.array(edges, dtype=np.int64))
    assert sizes.min() >= (1 << 32) // 100 and sizes.max() <= (1 << 32) // 100 + 1
    for c in (0, 37, 99):
        assert (int(edges[c]) * 100) >> 32 == c and ((int(edges[c]) - 1) * 100) >> 32 == c - 1
    assert all(abs((e1 - e0) // 4 - ((e1 - e0 + 3) // 4)) <= 1 for e0, e1 in zip(edges, edges[1:]))
