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
    """Contract v2: one Philox call = the 4 envs of the aligned group env >> 2 at one step (counter = (group, step)),
    word index env & 3; ja = mulhi(w, 25), step draw r32 = lo32(25 w), reset draw (w >> 2) & 3."""
    rs = np.random.RandomState(0)
    for _ in range(200):
        seed, env, step = (int(rs.randint(0, 2 ** 62)) for _ in range(3))
        grp = env >> 2
        words = philox_py((grp & 0xFFFFFFFF, grp >> 32, step & 0xFFFFFFFF, step >> 32), (seed & 0xFFFFFFFF, seed >> 32))
        w = oracle.philox_word(seed, env, step)
        assert w == words[env & 3]
        ja, r32 = (w * 25) >> 32, (w * 25) & 0xFFFFFFFF
        assert oracle.philox_decode(w) == (ja // 5, ja % 5, r32 >> 30, (w >> 2) & 3)
        assert oracle.philox_r32(w) == r32
        # the table column index the kernels use: jr = mulhi(w, 100) = ja * 4 + (r32 >> 30)
        assert (w * 100) >> 32 == ja * 4 + (r32 >> 30)
    # the initial reset uses step 2^64 - 1
    assert oracle.philox_word(5, 9, (1 << 64) - 1) == philox_py((2, 0, 0xFFFFFFFF, 0xFFFFFFFF), (5, 0))[1]


def test_decode_is_uniform_over_joint_action_and_draw(oracle):
    # every (joint action, 2-bit step draw) cell owns 2^32/100 +- 1 of the 2^32 words, and within a cell
    # the reset draw ((w >> 2) & 3) is uniform to within four words
    edges = [-(-(c << 32) // 100) for c in range(101)]          # first w with mulhi(w, 100) == c
    sizes = np.diff(np.array(edges, dtype=np.int64))
    assert sizes.min() >= (1 << 32) // 100 and sizes.max() <= (1 << 32) // 100 + 1
    for c in (0, 37, 99):
        assert (int(edges[c]) * 100) >> 32 == c and ((int(edges[c]) - 1) * 100) >> 32 == c - 1
    assert all(abs((e1 - e0) // 16 - ((e1 - e0 + 15) // 16)) <= 1 for e0, e1 in zip(edges, edges[1:]))
    # r32 = lo32(25 w): 25 is odd, so w -> r32 is a bijection of the 32-bit words (exactly uniform draw); inside one
    # joint action the draws are the arithmetic progression r0 + 25 j, i.e. uniform at a resolution of 25 / 2^32
    assert pow(25, -1, 1 << 32) * 25 % (1 << 32) == 1
    w0 = int(edges[4 * 7])                                      # first word of joint action 7
    assert [((w0 + j) * 25) & 0xFFFFFFFF for j in range(3)] == [(w0 * 25 + 25 * j) & 0xFFFFFFFF for j in range(3)]
    assert (w0 * 25) & 0xFFFFFFFF < 25
