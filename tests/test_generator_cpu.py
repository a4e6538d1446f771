"""CPU checks of the product's counter-based generator (no device needed).

The reference draws its normals with numpy's PCG64 (processes.py:23-24,48); the production kernels replace
it with Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC11).
The known answers below are the Random123 distribution's `kat_vectors` entries for philox4x32-10; the
library exports the very function its kernels inline (`optp_philox4x32_10`, host build of
csrc/optp_common.cuh:philox4x32_10), so a wrong multiplier, key schedule or word order fails here."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def nv():
    import __graft_entry__  # noqa: F401
    from optpricer_b200 import _native
    _native.load()
    return _native


KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),   # digits of pi
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox4x32_10_known_answers(nv, ctr, key, want):
    assert nv.philox4x32_10(ctr, key) == want


def _philox_py(ctr, key, rounds=10):
    """Independent restatement of the published algorithm (Random123 philox.h: M0 = 0xD2511F53,
    M1 = 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85)."""
    c = list(ctr)
    k = list(key)
    for _ in range(rounds):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return tuple(c)


def test_philox_restatement_pins_the_vectors_and_random_counters(nv):
    for ctr, key, want in KAT:
        assert _philox_py(ctr, key) == want
    g = np.random.default_rng(0)
    for _ in range(200):
        ctr = tuple(int(x) for x in g.integers(0, 2**32, 4, dtype=np.uint64))
        key = tuple(int(x) for x in g.integers(0, 2**32, 2, dtype=np.uint64))
        assert nv.philox4x32_10(ctr, key) == _philox_py(ctr, key)


def test_counter_layout_of_the_kernels_is_path_block_stream(nv):
    """csrc/optp_common.cuh:philox_words: counter = (path lo, path hi, block, stream), key = (seed lo, hi).
    Distinct (seed, path, block) give distinct outputs -- the property the multi-GPU sharding relies on
    (a rank's path range never repeats another rank's draws)."""
    seen = set()
    for seed in (1, 2**40 + 7):
        for path in (0, 1, 2**32, 2**32 + 1):
            for block in (0, 1, 62):
                out = nv.philox4x32_10((path & 0xFFFFFFFF, path >> 32, block, 0), (seed & 0xFFFFFFFF, seed >> 32))
                assert out not in seen
                seen.add(out)
    assert len(seen) == 24
