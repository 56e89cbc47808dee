"""Counter-based stream spec: Random123 known-answer vectors + helper consistency."""
import numpy as np

from oracle.philox import Stream, philox4x32_10, philox_block_numpy, u32_to_range, u32_to_unit


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    f = 0xffffffff
    assert philox4x32_10(f, f, f, f, f, f) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_stream_matches_blocks_and_vectorised():
    s = Stream(seed=(5 << 32) | 77, env_id=9, tag=1)
    words = [s.next_u32() for _ in range(10)]
    blk = philox_block_numpy([0, 1, 2], [9, 9, 9], 1, (5 << 32) | 77).reshape(-1)
    assert words == [int(x) for x in blk[:10]]
    assert s.counter == 10


def test_unit_and_range_maps():
    assert u32_to_unit(0) == 0.0
    assert u32_to_unit(0xffffffff) == 1.0 - 2.0 ** -24
    assert np.float32(u32_to_unit(0x12345678)) == u32_to_unit(0x12345678)  # exact in fp32
    assert u32_to_range(0, 3, 8) == 3 and u32_to_range(0xffffffff, 3, 8) == 7
    s = Stream(1, 2)
    a = s.rand(2, 3)
    assert a.shape == (2, 3) and s.counter == 6 and ((0 <= a) & (a < 1)).all()
    assert 0 <= s.randint(5) < 5 and 2 <= s.randint(2, 4) < 4
    assert s.uniform(0.0, 0.0) == 0.0 and s.counter == 9
