"""The Philox4x32-10 restatement behind the noise-augmentation checker, against the published known-answer vectors
(Random123 kat_vectors: `philox4x32 10 <ctr> <key> <expected>`), and the moments of the normal stream."""
import numpy as np

from oracle import philox

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(np.asarray([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(v) for v in got) == want, (ctr, key, [hex(int(v)) for v in got])


def test_normal_stream_is_a_function_of_seed_and_offset():
    a = philox.normals(4000, seed=7)
    b = philox.normals(4000, seed=7)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, philox.normals(4000, seed=8))
    # advancing the offset by the groups consumed continues the stream
    c = philox.normals(2000, seed=7, offset=500)
    assert np.array_equal(a[2000:], c)
    z = philox.normals(400000, seed=1)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs((z ** 3).mean()) < 2e-2 and abs((z ** 4).mean() - 3.0) < 5e-2
