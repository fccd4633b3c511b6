"""CPU: the numpy restatement of the engine's counter-based noise stream against the published Philox4x32-10
known-answer vectors (Random123 kat_vectors) and basic distribution sanity."""
import numpy as np

from oracle import philox_oracle as po


def test_philox4x32_10_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = po.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_normal_stream_properties():
    a = po.normal(3, 1001, seed=1234, sample_index_base=5, step_id=7)
    assert a.shape == (3, 1001) and a.dtype == np.float32 and np.isfinite(a).all()
    # a sample's noise depends on its GLOBAL index only: rows 1.. of base 5 == rows 0.. of base 6
    b = po.normal(2, 1001, seed=1234, sample_index_base=6, step_id=7)
    assert np.array_equal(a[1:], b)
    # different step / seed => different stream
    assert not np.array_equal(a, po.normal(3, 1001, 1234, 5, 8))
    assert not np.array_equal(a, po.normal(3, 1001, 1235, 5, 7))
    big = po.normal(4, 51548, seed=1, sample_index_base=0, step_id=-1)
    assert abs(float(big.mean())) < 0.01 and abs(float(big.std()) - 1.0) < 0.01
