"""The oracle's restatement of numpy's Generator(PCG64(SeedSequence(seed))) stream must be
bit-identical to the installed numpy (numpy is available on both boxes, so this runs anywhere)."""
import ctypes

import numpy as np

from oracle import cda_oracle


def test_seedsequence_pcg64_state_matches_numpy():
    L = cda_oracle.lib()
    for seed in [0, 1, 2, 12345, 2**32 - 1, 2**32, 2**40 + 17, 2**63 + 5, 2**64 - 1]:
        out = (ctypes.c_uint64 * 4)()
        L.orc_test_seed(ctypes.c_uint64(seed), out)
        st = np.random.PCG64(np.random.SeedSequence(seed)).state["state"]
        assert (out[0] << 64 | out[1]) == st["state"], seed
        assert (out[2] << 64 | out[3]) == st["inc"], seed


def test_mixed_stream_normal_integers_permutation_bit_exact():
    """standard_normal (ziggurat incl. wedge/tail paths), integers (Lemire, buffered uint32) and
    permutation (masked-rejection Fisher-Yates) interleaved exactly like the env interleaves them."""
    L = cda_oracle.lib()
    rs = np.random.default_rng(5)
    total = 0
    for trial in range(40):
        seed = int(rs.integers(0, 2**62))
        n = 4000
        ops = rs.choice([0, 0, 0, 0, -1, 2, 3, 4, 5, 8, 16, 31], size=n).astype(np.int32)
        outn = np.zeros(n)
        outp = np.zeros(32 * n, dtype=np.int32)
        lo, hi = 10, 101
        L.orc_test_stream(ctypes.c_uint64(seed), n, ops.ctypes.data_as(ctypes.c_void_p),
                          outn.ctypes.data_as(ctypes.c_void_p), outp.ctypes.data_as(ctypes.c_void_p), lo, hi)
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        for i, o in enumerate(ops):
            if o == 0:
                assert g.standard_normal() == outn[i]
            elif o == -1:
                assert g.integers(lo, hi) == outn[i]
            else:
                assert list(g.permutation(int(o))) == list(outp[32 * i:32 * i + o])
            total += 1
    assert total == 160000


def test_normal_loc_scale_is_two_roundings():
    """numpy's normal(loc, scale) is loc + scale*z with separately rounded ops (no FMA) on the
    float32->float64 promoted arguments — the formula the oracle and the kernel use."""
    a = np.random.Generator(np.random.PCG64(np.random.SeedSequence(11)))
    b = np.random.Generator(np.random.PCG64(np.random.SeedSequence(11)))
    src = np.random.default_rng(1)
    for _ in range(5000):
        loc = np.array([src.uniform(-500, 500)], np.float32)
        sc = np.array([src.uniform(0, 1)], np.float32)
        x = a.normal(loc, sc, 1)[0]
        z = b.standard_normal()
        assert x == float(loc[0]) + float(sc[0]) * z
