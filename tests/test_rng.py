"""CPU tier: the CPU twin of the device RNG (oracle/rng.py) against published
Philox known answers and numpy's legacy stream."""
import numpy as np

from oracle import rng


def test_philox4x32_10_random123_kat():
    M = 0xFFFFFFFF
    assert [int(x) for x in rng.philox4x32(0, 0, 0, 0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(x) for x in rng.philox4x32(M, M, M, M, (M << 32) | M)] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    got = rng.philox4x32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, (0x299F31D0 << 32) | 0xA4093822)
    assert [int(x) for x in got] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_mulhi64_exact():
    a = np.array([2 ** 64 - 1, 12345678901234567, 0, 1 << 63], dtype=np.uint64)
    for b in (1, 1000003, 2 ** 40 + 7):
        assert [int(x) for x in rng.mulhi64(a, b)] == [(int(x) * b) >> 64 for x in a]


def test_feistel_is_a_bijection():
    for m in (1, 2, 3, 17, 1000, 4096, 100003):
        p = rng.feistel_permutation(m, 2, 27)
        assert np.array_equal(np.sort(p), np.arange(m))
    assert not np.array_equal(rng.feistel_permutation(1000, 0, 27), rng.feistel_permutation(1000, 1, 27))


def test_null_cols_in_range_and_uniform():
    c = rng.philox_null_cols(1000, 200000, 0, 27)
    assert c.min() >= 0 and c.max() < 1000
    assert abs(np.bincount(c, minlength=1000).std() / 200 - 1 / np.sqrt(200)) < 0.02


def test_pvalue_draws_sharding_invariant():
    full = rng.philox_pvalue_draws(np.arange(50), 101, 30, 27)
    part = rng.philox_pvalue_draws(np.arange(20, 30), 101, 30, 27)
    assert np.array_equal(full[20:30], part) and full.max() < 30 and full.min() >= 0


def test_numpy_replay_known_answers():
    # seed(27); choice(1902772, 3, replace=False)
    assert list(rng.numpy_repeat_seeds(27, 3)) == [1076532, 128070, 219018]
    cols, perm, rs = rng.numpy_null_graph(50, 3, 5)
    np.random.seed(5)
    c2 = np.random.choice(np.arange(50), 150, replace=True)
    x = np.arange(150) * 1.5
    np.random.shuffle(x)
    assert np.array_equal(cols, c2) and np.array_equal(x, (np.arange(150) * 1.5)[perm])
    assert np.array_equal(rs.randint(0, 30, size=(4, 7)), np.random.randint(0, 30, size=(4, 7)))


def test_pvalue_draw_digits_are_uniform_and_independent():
    """Three draws per Philox word (base-nb digits of w / 2^32): each digit uniform on [0, nb), consecutive digits of one
    word independent (chi-square on singles and on pairs), and the single-draw scheme is kept for nb > 64."""
    nb = 30
    d = rng.philox_pvalue_draws(np.arange(4000), 1200, nb, 27)
    assert d.min() == 0 and d.max() == nb - 1
    for j in range(3):                                       # digit j of every word
        c = np.bincount(d[:, j::3].ravel(), minlength=nb)
        chi2 = ((c - c.mean()) ** 2 / c.mean()).sum()
        assert chi2 < 2.5 * nb, (j, chi2)                    # dof 29: far inside the 1e-6 tail bound
    first, second = d[:, 0::3], d[:, 1::3]
    m = min(first.shape[1], second.shape[1])
    c = np.bincount((first[:, :m] * nb + second[:, :m]).ravel(), minlength=nb * nb)
    chi2 = ((c - c.mean()) ** 2 / c.mean()).sum()
    assert chi2 < 1.25 * nb * nb, chi2                       # dof 899
    wide = rng.philox_pvalue_draws(np.arange(50), 101, 200, 27)
    assert wide.shape == (50, 101) and wide.max() < 200
    # prefix property: fewer permutations are a prefix of more (sharding by permutation ranges relies on it)
    assert np.array_equal(rng.philox_pvalue_draws(np.arange(7), 100, nb, 5), rng.philox_pvalue_draws(np.arange(7), 250, nb, 5)[:, :100])


def test_label_permutation_is_a_keyed_bijection():
    """The label-permutation null (extension) relabels cells with a Feistel bijection keyed by (permutation, seed):
    a permutation of [0, n), different per permutation id and seed, and disjoint from the K4b weight-shuffle streams."""
    from oracle import restate as R
    from oracle import rng as orng
    n = 10_007
    base = R.LABEL_PERM_BASE
    p0 = orng.feistel_permutation(n, base + 0, 27)
    p1 = orng.feistel_permutation(n, base + 1, 27)
    q0 = orng.feistel_permutation(n, base + 0, 28)
    k0 = orng.feistel_permutation(n, 0, 27)                     # what K4b repeat 0 would use on the same domain
    for p in (p0, p1, q0):
        assert np.array_equal(np.sort(p), np.arange(n))
    assert (p0 != p1).mean() > 0.99 and (p0 != q0).mean() > 0.99 and (p0 != k0).mean() > 0.99
    fixed = (p0 == np.arange(n)).sum()
    assert fixed < 10                                           # ~1 fixed point expected for a random permutation
