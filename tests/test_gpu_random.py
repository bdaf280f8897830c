"""-m gpu: randomized parity sweep.  Every case builds two small random files (random shape, ragged headers,
key domain, recid order incl. duplicate recids, string lengths around the 32-byte key-prefix boundary, junk
after the NUL) and requires bit-exact images from all four operators against the CANON oracle."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def random_file(orc, rng, nblocks, mode):
    f = orc.new_blocks(nblocks)
    n = nblocks * 100
    e = f["entries"].reshape(-1).copy()
    dom = int(rng.choice([3, 50, 5000, 2**20, 2**32]))
    e["num"] = rng.integers(0, dom, size=n, dtype=np.uint64).astype(np.uint32)
    if mode["recid"] == "ascending":
        e["recid"] = np.arange(n, dtype=np.uint32) + rng.integers(0, 1000)
    elif mode["recid"] == "shuffled":
        e["recid"] = rng.permutation(n).astype(np.uint32)
    else:  # duplicates: ties on (key, recid) must keep file order
        e["recid"] = rng.integers(0, max(n // 4, 1), size=n).astype(np.uint32)
    maxlen = {"short": 6, "boundary": 34, "long": 119}[mode["strlen"]]
    alphabet = rng.choice([2, 26])
    s = np.zeros((n, 120), dtype=np.uint8)
    lens = rng.integers(0, maxlen + 1, size=n)
    body = rng.integers(97, 97 + alphabet, size=(n, 120)).astype(np.uint8)
    if mode["strlen"] != "short":
        body[:, :28] = ord("p")  # long common prefix: the order is decided near / past the 32-byte prefix
    mask = np.arange(120)[None, :] < lens[:, None]
    s[mask] = body[mask]
    junk = rng.integers(1, 256, size=(n, 120)).astype(np.uint8)
    after = np.arange(120)[None, :] > lens[:, None]
    s[after] = junk[after]  # bytes after the NUL are don't-care for strcmp
    e["str"] = s.view("V120").reshape(-1)
    e["valid"] = 1
    e["dummy1"] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    e["dummy2"] = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    f["entries"] = e.reshape(nblocks, 100)
    f["blockid"] = np.arange(nblocks)
    f["valid"] = 1
    if mode["ragged"]:
        f["nreserved"] = rng.integers(0, 101, size=nblocks).astype(np.uint32)
    else:
        f["nreserved"] = 100
        if nblocks and rng.random() < 0.5:
            f["nreserved"][-1] = rng.integers(1, 101)
    f["dummy"] = f["nreserved"]
    return f


CASES = []
_r = np.random.default_rng(2024)
for i in range(24):
    CASES.append({"seed": 1000 + i, "recid": str(_r.choice(["ascending", "shuffled", "duplicates"])),
                  "strlen": str(_r.choice(["short", "short", "boundary", "long"])), "ragged": bool(_r.random() < 0.4),
                  "nb": int(_r.integers(1, 25))})


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"s{c['seed']}-{c['recid'][:3]}-{c['strlen'][:2]}-{'rag' if c['ragged'] else 'dense'}-{c['nb']}")
def test_random_parity(dbt, orc, case):
    rng = np.random.default_rng(case["seed"])
    r = random_file(orc, rng, case["nb"], case)
    s = random_file(orc, rng, int(rng.integers(1, 25)), case)
    kw = 30 if case["strlen"] != "short" else 8
    for field in "0123":
        got, n = H.dev_sort(dbt, orc, r, field, kw=kw)
        want = orc.sort(r, field)
        assert H.same_image(got, want), ("sort", field, H.first_diff(got, want))
        got, n, u = H.dev_dedup(dbt, orc, r, field, kw=kw)
        want = orc.dedup(r, field)
        assert H.same_image(got, want), ("dedup", field, H.first_diff(got, want))
        want = orc.hashjoin(r, s, field)
        got, n = H.dev_hashjoin(dbt, orc, r, s, field, cap_blocks=max(H.nb(orc.count_rows(want)), 1), kw=kw)
        assert H.same_image(got, want), ("hashjoin", field, H.first_diff(got, want))
        got, ur, us, info = H.dev_mergejoin(dbt, orc, r, s, field, kw=kw)
        want, wur, wus, winfo = orc.mergejoin(r, s, field)
        assert info == winfo, ("mergejoin counters", field, info, winfo)
        assert H.same_image(got, want) and H.same_image(ur, wur) and H.same_image(us, wus), ("mergejoin", field)
