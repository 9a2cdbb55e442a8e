"""The synthetic-input helpers (bench.py / the GPU tests build their trees from them): the torch generators, run on the CPU
device here, produce the same stream and the same well-formed indexed tree as the numpy ones, and that tree satisfies the
reference's linked-list invariants (indexed_merkle_tree.rs:632-660: sorted successor links, head at slot 0, largest -> (0, 0))."""
import numpy as np
import pytest

import imt_b200
from imt_b200 import synth
import oracle as O


@pytest.mark.parametrize("n,m", [(8, 1), (8, 2), (64, 20), (256, 256)])
def test_numpy_and_torch_generators_agree_and_are_well_formed(n, m):
    a = synth.indexed_preimages(n, m, seed=3)
    b = synth.indexed_preimages_torch(n, m, seed=3, device="cpu").numpy().view(np.uint64)
    assert np.array_equal(a, b)
    vals = {O.to_int(a[i, 0]): i for i in range(1, m)}
    assert len(vals) == m - 1 and all(0 < v < imt_b200.P for v in vals)
    cur, seen = 0, 0
    order = sorted(vals)
    for v in order:                                                     # follow the list from the head: ascending, every slot once
        assert O.to_int(a[cur, 1]) == v and int(a[cur, 2, 0]) == vals[v]
        cur = vals[v]
        seen += 1
    assert O.to_int(a[cur, 1]) == 0 and O.to_int(a[cur, 2]) == 0 and seen == m - 1
    assert not a[m:].any()                                              # empty slots are {0, 0, 0}


def test_field_element_streams_agree():
    a = synth.field_elements(1000, seed=11)
    b = synth.field_elements_torch(1000, seed=11, device="cpu").numpy().view(np.uint64)
    assert np.array_equal(a, b)
    assert np.array_equal(synth.field_elements(10, seed=11, first=990), a[990:])
