"""GPU: radix sort / segment primitives through the C ABI, bit-exact against numpy's stable argsort."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,hi", [(1, 5), (31, 3), (2048, 6041), (2049, 6041), (5000, 300), (100_000, 1 << 20),
                                  (1_000_003, 10_000_000), (70_000, 2)])
def test_sort_pairs_is_numpy_stable_argsort(n, hi):
    from recbole_fairrec_b200 import kernels
    rng = np.random.default_rng(n)
    keys = rng.integers(0, hi, size=n).astype(np.int32)
    k = torch.from_numpy(keys).cuda()
    bits = max(1, int(hi - 1).bit_length())
    ko, vo = kernels.sort_pairs(k, None, bits)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(vo.cpu().numpy(), order.astype(np.int32))
    np.testing.assert_array_equal(ko.cpu().numpy(), keys[order])


def test_sort_pairs_with_values():
    from recbole_fairrec_b200 import kernels
    rng = np.random.default_rng(7)
    keys = rng.integers(0, 1000, size=33_333).astype(np.int32)
    vals = rng.integers(0, 1 << 30, size=33_333).astype(np.int32)
    ko, vo = kernels.sort_pairs(torch.from_numpy(keys).cuda(), torch.from_numpy(vals).cuda(), 10)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(vo.cpu().numpy(), vals[order])
