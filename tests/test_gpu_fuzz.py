"""Random shapes through both GF(256) elimination paths (tools/fuzz_tc_vs_table.py): the tensor-core
path must leave the same ranks, pivot columns and bytes as the shared-memory table kernel, which
test_gpu_solve.py pins against the oracle and the compiled reference."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_tensor_core_path_equals_table_kernel_on_random_shapes(seed, cuda_lib):
    import fuzz_tc_vs_table as fz
    from oblas_b200 import device

    device.bind(0)
    rng = random.Random(seed)
    for i in range(10):
        ok, desc = fz.one_case(cuda_lib, rng, seed * 1000 + i)
        assert ok, desc
