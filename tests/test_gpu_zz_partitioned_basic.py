"""GPU parity of the partitioned basic stencils (`stencils b200 basic partitioned-*`).

The host logic (scatter with the neighbours' rows as j halo, one sweep per device, gather) is
covered on emulated devices in tests/test_host_datapath.py; here the same classes run on real
devices: on one GPU (`gpus=1`, any box) and on every pair / all GPUs of a multi-GPU box.
(The file sorts after the other GPU tests on purpose: it was added last; first GPU run:
profiles/gputests_partitioned_basic_r02.log.)
"""

import numpy as np
import pytest

from oracle import stencils
from stencil_benchmarks_b200 import capi
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import basic

pytestmark = pytest.mark.gpu

CASES = {
    "copy": (basic.PartitionedCopy, {}, lambda f, h: stencils.copy(f, h)),
    "onesided_j": (basic.PartitionedOnesidedAverage, dict(axis=1), lambda f, h: stencils.onesided_average(f, h, 1)),
    "symmetric_j": (basic.PartitionedSymmetricAverage, dict(axis=1), lambda f, h: stencils.symmetric_average(f, h, 1)),
    "symmetric_k": (basic.PartitionedSymmetricAverage, dict(axis=2), lambda f, h: stencils.symmetric_average(f, h, 2)),
    "laplacian_ij": (basic.PartitionedLaplacian, {}, lambda f, h: stencils.laplacian(f, h, (True, True, False))),
    "laplacian_ijk": (basic.PartitionedLaplacian, dict(along_z=True),
                      lambda f, h: stencils.laplacian(f, h, (True, True, True))),
}


def gpu_counts():
    count = capi.device_count()
    return sorted({1, min(2, count), count} - {0})


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_partitioned_basic_on_the_gpus_of_the_box(name, dtype):
    cls, parameters, oracle = CASES[name]
    for gpus in gpu_counts():
        bench = cls(domain=(301, 37, 7), halo=(3, 3, 3), dtype=dtype, gpus=gpus, verify=False, seed=23,
                    **parameters)
        data = bench.data()
        inp0, out0 = np.array(data.inp, copy=True), np.array(data.out, copy=True)
        result = bench.run()
        assert result["gpus"] == gpus and result["time"] > 0 and result["bandwidth"] > 0
        expected = oracle(inp0, bench.halo)
        inner = bench.inner_slice()
        assert np.allclose(data.out[inner], expected[inner], **stencils.tolerances(dtype)), (name, dtype, gpus)
        assert np.array_equal(data.inp, inp0)
        outside = np.ones(out0.shape, dtype=bool)
        outside[inner] = False
        assert np.array_equal(data.out[outside], out0[outside]), "only interior points of out may change"


def test_more_gpus_than_present_is_an_execution_error():
    from stencil_benchmarks_b200.benchmark import ExecutionError

    bench = basic.PartitionedCopy(domain=(64, 64, 4), gpus=capi.device_count() + 1, verify=False)
    with pytest.raises(ExecutionError, match="GPUs requested"):
        bench.run()
