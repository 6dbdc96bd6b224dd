"""div_known (csrc/sfxm_device.cuh) must equal the IEEE quotient bit for bit: it replaces every division of the
reference's stencil arithmetic, so a single wrong rounding would break bit-exact parity."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,n", [(0, 1 << 28), (1, 1 << 28), (2, 1 << 28)])
def test_division_by_known_reciprocal_is_correctly_rounded(mode, n):
    import torch
    from splitfxm_b200 import _lib
    lib = _lib.load()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    for seed in (1, 0xDEADBEEF):
        _lib.check(lib.sfxm_selftest_div(n, seed, mode, bad.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert int(bad[0]) == 0
