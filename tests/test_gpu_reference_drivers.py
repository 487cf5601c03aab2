"""The reference's own example drivers, compiled unchanged against the generated B200 host class
(tests/refdrivers.py; linked by __graft_entry__.build() where /root/reference is mounted), run on the GPU: their
stdout must equal, byte for byte, what the same drivers print on the reference-style C++ class (tests/golden/driver_*.txt)."""
import os

import pytest

from tests import refdrivers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key", refdrivers.KEYS)
def test_reference_driver_prints_the_reference_output(key):
    exe = refdrivers.exe_path(key)
    if not os.path.exists(exe):
        if not os.path.isdir(refdrivers.REF):
            pytest.skip("driver not prebuilt and /root/reference not mounted")
        exe = refdrivers.link_b200(key)
    with open(refdrivers.golden_path(key)) as f:
        want = f.read()
    got = refdrivers.run(key, exe)
    assert got == want
