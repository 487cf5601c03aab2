"""The bit-exact build's branch-free division / square root (om_div_rn with its shared reciprocal refinement, om_sqrt_rn)
return the IEEE-754 correctly rounded result — the bits of div.rn.f64 / sqrt.rn.f64 — on 10^8 random operand pairs each,
on the values Hydro actually divides (zero numerators, equal operands, ...), and raise their range flag outside 2^+-400."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_check_lib():
    import subprocess
    from paraiso_b200.build import NVCC_ARCH, nvcc_path
    d = os.path.join(ROOT, "tests", "cuda", "_build")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, "libdivsqrt_check.so")
    src = os.path.join(ROOT, "tests", "cuda", "divsqrt_check.cu")
    hdr = os.path.join(ROOT, "paraiso_b200", "csrc", "om_runtime.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run([nvcc_path()] + NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", "-fmad=false", "-shared", "-Xcompiler", "-fPIC",
                       "-I", os.path.dirname(hdr), src, "-o", so], check=True)
    return so


def _random_doubles(torch, n, gen, emin=-400, emax=400, signed=True):
    """n doubles with uniformly random mantissa bits and exponents in [emin, emax]."""
    mant = torch.randint(0, 1 << 52, (n,), generator=gen, device="cuda", dtype=torch.int64)
    expo = torch.randint(1023 + emin, 1023 + emax + 1, (n,), generator=gen, device="cuda", dtype=torch.int64)
    bits = mant | (expo << 52)
    if signed:
        bits = bits | (torch.randint(0, 2, (n,), generator=gen, device="cuda", dtype=torch.int64) << 63)
    return bits.view(torch.float64)


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(build_check_lib())


def _div(lib, torch, a, b):
    ours, ieee = torch.empty_like(a), torch.empty_like(a)
    bad = torch.zeros(2, dtype=torch.int32, device="cuda")
    rc = lib.om_check_div(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(ours.data_ptr()),
                          ctypes.c_void_p(ieee.data_ptr()), ctypes.c_void_p(bad.data_ptr()), ctypes.c_longlong(a.numel()), None)
    assert rc == 0
    torch.cuda.synchronize()
    return ours.view(torch.int64), ieee.view(torch.int64), (int(bad[0].item()), int(bad[1].item()))


def _sqrt(lib, torch, x):
    ours, ieee = torch.empty_like(x), torch.empty_like(x)
    bad = torch.zeros(2, dtype=torch.int32, device="cuda")
    rc = lib.om_check_sqrt(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(ours.data_ptr()), ctypes.c_void_p(ieee.data_ptr()),
                           ctypes.c_void_p(bad.data_ptr()), ctypes.c_longlong(x.numel()), None)
    assert rc == 0
    torch.cuda.synchronize()
    return ours.view(torch.int64), ieee.view(torch.int64), (int(bad[0].item()), int(bad[1].item()))


def test_division_is_correctly_rounded_on_1e8_random_pairs(lib):
    import torch
    gen = torch.Generator(device="cuda").manual_seed(20261017)
    n, total = 1 << 24, 0
    for it in range(6):
        # wide exponents (numerators down to the guard's 2^-900, denominators across its 2^+-100); then operands of similar
        # magnitude (quotients near 1, where a wrong last bit would show most often)
        sa, sb = (890, 100) if it < 3 else (2, 2)
        a, b = _random_doubles(torch, n, gen, -sa, sa), _random_doubles(torch, n, gen, -sb, sb)
        ours, ieee, (bad, nslow) = _div(lib, torch, a, b)
        assert bad == 0 and nslow == 0          # inside the guards: never sent to the slow path ...
        assert bool((ours == ieee).all()), f"{int((ours != ieee).sum())} of {n} quotients differ from div.rn.f64"      # ... and IEEE's bits
        total += n
    assert total >= 10 ** 8


def test_division_special_operands(lib):
    import torch
    gen = torch.Generator(device="cuda").manual_seed(7)
    n = 1 << 20
    b = _random_doubles(torch, n, gen, -100, 100)
    zeros = torch.zeros(n, dtype=torch.float64, device="cuda")
    for a in (zeros, -zeros, b.clone(), -b, torch.ones_like(b), b * 1.5, torch.full_like(b, 5.0 / 3.0)):
        ours, ieee, (bad, nslow) = _div(lib, torch, a, b)
        assert bad == 0 and nslow == 0 and bool((ours == ieee).all())      # bit patterns: the sign of a zero quotient included
    # mantissas of all ones / all zeros in numerator and denominator
    edge = torch.tensor([0x3FF0000000000000, 0x3FFFFFFFFFFFFFFF, 0x3FF0000000000001, 0x4000000000000000, 0x3FEFFFFFFFFFFFFF,
                         0x3FE0000000000001, 0x3FF8000000000000, 0x3FF7FFFFFFFFFFFF], dtype=torch.int64, device="cuda").view(torch.float64)
    a, b = torch.meshgrid(edge, edge, indexing="ij")
    ours, ieee, (bad, nslow) = _div(lib, torch, a.reshape(-1).contiguous(), b.reshape(-1).contiguous())
    assert bad == 0 and nslow == 0 and bool((ours == ieee).all())


def test_sqrt_is_correctly_rounded_on_1e8_random_operands(lib):
    import torch
    gen = torch.Generator(device="cuda").manual_seed(99)
    n, total = 1 << 24, 0
    for it in range(6):
        x = _random_doubles(torch, n, gen, -400 if it < 4 else -1, 400 if it < 4 else 1, signed=False)
        ours, ieee, (bad, nslow) = _sqrt(lib, torch, x)
        assert bad == 0 and nslow == 0
        assert bool((ours == ieee).all()), f"{int((ours != ieee).sum())} of {n} roots differ from sqrt.rn.f64"
        total += n
    assert total >= 10 ** 8
    # perfect squares, zero (both signs), values next to a square
    k = torch.arange(1, 1 << 20, device="cuda", dtype=torch.float64)
    for x in (k * k, torch.zeros(8, dtype=torch.float64, device="cuda"), -torch.zeros(8, dtype=torch.float64, device="cuda"),
              torch.nextafter(k * k, torch.full_like(k, 1e300)), torch.nextafter(k * k, torch.zeros_like(k))):
        ours, ieee, (bad, nslow) = _sqrt(lib, torch, x.contiguous())
        assert bad == 0 and nslow == 0 and bool((ours == ieee).all())


def test_tiny_operands_raise_the_slow_flag_and_specials_the_state_guard(lib):
    """Tiny non-zero operands (the denormal velocities at the front of a spreading perturbation) must send the cell to the
    IEEE slow path; NaN / Inf results are caught by the state guard when stored."""
    import torch
    one = torch.ones(4, dtype=torch.float64, device="cuda")
    for a, b in ((one * 1e-300, one), (one * 5e-324, one), (one, one * 1e-40), (one, one * 1e40), (one, one * 5e-324)):
        assert _div(lib, torch, a.contiguous(), b.contiguous())[2][1] == 4
    for a, b in ((one * 0.0, one), (one * 1e-270, one * 1e30), (one * 1e250, one * 1e-30)):
        assert _div(lib, torch, a.contiguous(), b.contiguous())[2] == (0, 0)
    for a, b in ((one * float("inf"), one), (one, one * float("nan")), (one, one * 0.0)):
        assert _div(lib, torch, a.contiguous(), b.contiguous())[2][0] == 1
    assert _sqrt(lib, torch, (one * 1e-300).contiguous())[2][1] == 4
    assert _sqrt(lib, torch, (one * 5e-324).contiguous())[2][1] == 4
    for x in (one * -1.0, one * float("inf"), one * float("nan")):
        assert _sqrt(lib, torch, x.contiguous())[2][0] == 1
    assert _sqrt(lib, torch, (one * 1e-250).contiguous())[2] == (0, 0)
    # wide random operands beyond the guards: wherever the flag stays down the bits are IEEE's
    gen = torch.Generator(device="cuda").manual_seed(5)
    a, b = _random_doubles(torch, 1 << 22, gen, -1000, 1000), _random_doubles(torch, 1 << 22, gen, -300, 300)
    bad = torch.zeros(2, dtype=torch.int32, device="cuda")
    ours, ieee, (_bad, nslow) = _div(lib, torch, a, b)
    assert 0 < nslow < a.numel()


def test_fast_math_division_and_sqrt_stay_within_their_error_bound(lib):
    """The fast_math build's sequences (om_frcp + om_fdiv_r, om_fsqrt) are not IEEE: the reciprocal is within 1 ulp, a quotient
    within 2 ulp, a square root within 1 ulp of the correctly rounded value on 2^26 random operands; sqrt(0) is exactly 0."""
    import torch
    gen = torch.Generator(device="cuda").manual_seed(31)
    n = 1 << 24
    worst_q = worst_r = 0
    for it in range(4):
        span = 250 if it < 2 else 2
        a = _random_doubles(torch, n, gen, -span, span, signed=False)
        b = _random_doubles(torch, n, gen, -span, span)
        quot, root = torch.empty_like(a), torch.empty_like(a)
        rc = lib.om_check_fast(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(quot.data_ptr()),
                               ctypes.c_void_p(root.data_ptr()), ctypes.c_longlong(n), None)
        assert rc == 0
        torch.cuda.synchronize()
        # same sign and both normal: the distance in ulps is the difference of the bit patterns
        worst_q = max(worst_q, int((quot.view(torch.int64) - (a / b).view(torch.int64)).abs().max().item()))
        worst_r = max(worst_r, int((root.view(torch.int64) - torch.sqrt(a).view(torch.int64)).abs().max().item()))
    assert worst_q <= 2 and worst_r <= 1, (worst_q, worst_r)
    # zero radicands / numerators (Hydro: velocity1 == 0), perfect squares, a radicand below the 1e-284 the seed's offset is exact for
    k = torch.arange(0, 1 << 16, device="cuda", dtype=torch.float64)
    x = torch.cat([k * k, torch.zeros(8, dtype=torch.float64, device="cuda"), torch.full((8,), 1e-290, dtype=torch.float64, device="cuda")]).contiguous()
    b = torch.full_like(x, 3.0)
    quot, root = torch.empty_like(x), torch.empty_like(x)
    assert lib.om_check_fast(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(quot.data_ptr()),
                             ctypes.c_void_p(root.data_ptr()), ctypes.c_longlong(x.numel()), None) == 0
    torch.cuda.synchronize()
    assert bool((root[: 1 << 16] == k).all()) and bool((root[1 << 16: (1 << 16) + 8] == 0).all()) and bool((quot[(1 << 16): (1 << 16) + 8] == 0).all())
    assert float((root[-8:] / 1e-145 - 1).abs().max()) < 1e-9


def test_exact_build_raises_when_the_state_leaves_the_normal_range():
    """Machine level: a NaN that reaches the stored state surfaces as a RuntimeError at the next host read; denormal
    velocities do not — those cells take the IEEE slow path and the state stays bit-identical to the oracle's."""
    from paraiso_b200.machines import hydro_machine, hydro_set_params
    size = (64, 48)
    m = hydro_machine(size)
    hydro_set_params(m, size)
    m.call("init")
    m.call("proceed")
    m.scalar("time")                                    # in range: no error
    p = m.get("velocity0")
    p[10, 10] = float("nan")
    m.set("velocity0", p)
    m.call("proceed")
    with pytest.raises(RuntimeError, match="exact_divsqrt"):
        m.scalar("time")
