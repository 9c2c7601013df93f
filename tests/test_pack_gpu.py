"""GPU parity: tpack / tunpack (CUDA, through the C-ABI and the quant_engine op) vs the oracle and the golden KATs."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_pack_kat
from gpu_util import c_tpack, c_tunpack

pytestmark = pytest.mark.gpu
TDT = {"float32": torch.float32, "float64": torch.float64, "float16": torch.float16, "int8": torch.int8,
       "int32": torch.int32}


@pytest.mark.parametrize("case", load_pack_kat(), ids=lambda c: f"b{c['n_bits']}{'s' if c['sign'] else 'u'}-{c['dtype']}-{len(c['values'])}")
def test_kat_through_engine(engine, case):
    x = torch.tensor(case["values"]).reshape(case["shape"]).to(TDT[case["dtype"]]).cuda()
    packed, des = engine.tpack(x, case["n_bits"], case["sign"])
    assert packed.dtype == torch.uint8 and packed.is_cuda and des.dtype == torch.int32 and des.is_cuda
    assert bytes(packed.cpu().numpy().tolist()).hex() == case["packed_hex"]
    assert des.cpu().tolist() == case["des"]
    back = engine.tunpack(packed, des)
    assert str(back.dtype).replace("torch.", "") == case["unpacked_dtype"]
    assert list(back.shape) == case["shape"]
    assert torch.equal(back.cpu().to(torch.int64), torch.tensor(case["values"]).reshape(case["shape"]).to(torch.int64))


SIZES = [1, 7, 8, 9, 255, 256, 257, 1023, 4096 + 5, 8 * 1024 * 4 + 3, (1 << 20) + 13]


@pytest.mark.parametrize("n_bits", range(1, 9))
@pytest.mark.parametrize("sign", [False, True])
def test_bytes_equal_oracle(n_bits, sign):
    rng = np.random.default_rng(n_bits * 2 + sign)
    lo, hi = (-(1 << (n_bits - 1)), (1 << (n_bits - 1)) - 1) if sign else (0, (1 << n_bits) - 1)
    for n in SIZES:
        v = rng.integers(lo, hi + 1, size=n)
        want, _ = oracle.tpack(v, n_bits, sign)
        for dt in (torch.float32, torch.int8 if hi < 128 else torch.int16, torch.float16, torch.int64, torch.float64,
                   torch.bfloat16, torch.int32, torch.uint8 if lo >= 0 else torch.int16):
            x = torch.tensor(v).to(dt).cuda()
            got, flag = c_tpack(x, n_bits, sign)
            assert flag == 0
            assert np.array_equal(got.cpu().numpy(), want), (n, dt)
        back = c_tunpack(torch.from_numpy(want).cuda(), n, n_bits, sign)
        assert np.array_equal(back.cpu().numpy().astype(np.int64), v)


def test_unaligned_and_empty():
    rng = np.random.default_rng(5)
    v = rng.integers(-8, 8, size=5000)
    base = torch.tensor(v).float().cuda()
    for off in (1, 3, 5):
        x = base[off:]
        want, _ = oracle.tpack(v[off:], 4, True)
        got, flag = c_tpack(x, 4, True)
        assert flag == 0 and np.array_equal(got.cpu().numpy(), want)
        pk = torch.zeros(want.size + 3, dtype=torch.uint8, device="cuda")
        pk[3:] = torch.from_numpy(want).cuda()
        back = c_tunpack(pk[3:], v.size - off, 4, True)
        assert np.array_equal(back.cpu().numpy(), v[off:])
    got, flag = c_tpack(torch.zeros(0, device="cuda"), 4, True)
    assert got.numel() == 0 and flag == 0


def test_range_check(engine):
    for vals, n_bits, sign in (([0, 8], 4, True), ([-9, 0], 4, True), ([-1], 4, False), ([16], 4, False),
                               ([float("nan")], 8, True), ([0] * 5000 + [300], 8, False)):
        with pytest.raises(RuntimeError, match="out of range"):
            engine.tpack(torch.tensor(vals, dtype=torch.float32).cuda(), n_bits, sign)
    with pytest.raises(RuntimeError, match=r"\(0, 8\]"):
        engine.tpack(torch.zeros(4).cuda(), 9, True)
    with pytest.raises(RuntimeError, match="too short"):
        engine.tunpack(torch.zeros(4, dtype=torch.uint8).cuda(), torch.tensor([4, 1], dtype=torch.int32).cuda())
    with pytest.raises(RuntimeError, match="uint8"):
        engine.tunpack(torch.zeros(4).cuda(), torch.tensor([4, 1, 8], dtype=torch.int32).cuda())


def test_cpu_tensors_are_staged_through_the_gpu(engine):
    # reference accepts CPU tensors (tpack.cu:241-251); checkpoint loading relies on it (quantconv2d.py:218-235)
    v = torch.randint(-8, 8, (3, 5, 7)).float()
    packed, des = engine.tpack(v, 4, True)
    assert not packed.is_cuda and not des.is_cuda
    back = engine.tunpack(packed, des)
    assert not back.is_cuda and torch.equal(back.float(), v)


def test_round_trip_at_microbench_size():
    """BASELINE config 5 size (2^27 elements): size-independent property instead of the oracle."""
    n = 1 << 27
    g = torch.Generator(device="cuda").manual_seed(0)
    for n_bits in (4, 3):
        x = torch.randint(-(1 << (n_bits - 1)), 1 << (n_bits - 1), (n,), generator=g, device="cuda", dtype=torch.int8)
        packed, flag = c_tpack(x, n_bits, True)
        assert flag == 0
        back = c_tunpack(packed, n, n_bits, True)
        assert torch.equal(back, x)
        # first 2^20 elements against the oracle, and a popcount checksum of the whole stream
        want, _ = oracle.tpack(x[: 1 << 20].cpu().numpy(), n_bits, True)
        assert np.array_equal(packed[: want.size].cpu().numpy(), want)
        del packed, back, x


def test_reference_extension_agrees():
    """The unmodified reference kernels (oracle/_ref, compiled from /root/reference) on the same GPU."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built")
    ref = build_ref.load()
    rng = np.random.default_rng(11)
    for n_bits, sign, n in ((4, True, 36864), (3, False, 10001), (8, True, 4099), (6, True, 777)):
        lo, hi = (-(1 << (n_bits - 1)), (1 << (n_bits - 1)) - 1) if sign else (0, (1 << n_bits) - 1)
        x = torch.tensor(rng.integers(lo, hi + 1, size=n)).float().cuda()
        rp, rd = ref.tpack(x, n_bits, sign)
        got, flag = c_tpack(x, n_bits, sign)
        assert flag == 0 and torch.equal(got, rp)
        assert torch.equal(c_tunpack(rp, n, n_bits, sign), ref.tunpack(rp, rd).reshape(-1))
