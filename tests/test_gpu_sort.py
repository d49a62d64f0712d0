"""GPU tests of the hand-written radix sort (csrc/sort.cu) through the C-ABI: stable, device-side count, skip word,
implicit values, 16- and 32-bit keys, every pass count.  The checker is torch.sort(stable=True) on the same keys."""
import numpy as np
import pytest
import torch

from dqo_map_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def run_sort(keys, bits, count=None, skip=0, implicit=False, capacity=None, key_dtype=torch.int32):
    L = _lib.lib()
    n_alloc = keys.numel() if capacity is None else capacity
    u16 = key_dtype == torch.int16
    ka = torch.zeros(max(n_alloc, 1), dtype=key_dtype, device=DEV)
    ka[:keys.numel()] = keys.to(key_dtype)
    kb = torch.full_like(ka, -1)
    va = torch.arange(max(n_alloc, 1), dtype=torch.int32, device=DEV) * 3 + 7
    vb = torch.full_like(va, -1)
    vals_in = va.clone()
    tmp = torch.empty(L.dqo_sort_pairs_temp_bytes(n_alloc, bits), dtype=torch.uint8, device=DEV)
    cnt = None if count is None else torch.tensor([count], dtype=torch.int32, device=DEV)
    skp = torch.tensor([skip], dtype=torch.int32, device=DEV)
    fn = L.dqo_sort_pairs_u16 if u16 else L.dqo_sort_pairs_u32
    _lib.check(fn(_lib.ptr(ka), _lib.ptr(kb), _lib.ptr(va), _lib.ptr(vb), int(implicit), _lib.ptr(cnt), _lib.ptr(skp),
                  n_alloc, bits, _lib.ptr(tmp), torch.cuda.current_stream().cuda_stream), "dqo_sort_pairs")
    torch.cuda.synchronize()
    passes = (bits + 7) // 8
    in_a = passes % 2 == 0
    return (ka if in_a else kb), (va if in_a else vb), vals_in


def reference_sort(keys, bits, n, vals):
    k = keys[:n].to(torch.int64) & 0xFFFFFFFF
    digit = k & ((1 << bits) - 1)
    order = torch.sort(digit, stable=True).indices
    return k[order], vals[:n][order]


@pytest.mark.parametrize("n", [1, 31, 256, 4095, 4096, 4097, 50_000, 1_000_003])
@pytest.mark.parametrize("bits", [5, 8, 12, 16, 20, 30, 32])
def test_sort_u32_matches_stable_torch_sort(n, bits):
    g = torch.Generator(device="cpu").manual_seed(n * 131 + bits)
    hi = 1 << min(bits + 2, 31)  # bits above `bits` must be ignored by the sort but carried along
    lo = -(1 << 31) if bits == 32 else 0  # all 32 bits in play, including the top one
    keys = torch.randint(lo, hi, (n,), generator=g, dtype=torch.int64).to(torch.int32).to(DEV)
    ks, vs, vals_in = run_sort(keys, bits)
    rk, rv = reference_sort(keys, bits, n, vals_in)
    assert torch.equal(ks[:n].to(torch.int64) & 0xFFFFFFFF, rk)
    assert torch.equal(vs[:n], rv)


@pytest.mark.parametrize("n", [1, 300, 4096, 70_001, 2_461_184])
@pytest.mark.parametrize("bits", [7, 12, 13, 16])
def test_sort_u16_few_distinct_keys_is_stable(n, bits):
    # tile ids: few distinct keys, long runs of equal keys -- the case where stability decides the blend order
    g = torch.Generator(device="cpu").manual_seed(n + bits)
    keys = torch.randint(0, min(1 << bits, 3225), (n,), generator=g, dtype=torch.int64).to(DEV)
    ks, vs, vals_in = run_sort(keys, bits, key_dtype=torch.int16)
    order = torch.sort(keys, stable=True).indices
    assert torch.equal(ks[:n].to(torch.int64) & 0xFFFF, keys[order])
    assert torch.equal(vs[:n], vals_in[:n][order])


def test_sort_device_count_skip_and_implicit_values():
    n_alloc, n = 200_000, 123_457
    g = torch.Generator(device="cpu").manual_seed(5)
    keys = torch.randint(0, 1 << 30, (n_alloc,), generator=g, dtype=torch.int64).to(torch.int32).to(DEV)
    # only the first `count` items are sorted; the rest of the output buffers is never written
    ks, vs, vals_in = run_sort(keys, 32, count=n, capacity=n_alloc)
    rk, rv = reference_sort(keys, 32, n, vals_in)
    assert torch.equal(ks[:n].to(torch.int64) & 0xFFFFFFFF, rk) and torch.equal(vs[:n], rv)
    # a count beyond the capacity is clamped
    ks, vs, vals_in = run_sort(keys, 32, count=n_alloc + 999, capacity=n_alloc)
    rk, rv = reference_sort(keys, 32, n_alloc, vals_in)
    assert torch.equal(ks.to(torch.int64) & 0xFFFFFFFF, rk) and torch.equal(vs, rv)
    # implicit values = input positions (one pass and four passes)
    for bits in (8, 32):
        ks, vs, _ = run_sort(keys, bits, implicit=True)
        order = torch.sort((keys.to(torch.int64) & 0xFFFFFFFF) & ((1 << bits) - 1), stable=True).indices
        assert torch.equal(vs.to(torch.int64), order)
    # skip word set: nothing moves (outputs keep their fill)
    ks, vs, _ = run_sort(keys, 8, skip=1)
    assert bool((ks == -1).all()) and bool((vs == -1).all())
    # zero count
    ks, vs, _ = run_sort(keys, 8, count=0, capacity=n_alloc)
    assert bool((ks == -1).all())


def test_sort_sorted_and_constant_inputs():
    n = 300_000
    for keys in (torch.arange(n, dtype=torch.int32, device=DEV), torch.full((n,), 0x3F800000, dtype=torch.int32, device=DEV),
                 torch.arange(n, 0, -1, dtype=torch.int32, device=DEV)):
        ks, vs, vals_in = run_sort(keys, 32)
        rk, rv = reference_sort(keys, 32, n, vals_in)
        assert torch.equal(ks.to(torch.int64) & 0xFFFFFFFF, rk) and torch.equal(vs, rv)
