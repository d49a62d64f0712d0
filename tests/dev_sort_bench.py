"""Developer benchmark of csrc/sort.cu alone (not collected by pytest): the two sorts of the c2 forward."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dqo_map_b200 import _lib

DEV = torch.device("cuda:0")
L = _lib.lib()


def bench(n, bits, u16, reps=20, label=""):
    dt = torch.int16 if u16 else torch.int32
    g = torch.Generator(device="cpu").manual_seed(1)
    if u16:
        keys = torch.randint(0, 3225, (n,), generator=g, dtype=torch.int64).to(dt).to(DEV)
    else:
        keys = (torch.rand(n, generator=g) * 4.5 + 0.5).view(torch.int32).to(DEV)
    ka, kb = keys.clone(), torch.empty_like(keys)
    va, vb = torch.arange(n, dtype=torch.int32, device=DEV), torch.empty(n, dtype=torch.int32, device=DEV)
    tmp = torch.empty(L.dqo_sort_pairs_temp_bytes(n, bits), dtype=torch.uint8, device=DEV)
    cnt = torch.tensor([n], dtype=torch.int32, device=DEV)
    fn = L.dqo_sort_pairs_u16 if u16 else L.dqo_sort_pairs_u32
    st = torch.cuda.current_stream().cuda_stream
    def run():
        _lib.check(fn(_lib.ptr(ka), _lib.ptr(kb), _lib.ptr(va), _lib.ptr(vb), 0, _lib.ptr(cnt), None, n, bits, _lib.ptr(tmp), st), "sort")
    for _ in range(3):
        ka.copy_(keys); run()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    # device time: everything queued back to back (the host runs ahead), minus the same number of restore copies alone
    e0.record()
    for _ in range(reps):
        ka.copy_(keys); run()
    e1.record()
    for _ in range(reps):
        ka.copy_(keys)
    e2.record()
    torch.cuda.synchronize()
    us = 1000 * (e0.elapsed_time(e1) - e1.elapsed_time(e2)) / reps
    print("%s n=%d bits=%d: %.1f us per sort (%d passes)" % (label, n, bits, us, (bits + 7) // 8), flush=True)


if __name__ == "__main__":
    bench(2_461_184, 12, True, label="tile sort c2 front")
    bench(24_614, 12, True, label="tile sort c2 back ")
    bench(1_000_000, 32, False, label="depth sort c2     ")
    bench(15_054_845, 12, True, label="tile sort c2 single-phase")
