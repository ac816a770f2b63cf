"""Run the decode-design microbenchmarks on a B200 (build first: `make -C tools/microbench`):
grid-barrier latency per protocol, same-address atomic reduction cost, and HBM→smem streaming rate vs CTA count."""
import ctypes as C
import os

import torch

here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "libmb.so"))
P = C.c_void_p
lib.mb_grid_barrier.argtypes = [C.c_int, C.c_int, C.c_int, P, P, C.POINTER(C.c_float)]
lib.mb_atomic_contention.argtypes = [C.c_int, C.c_int, C.c_int, P, P, C.POINTER(C.c_float)]
lib.mb_stream.argtypes = [C.c_int, C.c_int, P, P, C.POINTER(C.c_float)]
dev = torch.device("cuda")
sms = torch.cuda.get_device_properties(0).multi_processor_count
mhz = 1e3 * float(os.environ.get("SM_GHZ", "1.9"))
cycles = torch.zeros(1024, dtype=torch.int64, device=dev)
scratch = torch.zeros(4, dtype=torch.int32, device=dev)
ms = C.c_float()

print(f"== device-wide barrier, {sms} co-resident CTAs (1 per SM)")
for variant, name in ((0, "monotonic counter, all poll it"), (1, "counter + epoch flag (last arriver publishes)"),
                      (2, "cooperative_groups grid.sync()")):
    for threads in (256, 512):
        for iters in (10, 1000):                      # 10 = warm-up
            rc = lib.mb_grid_barrier(variant, iters, threads, scratch.data_ptr(), cycles.data_ptr(), C.byref(ms))
            assert rc > 0, rc
        c = cycles[:sms].float()
        print(f"  {name:48s} {threads:4d} thr: {ms.value * 1e3 / iters:6.2f} us per barrier "
              f"(clock64: mean {float(c.mean()) / iters:7.0f}, max {float(c.max()) / iters:7.0f} cycles)")

print("== same-address reduction: every CTA adds a [rows x 1024] fp32 tile into one global tile (red.global.add.v4.f32)")
dst = torch.zeros(64 * 1024, dtype=torch.float32, device=dev)
for rows in (1, 16, 64):
    for ctas in (19, 37, 74, 148):
        for rounds in (1, 20):
            assert lib.mb_atomic_contention(ctas, rows, rounds, dst.data_ptr(), cycles.data_ptr(), C.byref(ms)) == 0
        c = cycles[:ctas].float()
        print(f"  rows {rows:3d} ctas {ctas:4d}: {ms.value * 1e3 / rounds:7.2f} us per round (clock64 max {float(c.max()) / rounds:8.0f} cycles)")

print("== streaming HBM → shared memory with cp.async.bulk (16 KB chunks, 8 stages), distinct 16 MB region per CTA")
chunks = 1024
src = torch.empty(sms * chunks * 16384, dtype=torch.uint8, device=dev)
src.zero_()
for ctas in (8, 16, 32, 64, 96, 128, sms):
    for rep in range(2):
        assert lib.mb_stream(ctas, chunks, src.data_ptr(), cycles.data_ptr(), C.byref(ms)) == 0
    gb = ctas * chunks * 16384 / 1e9
    print(f"  {ctas:4d} CTAs: {gb / (ms.value / 1e3):8.1f} GB/s aggregate, {gb / (ms.value / 1e3) / ctas:6.1f} GB/s per CTA")
