"""vg_attn_decode alone: achieved HBM GB/s per launch at decode batch B, cache length Tk (bf16, H=16, D=64).
usage: python tools/attn_decode_bench.py [Tk] [B ...]   (VG_ATTN_DECODE_STREAM=0 selects the register-load kernel)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_gslm_b200 import _lib, ops

Tk = int(sys.argv[1]) if len(sys.argv) > 1 else 402
Bs = [int(x) for x in sys.argv[2:]] or [1, 8, 64, 128, 256]
H, D, L = 16, 64, 16
dev = torch.device("cuda", 0)
_lib.load()
slopes = torch.tensor(ops.alibi_slopes(H), device=dev)
for B in Bs:
    Tmax = Tk + 8
    # one cache per layer, as in the step: consecutive launches read different memory (nothing is L2-resident)
    kc = (0.5 * torch.randn(L, B, H, Tmax, D, device=dev)).bfloat16()
    vc = (0.5 * torch.randn(L, B, H, Tmax, D, device=dev)).bfloat16()
    qkv = (0.5 * torch.randn(B, 3 * H * D, device=dev)).bfloat16()
    out = torch.empty(B, H * D, device=dev, dtype=torch.bfloat16)
    tickets = torch.zeros(B * H, dtype=torch.int32, device=dev)
    pos_dev = torch.tensor([Tk - 1], dtype=torch.int32, device=dev)
    for splits in (None, 1, 2, 4):
        def run():
            for i in range(L):
                ops.attention_decode(qkv, kc[i], vc[i], 0, slopes, pos_dev=pos_dev, out=out, tickets=tickets, splits=splits)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()          # the launch rate of the python wrapper must not be what is measured
        with torch.cuda.graph(graph):
            run()
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * L)
        nbytes = B * H * Tk * D * 2 * 2
        print(f"B={B:4d} Tk={Tk} splits={splits}: {us:7.1f} us / launch   {nbytes / us / 1e3:7.1f} GB/s", flush=True)
