"""Issue rates of single instruction kinds and of the thrower's random recipe on the GPU at hand
(wb200_microbench 6..13): the evidence behind DESIGN.md 4.1b (IMAD.WIDE.U32 at a quarter of the
issue rate bounds a Philox4x32-10 call at ~80 cycles per warp).  Usage (GPU box):
    python tools/pipe_rates.py > gpurun_out/pipes.txt"""
import ctypes as C, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.cuda.init(); torch.zeros(1, device='cuda')
from wayne_b200 import _lib
names = {6: 'philox4x32-10 calls', 7: 'philox+BM electrons', 8: 'IMAD.WIDE.U32(+IADD)', 9: 'IMAD', 10: 'LOP3', 11: 'MUFU.LG2', 12: 'FFMA 3-reg', 13: 'I2FP(+LOP)'}
for w in (6, 7, 8, 9, 10, 11, 12, 13):
    ms, ops = C.c_double(), C.c_double()
    _lib.check(_lib.lib.wb200_microbench(w, 2000 if w < 8 else 4000, C.byref(ms), C.byref(ops)), 'mb')
    rate = ops.value / (ms.value * 1e-3)
    # warp-instructions per cycle per SMSP at 1.965 GHz, 592 SMSPs
    print(w, names[w], '%.1f Gops/s' % (rate / 1e9), 'per SMSP per cycle (thread ops / 32): %.3f' % (rate / 32 / 592 / 1.965e9))
