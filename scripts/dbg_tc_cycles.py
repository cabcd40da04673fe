import ctypes, torch, sys
sys.path.insert(0,'/root/repo')
from starst3r_b200 import match, _lib
lib=_lib.load(); dev=torch.device('cuda:0')
g=torch.Generator().manual_seed(0)
for M in (4096, 32768):
    A=torch.nn.functional.normalize(torch.randn(M,24,generator=g),dim=-1).to(dev)
    B=torch.nn.functional.normalize(torch.randn(262144,24,generator=g),dim=-1).to(dev)
    match.nn_argmax(A,B,impl='tcgen05'); torch.cuda.synchronize()
    lib.st3r_debug_nn_tc_cycles(None,1)
    match.nn_argmax(A,B,impl='tcgen05'); torch.cuda.synchronize()
    out=(ctypes.c_ulonglong*4)(); lib.st3r_debug_nn_tc_cycles(out,0)
    t,w,e,tot=[int(x) for x in out]
    print(f"M={M}: tiles {t}  wait/tile {w/t:.0f}  epilogue/tile {e/t:.0f}  loop/tile {tot/t:.0f} cycles")
