import sys; sys.path.insert(0,'.')
from simplesr_b200 import _lib as L
ctx=L.Context(0)
for n in (32,48,64,96,128,144,192,256):
    print(n, [round(ctx.diag_mma_rate(n, 8192, sh),1) for sh in (0,1,2,3,4,8,18,19)], flush=True)
