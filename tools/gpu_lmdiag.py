import sys,os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, gen, numpy as np
from rust_compression_b200 import device as dv
data=gen.text(1, 64<<20)
ctx=dv.Context()
d_in=torch.frombuffer(bytearray(data),dtype=torch.uint8).cuda()
out=dv.compress_tensor(ctx,9,d_in)
nb=len(ctx.block_table(with_crc=False)[0])-1
tot=0
for b in range(nb):
    i=ctx.debug_stage(b,"info")
    tot+=i["lm_used"]
print("blocks",nb,"lm total",tot, {k:v for k,v in i.items() if k!='in_use'})
