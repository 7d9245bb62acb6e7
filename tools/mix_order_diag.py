# diagnostic: the 50/50 mix with alternating frames (bench.py) against the same frames in two blocks
import sys, os
sys.path.insert(0, "/root/repo")
import torch, lanczos_hls_b200 as lz
F=64
g=torch.Generator(device="cuda"); g.manual_seed(1)
yy=torch.arange(1080,device="cuda",dtype=torch.float32).view(1080,1,1); xx=torch.arange(1920,device="cuda",dtype=torch.float32).view(1,1920,1); cc=torch.arange(3,device="cuda",dtype=torch.float32).view(1,1,3)
def img(f):
    b=128+90*torch.sin(0.05*xx+cc+0.3*f)*torch.cos(0.037*yy); b+=torch.randint(-8,8,(1080,1920,3),device="cuda",generator=g); return b.clamp_(0,255).to(torch.uint8)
def noise(): return torch.randint(0,256,(1080,1920,3),dtype=torch.uint8,device="cuda",generator=g)
d_in=torch.empty((F,1080,1920,3),dtype=torch.uint8,device="cuda"); d_out=torch.empty((F,2160,3840,3),dtype=torch.uint8,device="cuda")
def t():
    for _ in range(3): lz.upscale_batch_device(d_in,d_out,a=3,scale_n=2,scale_d=1)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lz.upscale_batch_device(d_in,d_out,a=3,scale_n=2,scale_d=1)
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/20
for name,order in (("alternating",[f%2 for f in range(F)]),("two blocks",[0]*32+[1]*32),("blocks of 8",[(f//8)%2 for f in range(F)]),("all image",[0]*F),("all noise",[1]*F)):
    for f in range(F): d_in[f]=noise() if order[f] else img(f)
    print("%-12s %.4f ms"%(name,t()))
