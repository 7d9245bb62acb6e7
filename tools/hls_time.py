import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, lanczos_hls_b200 as lz
F=16
d_in=torch.randint(0,256,(F,1080,1920,3),dtype=torch.uint8,device='cuda'); d_out=torch.empty((F,2160,3840,3),dtype=torch.uint8,device='cuda')
for _ in range(3): lz.upscale_hls_device(d_in,d_out)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lz.upscale_hls_device(d_in,d_out)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print("HLS mode 1080p->2160p x16: %.3f ms  %.1f Gpix/s  %.1f GB/s"%(ms, F*3840*2160/ms/1e6, F*31104000/ms/1e6))
