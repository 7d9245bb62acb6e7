"""Debug driver: one launch of a given config, checked against the generic kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lanczos_hls_b200 as lz
f, ih, iw, c, n, d, a = (int(v) for v in sys.argv[1:8])
oh, ow = ih * n // d, iw * n // d
torch.manual_seed(1)
d_in = torch.randint(0, 256, (f, ih, iw, c), dtype=torch.uint8, device="cuda")
d_out = torch.zeros((f, oh, ow, c), dtype=torch.uint8, device="cuda")
d_ref = torch.zeros((f, oh, ow, c), dtype=torch.uint8, device="cuda")
lz.upscale_batch_device(d_in, d_ref, a=a, scale_n=n, scale_d=d, flags=lz.FLAG_GENERIC_KERNEL)
torch.cuda.synchronize()
lz.upscale_batch_device(d_in, d_out, a=a, scale_n=n, scale_d=d)
torch.cuda.synchronize()
st = lz.stats()
diff = (d_out != d_ref)
print("ok", sys.argv[1:8], st, "mismatches", int(diff.sum()))
if diff.any():
    idx = diff.nonzero()[:5].tolist()
    print(" first mismatches (f,y,x,c):", idx)
