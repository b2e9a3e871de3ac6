"""Developer script (not a test): accuracy + CUDA-event timing of every conv implementation."""
import sys; sys.path.insert(0,'.')
import numpy as np, torch
from oracle import ref_numpy as RN
from pbml_mantle_convection_b200 import ops, _lib as L
DEV='cuda:0'
cu=lambda a: torch.tensor(np.ascontiguousarray(a),dtype=torch.float32,device=DEV)
r=np.random.default_rng(0)
IMPLS=["ffma","umma_f16x2","umma_bf16"]
for (B,Ci,Co,H,W,k,pad) in [(1,16,16,40,70,3,'replicate'),(1,8,16,20,20,3,'zeros'),(1,16,16,21,66,5,'zeros')]:
    x=r.standard_normal((B,Ci,H,W)); w=r.standard_normal((Co,Ci,k,k))/np.sqrt(Ci*k*k); b=r.standard_normal(Co)
    ref=RN.conv2d_same(x,w,b,pad)
    for impl in IMPLS:
        out,_,_=ops.conv_fwd([ops.Source(ops.pack_nchw(cu(x)))],ops.pack_conv_weight(cu(w),[Ci]),ops.pad_vec(cu(b),Co,DEV),Co,k,pad,impl=impl,wpk_umma=ops.pack_conv_weight_umma(cu(w),[Ci]))
        torch.cuda.synchronize()
        y=ops.unpack_nchw(out,Co).cpu().numpy()
        print((B,Ci,Co,H,W,k,pad),impl,'rel',np.linalg.norm(y-ref)/np.linalg.norm(ref))
def timeit(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize(); return a.elapsed_time(b)/n*1e3
g=torch.Generator(device=DEV).manual_seed(1)
for (B,H,W) in [(1,512,512),(32,256,256),(1,128,128)]:
    x16=torch.randn(B,4,H,W,4,device=DEV,generator=g)
    stats=torch.stack([x16.double().sum((2,3,4)),(x16.double()**2).sum((2,3,4))],-1).contiguous()
    gam=torch.ones(16,device=DEV); bet=torch.zeros(16,device=DEV)
    w=torch.randn(16,16,3,3,device=DEV)/12; wpk=ops.pack_conv_weight(w,[16]); wum=ops.pack_conv_weight_umma(w,[16]); bias=torch.zeros(16,device=DEV)
    src=ops.Source(x16,L.XFORM_GN_GELU,stats,gam,bet)
    o=torch.empty_like(x16); st=torch.zeros_like(stats)
    w1=torch.randn(16,103,3,3,device=DEV)/30; wpk1=ops.pack_conv_weight(w1,[16]*6+[7]); wum1=ops.pack_conv_weight_umma(w1,[16]*6+[7])
    srcs=[src]+[ops.Source(torch.randn(B,4,H,W,4,device=DEV,generator=g)) for _ in range(5)]+[ops.Source(torch.randn(B,2,H,W,4,device=DEV,generator=g))]
    for impl in IMPLS:
        t=timeit(lambda: ops.conv_fwd([src],wpk,bias,16,3,'replicate',impl=impl,wpk_umma=wum,out=o,stats=st))
        t1=timeit(lambda: ops.conv_fwd(srcs,wpk1,bias,16,3,'replicate',impl=impl,wpk_umma=wum1,out=o,stats=st))
        cells=B*H*W
        print(f"B{B} {H}x{W} {impl:12s} conv16x16 {t:8.1f} us ({cells*128/t/1e3:7.1f} GB/s, {cells*4608/t/1e6:6.1f} TF)   conv1 {t1:8.1f} us ({cells*29664/t1/1e6:6.1f} TF)")
