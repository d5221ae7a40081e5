import sys, time
sys.path.insert(0, '.')
import torch, ofdg_b200 as o
B=64
g=o.Generator(device=0, mode=7, max_batch=B)
g.synth_textures(200, 1024, 768, seed=0)
img0=torch.empty((B,3,384,512),device='cuda'); img1=torch.empty_like(img0); fl=torch.empty((B,2,384,512),device='cuda')
st=torch.cuda.Stream(); torch.cuda.set_stream(st)
for fg in (1, 4, 8, 0, 40):
    ps=o.ParamStream(7, fg_override=fg)
    prep=[g.prepare(ps.generate(B)) for _ in range(3)]
    for i in range(20): g.render_prepared(prep[i%3], img0,img1,fl, st.cuda_stream)
    torch.cuda.synchronize(); g.kernel_times()
    for i in range(200): g.render_prepared(prep[i%3], img0,img1,fl, st.cuda_stream)
    torch.cuda.synchronize()
    p,r,n=g.kernel_times()
    print('fg_override',fg,'render ms',r/n,'prep ms',p/n)
