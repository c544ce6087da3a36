import sys, os
sys.path.insert(0,"cuda-akaze_b200"); sys.path.insert(0,"tests"); sys.path.insert(0,".")
import torch, numpy as np, akaze_b200 as ab, bench as BN
F=256
frames8=BN.make_frames(F,"shapes")
dev=(torch.from_numpy(frames8).float()*np.float32(1/255.)).cuda() if False else torch.from_numpy((frames8.astype(np.float32)*np.float32(1.0/255.0))).cuda()
def run(nctx, chunk):
    ctxs=[ab.Context(BN.W,BN.H,max_batch=chunk,max_pts=10000) for _ in range(nctx)]
    per=F//nctx
    res=[c.alloc_results(per,True) for c in ctxs]
    streams=[c.torch_stream() for c in ctxs]
    def step():
        for i,c in enumerate(ctxs):
            c.detect_and_compute(dev[i*per:(i+1)*per], True, out=res[i])
    for _ in range(2): step()
    torch.cuda.synchronize()
    cur=torch.cuda.current_stream()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for s in streams: s.wait_event(e0)
    for _ in range(3): step()
    for s in streams:
        ev=torch.cuda.Event(); ev.record(s); cur.wait_event(ev)
    e1.record(cur); e1.synchronize()
    ms=e0.elapsed_time(e1)/3
    for c in ctxs: c.close()
    return F/(ms*1e-3)
for nctx,chunk in ((1,32),(2,16),(2,32),(4,16)):
    print(nctx,chunk, round(run(nctx,chunk),1), flush=True)
