"""Survivor rates of FAST early-reject tests on the bench frames (numpy + cv2 resize; DESIGN.md section 4, profiles/r2_fast_survivors.txt)."""
import sys, numpy as np, cv2
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tests')
from synth import synth_frame
RING=[(0,3),(1,3),(2,2),(3,1),(3,0),(3,-1),(2,-2),(1,-3),(0,-3),(-1,-3),(-2,-2),(-3,-1),(-3,0),(-3,1),(-2,2),(-1,3)]
def stats(img, ts=(20,7)):
    h,w=img.shape
    I=img.astype(np.int32)
    c=I[3:h-3,3:w-3]
    r=np.stack([I[3+dy:h-3+dy,3+dx:w-3+dx] for dx,dy in RING])
    out={}
    # best score
    d=r-c[None]
    ext=np.concatenate([d,d[:8]],0)
    mn9=np.stack([ext[k:k+9].min(0) for k in range(16)]); mx9=np.stack([ext[k:k+9].max(0) for k in range(16)])
    best=np.maximum(mn9.max(0), (-mx9).max(0)); best=np.maximum(best,0)
    for t in ts:
        b=r> c[None]+t; dk=r< c[None]-t
        compB=(b[0]|b[8])&(b[4]|b[12]); compD=(dk[0]|dk[8])&(dk[4]|dk[12])
        comp=compB|compD
        p8B=np.all(b[:8]|b[8:],0); p8D=np.all(dk[:8]|dk[8:],0)
        p8=p8B|p8D
        ns=np.abs(d)>t
        compN=(ns[0]|ns[8])&(ns[4]|ns[12])
        p8N=np.all(ns[:8]|ns[8:],0)
        # diag pairs too: 2,10 and 6,14
        comp8B=compB&(b[2]|b[10])&(b[6]|b[14]); comp8D=compD&(dk[2]|dk[10])&(dk[6]|dk[14])
        out[t]=dict(compass=comp.mean(), compass_absdiff=compN.mean(), four_pairs=(comp8B|comp8D).mean(), pairs8=p8.mean(), pairs8_abs=p8N.mean(), corner=(best>t).mean())
    return out
for (H,W) in [(480,752),(720,1280)]:
    img=synth_frame(H,W,7)
    acc={}
    tot=0
    sf=1.0
    for l in range(8):
        w=int(round(W/sf)); h=int(round(H/sf))
        lv=img if l==0 else cv2.resize(img,(w,h),interpolation=cv2.INTER_LINEAR)
        s=stats(lv)
        n=lv.size; tot+=n
        for t in s:
            for k,v in s[t].items(): acc[(t,k)]=acc.get((t,k),0)+v*n
        if l in (0,3,7): print(H,W,'level',l,{t:{k:round(v,4) for k,v in s[t].items()} for t in s})
        sf*=1.2
    print(H,W,'ALL',{k:round(v/tot,4) for k,v in acc.items()})
