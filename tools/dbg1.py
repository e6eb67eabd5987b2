import sys
sys.path[:0]=['/root/repo','/root/repo/tests','/root/repo/orb-slam3_byzyh_b200']
import numpy as np, synth, orbfe
from oracle import oracle as O
img=synth.synth_frame(480,752,0)
ex=orbfe.ORBextractor(1000); exc=O.Extractor(1000)
ex(img,None,(0,1000)); exc(img,(0,1000))
for lvl in (0,3):
    L=exc.level(lvl)
    roi=L["padded"][19:-19,19:-19]
    sc=ex.debug_score(lvl)
    c=O.fast(roi[16:-16,16:-16].copy(),7,False)   # corners at minTh, no NMS, coords rel to (16,16)
    ref=np.zeros_like(roi)
    ref[c[:,1]+16,c[:,0]+16]=c[:,2]+1
    dom=np.zeros_like(roi,bool); dom[19:-19,19:-19]=True
    diff=(sc!=ref)&dom
    print("level",lvl,"shape",roi.shape,"mismatch",diff.sum(),"of",dom.sum(),"nonzero ref",(ref>0).sum(),"gpu",((sc>0)&dom).sum())
    ys,xs=np.nonzero(diff)
    for y,x in list(zip(ys,xs))[:12]:
        print(" ",x,y,"gpu",sc[y,x],"ref",ref[y,x])
    if diff.sum():
        print("mismatch x hist (mod 64 of x-19):",np.bincount((xs-19)%64,minlength=64))
        print("mismatch y hist (mod 16 of y-19):",np.bincount((ys-19)%16,minlength=16))
