"""CPU: the index arithmetic of the shared-memory small-channel conv kernels (csrc/gg_conv_small.cu) — tile origins, the
zero-filled dy / input patches and their extents, the stride-parity thread classes of the dgrad kernel, ragged tiles —
emulated thread by thread in Python and compared with torch autograd.  The CUDA kernels follow this arithmetic line by line;
their numerics are checked on the GPU by tests/test_gpu_kernels.py."""
import pytest
import numpy as np
import torch
F=torch.nn.functional
def floor_div(a,b): return a//b
def emu_dgrad(dy,w,B,H,W,Ci,Co,k,s,pt,pl,Ho,Wo):
    kDH,kDW=8,16
    PH=(kDH-1+k-1)//s+2; PW=(kDW-1+k-1)//s+2
    dx=np.full((B,H,W,Ci),np.nan)
    tiles_w=-(-W//kDW); tiles_h=-(-H//kDH)
    for b in range(B):
      for blk in range(tiles_w*tiles_h):
        h0=(blk//tiles_w)*kDH; w0=(blk%tiles_w)*kDW
        hy_min=(h0+pt-(k-1))//s; wx_min=(w0+pl-(k-1))//s
        patch=np.zeros((PH,PW,Co))
        for ph in range(PH):
          for pw in range(PW):
            hy,wx=hy_min+ph,wx_min+pw
            if 0<=hy<Ho and 0<=wx<Wo: patch[ph,pw]=dy[b,hy,wx]
        ncls=s*s; per=128//ncls; cols=kDW//s
        for tid in range(128):
            cls=tid//per; idx=tid%per; ch=cls//s; cw=cls%s
            h=h0+(idx//cols)*s+ch; wv=w0+(idx%cols)*s+cw
            acc=np.zeros(Ci)
            r0=((h+pt)%s+s)%s; s0=((wv+pl)%s+s)%s
            for r in range(r0,k,s):
                t=h+pt-r; hy=int(t/s)-hy_min   # C truncation
                assert t%s==0 and 0<=hy<PH,(h,r,hy,PH)
                for ss in range(s0,k,s):
                    t2=wv+pl-ss; wx=int(t2/s)-wx_min
                    assert t2%s==0 and 0<=wx<PW,(wv,ss,wx,PW)
                    acc+= w[r,ss]@patch[hy,wx]     # [Ci,Co]@[Co]
            if h<H and wv<W:
                assert np.isnan(dx[b,h,wv,0])
                dx[b,h,wv]=acc
    assert not np.isnan(dx).any()
    return dx
def emu_fwd(x,w,B,H,W,Ci,Co,k,s,pt,pl,Ho,Wo):
    T=8; IW=(T-1)*s+k
    y=np.full((B,Ho,Wo,Co),np.nan)
    tiles_w=-(-Wo//T); tiles_h=-(-Ho//T)
    for b in range(B):
      for blk in range(tiles_w*tiles_h):
        ho0=(blk//tiles_w)*T; wo0=(blk%tiles_w)*T
        hi0=ho0*s-pt; wi0=wo0*s-pl
        sx=np.zeros((IW,IW,Ci))
        for ih in range(IW):
          for iw in range(IW):
            hi,wi=hi0+ih,wi0+iw
            if 0<=hi<H and 0<=wi<W: sx[ih,iw]=x[b,hi,wi]
        sxf=sx.reshape(-1)
        for pl_ in range(16):
          for j in range(4):
            pix=pl_+16*j; py,px=pix//T,pix%T
            xo=((py*s)*IW+px*s)*Ci
            acc=np.zeros(Co)
            for r in range(k):
              for ss in range(k):
                xoff=(r*IW+ss)*Ci
                for c in range(Ci):
                    acc+=sxf[xo+xoff+c]*w[r,ss,c]
            ho,wo=ho0+py,wo0+px
            if ho<Ho and wo<Wo: y[b,ho,wo]=acc
    assert not np.isnan(y).any()
    return y
def wgrad_plan(B,H,W,Ci,Co,k,s,Ho,Wo,resident=15):
    """mirror of wgrad_plan() + the resident-cluster cap of conv_small_wgrad() in csrc/gg_conv_small.cu"""
    TK,CL,MAXC=5,8,18
    K=k*k*Ci
    assert 1<=Ci<=4 and Co%8==0 and K<=100
    rh_max=16384//(Wo*Co); assert rh_max>=1
    RH=min(Ho,rh_max)
    PW=(Wo-1)*s+k
    NT=(-(-K//TK))*(Co//8); assert NT<=512
    PS=min(512//NT,4)
    threads=max(128,-(-NT*PS//32)*32)
    clusters=min(-(-(B*Ho)//(2*CL)),MAXC,resident)
    return dict(RH=RH,PW=PW,NT=NT,PS=PS,threads=threads,grid=clusters*CL,clusters=clusters)
def emu_wgrad(x,dy,B,H,W,Ci,Co,k,s,pt,pl,Ho,Wo):
    """thread-by-thread emulation of conv_small_wgrad_kernel: equal contiguous row ranges per CTA walked in units of <= RH rows
    of one image, x patch / dy rows in 'shared memory', the (5 filter rows x 8 channels) register tile, pixel groups, the CTA
    fold, the cluster fold by float4 slices and the last-cluster fold"""
    P=wgrad_plan(B,H,W,Ci,Co,k,s,Ho,Wo)
    RH,PW,NT,PS,grid,ncl=P['RH'],P['PW'],P['NT'],P['PS'],P['grid'],P['clusters']
    K=k*k*Ci; N=K*Co; nq8=Co//8
    tiles=np.zeros((grid,N))
    total=B*Ho; seen=np.zeros(total,int)
    for cta in range(grid):
        accs=np.zeros((NT*PS,5,8))
        g_lo,g_hi=total*cta//grid,total*(cta+1)//grid
        g0=g_lo
        while g0<g_hi:
            b,ho0=g0//Ho,g0%Ho
            rows=min(Ho-ho0,RH,g_hi-g0)
            seen[g0:g0+rows]+=1
            g0+=rows
            npix=rows*Wo; PP=-(-npix//PS)
            hi0=ho0*s-pt; wi0=-pl
            nx=((rows-1)*s+k)*PW*Ci
            sx=np.zeros(nx)
            for i in range(nx):
                c=i%Ci; iw=(i//Ci)%PW; ih=i//(Ci*PW); hi,wi=hi0+ih,wi0+iw
                if 0<=hi<H and 0<=wi<W: sx[i]=x[b,hi,wi,c]
            sdy=dy[b].reshape(-1)[ho0*Wo*Co:(ho0+rows)*Wo*Co]
            for tid in range(NT*PS):
                ps,tt=tid//NT,tid%NT; g,q=tt//nq8,tt%nq8
                off=[]
                for j in range(5):
                    kk=g*5+j; tap,c=kk//Ci,kk%Ci
                    off.append(((tap//k)*PW+(tap%k))*Ci+c if kk<K else 0)
                p_lo=min(npix,ps*PP); p_hi=min(npix,p_lo+PP)
                pr,wo=p_lo//Wo,p_lo%Wo
                for pix in range(p_lo,p_hi):
                    xo=((pr*s)*PW+wo*s)*Ci
                    d0=sdy[pix*Co+q*4:pix*Co+q*4+4]; d1=sdy[pix*Co+(q+nq8)*4:pix*Co+(q+nq8)*4+4]
                    for j in range(5):
                        xv=sx[xo+off[j]]
                        accs[tid,j,:4]+=xv*d0; accs[tid,j,4:]+=xv*d1
                    wo+=1
                    if wo==Wo: wo=0; pr+=1
        for sgrp in range(PS):
            for tid in range(NT*PS):
                ps,tt=tid//NT,tid%NT; g,q=tt//nq8,tt%nq8
                if ps!=sgrp: continue
                for j in range(5):
                    kk=g*5+j
                    if kk<K:
                        tiles[cta,kk*Co+q*4:kk*Co+q*4+4]+=accs[tid,j,:4]
                        tiles[cta,kk*Co+(q+nq8)*4:kk*Co+(q+nq8)*4+4]+=accs[tid,j,4:]
    assert (seen==1).all()
    n4=N//4; per=-(-n4//8)
    part=np.full((ncl,N),np.nan)
    for cid in range(ncl):
        for rank in range(8):
            lo4,hi4=rank*per,min(n4,rank*per+per)
            for i in range(lo4,hi4):
                part[cid,4*i:4*i+4]=sum(tiles[cid*8+r,4*i:4*i+4] for r in range(8))
    assert not np.isnan(part).any()
    return part.sum(0).reshape(k,k,Ci,Co)
def case(B,H,W,Ci,Co,k,s,padding):
    if padding=='SAME':
        Ho=-(-H//s); Wo=-(-W//s); ph=max((Ho-1)*s+k-H,0); pw=max((Wo-1)*s+k-W,0); pt,pl=ph//2,pw//2
    else:
        Ho=(H-k)//s+1; Wo=(W-k)//s+1; ph=pw=0; pt=pl=0
    rs=np.random.RandomState(0)
    x=torch.tensor(rs.randn(B,Ci,H,W),requires_grad=True); w=torch.tensor(rs.randn(k,k,Ci,Co),requires_grad=True)
    xp=F.pad(x,(pl,pw-pl,pt,ph-pt))
    y=F.conv2d(xp,w.permute(3,2,0,1),stride=s)
    gy=torch.tensor(rs.randn(*y.shape))
    dx,=torch.autograd.grad(y,x,gy,retain_graph=True)
    got=emu_dgrad(gy.permute(0,2,3,1).numpy(),w.detach().numpy(),B,H,W,Ci,Co,k,s,pt,pl,Ho,Wo)
    e1=np.abs(got-dx.permute(0,2,3,1).numpy()).max()
    got2=emu_fwd(x.detach().permute(0,2,3,1).numpy(),w.detach().numpy(),B,H,W,Ci,Co,k,s,pt,pl,Ho,Wo)
    e2=np.abs(got2-y.detach().permute(0,2,3,1).numpy()).max()
    print((B,H,W,Ci,Co,k,s,padding),"dgrad err %.2e fwd err %.2e"%(e1,e2))
    assert e1<1e-9 and e2<1e-9
    if Co%8==0:
        dw,=torch.autograd.grad(y,w,gy)
        got3=emu_wgrad(x.detach().permute(0,2,3,1).numpy(),gy.permute(0,2,3,1).numpy(),B,H,W,Ci,Co,k,s,pt,pl,Ho,Wo)
        e3=np.abs(got3-dw.numpy()).max()
        print("wgrad err %.2e"%e3)
        assert e3<1e-9


@pytest.mark.parametrize("geom", [(1, 32, 32, 3, 8, 5, 2, 'SAME'), (1, 28, 28, 1, 4, 5, 2, 'SAME'), (1, 14, 14, 2, 4, 5, 2, 'SAME'),
                                  (1, 7, 7, 3, 4, 5, 2, 'SAME'), (1, 16, 20, 3, 4, 3, 1, 'SAME'), (1, 9, 11, 4, 4, 4, 1, 'VALID'),
                                  (1, 5, 6, 2, 4, 5, 2, 'SAME'), (3, 12, 10, 3, 16, 5, 2, 'SAME'), (2, 9, 11, 1, 8, 4, 1, 'VALID'),
                                  (150, 3, 4, 4, 8, 3, 1, 'SAME')])
def test_small_channel_kernel_index_math(geom):
    case(*geom)
