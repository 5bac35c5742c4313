"""Prototype (CPU, NumPy): evaluate the reference's wide Gaussians at reduced
resolution and measure the parity impact on the uint8 mosaic."""
import sys, numpy as np, cv2
sys.path.insert(0, '/root/repo')
from oracle import restate as rs, cv_semantics as cvs
from pano360_b200 import synth

def reflect101(p, n):
    return cvs.reflect101(p, n)

def decimate(img, f, R):
    """area-average f x f cells of the reflect-101 extension of img over full-res range [-R, n+R)."""
    h, w = img.shape[:2]
    ys = reflect101(np.arange(-R, h + R + (-(h + 2*R)) % f), h)
    xs = reflect101(np.arange(-R, w + R + (-(w + 2*R)) % f), w)
    ext = img[ys][:, xs]
    H, W = ext.shape[:2]
    return ext.reshape(H//f, f, W//f, f, -1).mean(axis=(1, 3), dtype=np.float32)

def lowres_blur(lr, sigma_lr):
    ks = max(3, int(np.rint(sigma_lr*8+1))|1)
    k = cvs.gaussian_kernel(sigma_lr, ks)
    return cv2.sepFilter2D(lr, -1, k, k, borderType=cv2.BORDER_REPLICATE)

def expand(lr, f, R, h, w):
    u = ((np.arange(w) + R + 0.5)/f - 0.5).astype(np.float32)
    v = ((np.arange(h) + R + 0.5)/f - 0.5).astype(np.float32)
    mx, my = np.meshgrid(u, v)
    return cv2.remap(lr, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REPLICATE)

def approx_blur(img, sigma, f):
    h, w = img.shape[:2]
    var = sigma**2 - (f*f-1)/12.0 - f*f/6.0
    s_lr = np.sqrt(var)/f
    R = f*int(np.ceil((4*sigma + f)/f)) 
    lr = decimate(img, f, R)
    lr = lowres_blur(lr, s_lr)
    return expand(lr, f, R, h, w)

FACTORS = None
def multiband_approx(patches, shape, n_levels, factors):
    own = rs.owner_map(patches, shape, 'stream')
    for i,(w,_,where) in enumerate(patches): w[...,3] = own[where]==i
    covered = np.zeros(shape,bool); out=np.zeros(shape+(3,),np.float32); prev=[None]*len(patches)
    for lvl in range(n_levels):
        sigma = rs.band_sigma(lvl)
        bs = np.zeros(shape+(3,),np.float32); ws=np.zeros(shape,np.float32); last = lvl==n_levels-1
        for i,(warped,inv,where) in enumerate(patches):
            tile = prev[i] if prev[i] is not None else warped.copy()
            if not last:
                f = factors[lvl]
                blur = cv2.GaussianBlur(warped,(0,0),sigma) if f==1 else approx_blur(warped, sigma, f)
                tile[...,:3] -= blur[...,:3]; tile[...,3]=blur[...,3]; prev[i]=blur
            bs[where] += tile[...,:3]*tile[...,3:4]; ws[where]+=tile[...,3]
            if lvl==0: covered[where] |= ~inv
        bs[~covered,:]=0; ws[ws==0]=1; out += bs/ws[...,None]
    return (255*np.clip(out,0,1)).astype(np.uint8)

def psnr(a,b):
    e=np.mean((a.astype(float)-b.astype(float))**2); return 99 if e==0 else 10*np.log10(255**2/e)

if __name__ == '__main__':
    rng = np.random.default_rng(0)
    img = rng.random((200,300,4),dtype=np.float32)
    img[60:, 100:, 3] = 1; img[:60,:,3]=0; img[:, :100, 3]=0
    for lvl in range(5):
        s = rs.band_sigma(lvl); ref = cv2.GaussianBlur(img,(0,0),s)
        for f in (2,4,8):
            if s*s - (f*f-1)/12 - f*f/6 <= 0.5: continue
            a = approx_blur(img, s, f)
            print('lvl',lvl,'f',f,'max err rgb(noise) %.2e  alpha(step) %.2e'%(np.abs(a-ref)[...,:3].max(), np.abs(a-ref)[...,3].max()))
    cases = {
      'cfg1/2 noise20': (synth.make_views(synth.workload('cfg1',scale=2.0), noise=20.0), 5),
      'cfg1 full': (synth.make_views(synth.workload('cfg1')), 5),
      'cfg1 noise40': (synth.make_views(synth.workload('cfg1'), noise=40.0), 5),
      'cfg3/8 L6 noise10': (synth.make_views(synth.workload('cfg3',scale=8.0), noise=10.0), 6),
    }
    for name,(regs,L) in cases.items():
        patches, pl = rs.build_patches(regs,'multiband',max_resolution=1e9)
        ref = rs.multiband([(w.copy(),m.copy(),s) for w,m,s in patches], pl.shape, L)
        for factors in ([2,4,4,4,4],[1,4,4,4,4],[2,2,4,4,4],[4,4,4,4,4],[2,4,8,8,8],[1,2,4,4,4]):
            got = multiband_approx([(w.copy(),m.copy(),s) for w,m,s in patches], pl.shape, L, factors)
            d = np.abs(got.astype(int)-ref.astype(int))
            print(name, factors, 'max',d.max(),'n>1',(d>1).sum(),'n>0 %.4f'%((d>0).mean()),'psnr %.1f'%psnr(got,ref))
