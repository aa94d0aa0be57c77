"""Prototype (numpy) of the pair-incompatibility certificate for the subset search."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from macaque_3d_pose_estimation_b200 import synth
from oracle import cameragroup as og, fixtures, camera_math as cm

def mu_of(cam):
    k1,k2,p1,p2,k3 = (list(cam.dist)+[0]*5)[:5]
    u = np.linspace(0, 1e4, 2000001)
    s = 1+k1*u+k2*u*u+k3*u**3
    gd = 1+3*k1*u+5*k2*u*u+7*k3*u**3
    P = abs(p1)+abs(p2)
    return float(np.min(np.minimum(s, gd) - 6*P*np.sqrt(u)))

def distort(cam, xy):
    k1,k2,p1,p2,k3 = (list(cam.dist)+[0]*5)[:5]
    x,y = xy[...,0], xy[...,1]
    r2 = x*x+y*y
    cd = 1+k1*r2+k2*r2*r2+k3*r2**3
    xd = x*cd+2*p1*x*y+p2*(r2+2*x*x)
    yd = y*cd+p1*(r2+2*y*y)+2*p2*x*y
    return np.stack([xd*cam.K[0,0]+cam.K[0,2], yd*cam.K[1,1]+cam.K[1,2]], -1)

def main():
    C = 8
    seed = 20261020
    cams = fixtures.cams_from_dicts(synth.make_rig(C, "pinhole", seed=seed))
    mus = [mu_of(c) for c in cams]
    print("mu", np.round(mus,3))
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    X = synth.make_tracks(F, 2, seed=seed).reshape(-1,3)
    p2d = synth.corrupt(og.project(cams, X), seed=seed, p_outlier=0.2, p_missing=0.1)
    N = p2d.shape[1]
    out, picked, p2, errs, sidx, nev = og.triangulate_ransac(cams, p2d, return_stats=True)
    print("N", N, "mean neval", nev.mean(), "selected", (sidx>=0).mean())
    U = og.undistort_points(cams, p2d)
    R = [cm.rodrigues(c.rvec) for c in cams]; t = [c.tvec for c in cams]
    fmin = [min(abs(c.K[0,0]), abs(c.K[1,1])) for c in cams]
    delta = np.stack([np.linalg.norm(p2d[c]-distort(cams[c], U[c]), axis=-1) for c in range(C)])
    print("delta max", np.nanmax(delta), "median", np.nanmedian(delta))
    T = 0.5
    bad = np.zeros((N, C, C), bool)
    for a in range(C):
        for b in range(a+1, C):
            Rba = R[b] @ R[a].T
            tba = t[b] - Rba @ t[a]
            tx = np.array([[0,-tba[2],tba[1]],[tba[2],0,-tba[0]],[-tba[1],tba[0],0]])
            E = tx @ Rba
            E /= np.linalg.norm(E)
            g22 = np.linalg.norm(E[:2,:2], 2)
            xa = np.concatenate([U[a], np.ones((N,1))], 1)
            xb = np.concatenate([U[b], np.ones((N,1))], 1)
            Ea = xa @ E.T           # E xa
            Etb = xb @ E            # E^T xb
            Fv = np.abs(np.sum(xb*Ea, 1))
            A = (np.abs(Etb[:,0])+np.abs(Etb[:,1]))/(mus[a]*fmin[a])
            B = (np.abs(Ea[:,0])+np.abs(Ea[:,1]))/(mus[b]*fmin[b])
            k = (~np.isnan(p2d[:,:,0])).sum(0)
            rho = T*k*(1+1e-9)+1e-6
            D = rho + delta[a]+delta[b]
            rhs = np.maximum(A,B)*D + g22/(mus[a]*fmin[a]*mus[b]*fmin[b])*D*D/4
            bd = Fv > rhs*(1+1e-9)
            bad[:,a,b] = bd; bad[:,b,a] = bd
    # simulate the skip search
    ncand = np.zeros(N, int); visits = np.zeros(N,int); viol = 0
    for n in range(N):
        V = [c for c in range(C) if not np.isnan(p2d[c,n,0])]
        k = len(V)
        if k < 2 or k <= 2: continue
        # was decided at s=0?
        if sidx[n] == 0: continue
        # local bit b <-> camera V[k-1-b]
        cam_of_bit = [V[k-1-b] for b in range(k)]
        s = 1; nsub = 1<<k; found=False
        target = sidx[n]
        while s < nsub:
            visits[n]+=1
            kept = [b for b in range(k) if not (s>>b)&1]
            jb = -1
            for b in kept:
                for a in kept:
                    if a > b and bad[n, cam_of_bit[a], cam_of_bit[b]]:
                        jb = max(jb, b)
            if jb >= 0:
                s2 = ((s>>jb)|1)<<jb
                if target >= 0 and s <= target < s2: viol += 1
                s = s2; continue
            if len(kept) < 2: s+=1; continue
            ncand[n]+=1
            if s == target: found=True; break
            s+=1
    und = (sidx!=0)
    print("violations", viol, "cand/pt (undecided)", ncand[und].mean(), "visits/pt", visits[und].mean(), "max cand", ncand.max(),
          "hist", np.bincount(ncand)[:12])
    # property test: random X, never e_a+e_b < rho when bad
    rng = np.random.default_rng(1)
    for trial in range(20):
        Xr = np.concatenate([X + rng.normal(0, 30, X.shape), rng.normal(0, 3000, X.shape), X*rng.uniform(-3,3,(N,1))])
        pr = og.project(cams, Xr)          # C, 3N, 2
        for a in range(C):
            for b in range(a+1,C):
                bb = np.tile(bad[:,a,b],3)
                ea = np.linalg.norm(np.tile(p2d[a],(3,1))-pr[a],axis=1)
                eb = np.linalg.norm(np.tile(p2d[b],(3,1))-pr[b],axis=1)
                kk = np.tile((~np.isnan(p2d[:,:,0])).sum(0),3)
                v = bb & (ea+eb < T*kk)
                if v.any(): print("PROPERTY VIOLATION", a,b, v.sum())
    print("property test done")
main()
