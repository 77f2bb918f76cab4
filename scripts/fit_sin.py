import mpmath as mp
mp.mp.dps = 60
Z = (mp.pi/2 * mp.mpf('1.002'))**2
def g(z):
    if z == 0: return mp.mpf(-1)/6
    r = mp.sqrt(z)
    return (mp.sin(r)/r - 1)/z
def fit(deg):
    n = deg+1
    # Chebyshev nodes on [0,Z]
    nodes = [ (Z/2)*(1+mp.cos(mp.pi*(2*k+1)/(2*n))) for k in range(n)]
    A = mp.matrix(n,n); b = mp.matrix(n,1)
    for i,z in enumerate(nodes):
        for j in range(n): A[i,j] = z**j
        b[i] = g(z)
    c = mp.lu_solve(A,b)
    return [c[j] for j in range(n)]
def remez(deg, iters=8):
    n = deg+2
    xs = [ (Z/2)*(1-mp.cos(mp.pi*k/(n-1))) for k in range(n)]
    # weight: error in sin relative: r*z*(g-Q)/sin(r)  -> weight w(z)= r*z/sin(r)
    def w(z):
        if z==0: return mp.mpf(0)
        r=mp.sqrt(z); return r*z/mp.sin(r)
    for it in range(iters):
        A = mp.matrix(n,n); b = mp.matrix(n,1)
        for i,z in enumerate(xs):
            for j in range(deg+1): A[i,j] = z**j
            ww = w(z) if z!=0 else mp.mpf('1e-30')
            A[i,deg+1] = (-1)**i / ww
            b[i] = g(z)
        sol = mp.lu_solve(A,b)
        c = [sol[j] for j in range(deg+1)]
        E = sol[deg+1]
        # find extrema of weighted error on dense grid
        N=4000
        grid=[Z*k/N for k in range(N+1)]
        err=[ (sum(c[j]*z**j for j in range(deg+1))-g(z))*w(z) for z in grid]
        # local extrema
        ext=[]
        for k in range(N+1):
            l = err[k-1] if k>0 else None; r_ = err[k+1] if k<N else None
            e=err[k]
            if (l is None or abs(e)>=abs(l) ) and (r_ is None or abs(e)>=abs(r_)) :
                if (l is None or (e-l)*(1 if e>0 else -1)>=0) and (r_ is None or (e-r_)*(1 if e>0 else -1)>=0):
                    ext.append((grid[k],e))
        # pick alternating n extrema with largest
        # merge same-sign neighbours
        merged=[]
        for z,e in ext:
            if merged and (merged[-1][1]>0)==(e>0):
                if abs(e)>abs(merged[-1][1]): merged[-1]=(z,e)
            else: merged.append((z,e))
        if len(merged)<n: break
        while len(merged)>n:
            if abs(merged[0][1])<abs(merged[-1][1]): merged.pop(0)
            else: merged.pop()
        xs=[m[0] for m in merged]
    maxerr=max(abs(e) for e in err)
    return c, maxerr
for deg in (7,8,9):
    c, me = remez(deg)
    print("deg", deg, "weighted max rel err", mp.nstr(me,5), "2^-53=",mp.nstr(mp.mpf(2)**-53,5))
    for cc in c: print("   ", mp.nstr(cc, 25), float(cc).hex())
