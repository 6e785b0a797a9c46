import mpmath as mp, numpy as np
mp.mp.dps = 60
A = (mp.pi/2)**2 * mp.mpf('1.0001')
def P(z):
    if z == 0: return -mp.mpf(1)/6
    r = mp.sqrt(z)
    return (mp.sin(r) - r) / (r*z)
def chebfit(N):
    # interpolate at Chebyshev nodes on [0, A], return monomial coefficients in z
    k = N + 1
    nodes = [ (A/2)*(1 + mp.cos(mp.pi*(2*j+1)/(2*k))) for j in range(k)]
    M = mp.matrix(k, k); b = mp.matrix(k, 1)
    for i, z in enumerate(nodes):
        for j in range(k): M[i, j] = z**j
        b[i] = P(z)
    c = mp.lu_solve(M, b)
    return [c[j] for j in range(k)]
def remez_like(N, iters=6):
    # minimise weighted error w(z)*(P - Phat), w = z * r/sin(r)  (relative error of sin), via iterated reweighted exchange (simple Remez)
    k = N + 2
    xs = [ (A/2)*(1 - mp.cos(mp.pi*j/(k-1))) for j in range(k)]
    def w(z):
        if z == 0: return mp.mpf(0)
        r = mp.sqrt(z); return z * r / mp.sin(r)
    c = None
    for it in range(iters):
        M = mp.matrix(k, k); b = mp.matrix(k, 1)
        for i, z in enumerate(xs):
            for j in range(N+1): M[i, j] = z**j
            wz = w(z) if z != 0 else mp.mpf('1e-30')
            M[i, N+1] = (-1)**i / wz
            b[i] = P(z)
        sol = mp.lu_solve(M, b)
        c = [sol[j] for j in range(N+1)]
        # find extrema of weighted error on a fine grid
        grid = [A*mp.mpf(i)/4000 for i in range(4001)]
        err = [w(z)*(P(z) - sum(c[j]*z**j for j in range(N+1))) for z in grid]
        # pick k alternating extrema
        ext = []
        sgn = 0
        for i in range(len(grid)):
            e = err[i]
            s = 1 if e > 0 else -1
            if not ext or s != sgn:
                ext.append(i); sgn = s
            elif abs(e) > abs(err[ext[-1]]):
                ext[-1] = i
        if len(ext) < k: break
        # keep the k largest consecutive
        while len(ext) > k:
            if abs(err[ext[0]]) < abs(err[ext[-1]]): ext.pop(0)
            else: ext.pop()
        xs = [grid[i] for i in ext]
        emax = max(abs(e) for e in err)
    return c, emax
for N in (7, 8):
    c, emax = remez_like(N)
    print("N", N, "weighted max rel err of sin ~", mp.nstr(emax, 5))
    cd = [float(x) for x in c]
    print("  coeffs:", ", ".join(repr(x) for x in cd))
    # test in double arithmetic (numpy, emulate fma-less Horner; fma differences are tiny)
    rng = np.random.default_rng(1)
    r = rng.uniform(-np.pi/2, np.pi/2, 200000)
    z = r*r
    p = np.full_like(z, cd[-1])
    for cc in cd[-2::-1]: p = p*z + cc
    s = r + (r*z)*p
    ref = np.array([float(mp.sin(mp.mpf(x))) for x in r[:20000]])
    refm = [mp.sin(mp.mpf(x)) for x in r[:20000]]
    ulp = np.spacing(np.abs(ref))
    e = np.array([abs(float(mp.mpf(float(si)) - rm)) for si, rm in zip(s[:20000], refm)]) / ulp
    print("  max ulp err (no-fma Horner, numpy):", e.max(), "mean", e.mean())
