"""numpy model of the blocked one-sided Jacobi (round-robin over 8-row blocks, rotations from the
block Gram) to count sweeps under different preconditioners.  Scratch analysis tool."""
import numpy as np, sys

def rr_pairs(n, rd):
    out = []
    for k in range(n // 2):
        if k == 0: a, b = n - 1, rd
        else:
            a = (rd + k) % (n - 1); b = (rd - k) % (n - 1)
        out.append((min(a, b), max(a, b)))
    return out

def block_jacobi(X, b=8, max_sweeps=40, stop=3e-8, inner='one'):
    X = X.copy(); p = X.shape[0]
    nb = -(-p // b); nb += nb & 1
    if nb * b > p:
        X = np.vstack([X, np.zeros((nb * b - p, X.shape[1]))])
    hist = []
    for sw in range(max_sweeps):
        mx = 0.0
        for ph in range(nb - 1):
            for (bi, bj) in rr_pairs(nb, ph):
                idx = np.r_[bi * b:(bi + 1) * b, bj * b:(bj + 1) * b]
                T = X[idx]
                G = T @ T.T
                W = np.eye(2 * b)
                # rotation rounds on G
                def rot(i, j):
                    nonlocal mx
                    a, bb, c = G[i, i], G[j, j], G[i, j]
                    if a <= 0 or bb <= 0 or c == 0: return
                    rel = abs(c) / np.sqrt(a * bb)
                    mx = max(mx, rel)
                    if rel <= 1e-15 * 16: return
                    tau = (bb - a) / (2 * c)
                    t = np.sign(tau) / (abs(tau) + np.sqrt(1 + tau * tau)) if tau != 0 else 1.0
                    cs = 1 / np.sqrt(1 + t * t); sn = cs * t
                    R = np.array([[cs, -sn], [sn, cs]])
                    G[[i, j], :] = R @ G[[i, j], :]
                    G[:, [i, j]] = G[:, [i, j]] @ R.T
                    W[[i, j], :] = R @ W[[i, j], :]
                reps = 1 if inner == 'one' else 3
                for _ in range(reps):
                    if ph == 0 or inner == 'full':
                        for h in range(2):
                            for rd in range(b - 1):
                                for (i, j) in rr_pairs(b, rd): rot(i + h * b, j + h * b)
                    for rd in range(b):
                        for i in range(b): rot(i, b + (i + rd) % b)
                X[idx] = W @ T
        hist.append(mx)
        if mx <= stop: break
    return sw + 1, hist, X[:p]

def make(m, c, parts, decade, rng):
    rp = c // parts
    A = rng.standard_normal((m, c))
    w = np.repeat(10.0 ** (-decade * np.arange(parts)), rp)
    B = np.linalg.qr(rng.standard_normal((c, c)))[0]
    return (A * w) @ B

if __name__ == '__main__':
    rng = np.random.default_rng(0)
    c = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    parts = c // 32
    decade = float(sys.argv[2]) if len(sys.argv) > 2 else 1.5
    M = make(c * 8, c, parts, decade, rng)
    R = np.linalg.qr(M, mode='r')
    sv = np.linalg.svd(R, compute_uv=False)
    def report(name, X, **kw):
        n, hist, Y = block_jacobi(X, **kw)
        s = np.sort(np.linalg.norm(Y, axis=1))[::-1]
        print(f"{name:34s} sweeps {n:2d}  relerr sv {np.max(np.abs(s - sv) / sv):.1e}  hist " + ' '.join(f'{h:.0e}' for h in hist))
    report('rows of R', R)
    report('rows of R^T', R.T)
    # sorted columns (norm pivoting), then QR
    perm = np.argsort(-np.linalg.norm(M, axis=0))
    Rp = np.linalg.qr(M[:, perm], mode='r')
    report('rows of R (col-norm sorted)', Rp)
    report('rows of R^T (col-norm sorted)', Rp.T)
    # second QR: R^T = Q1 R1
    R1 = np.linalg.qr(R.T, mode='r')
    report('rows of R1 (2nd QR of R^T)', R1)
    report('rows of R1^T', R1.T)
    R2 = np.linalg.qr(R1.T, mode='r')
    report('rows of R2 (3rd QR)', R2)
    report('rows of R2^T', R2.T)
    # rows sorted by norm
    o = np.argsort(-np.linalg.norm(R1.T, axis=1))
    report('rows of R1^T sorted by norm', R1.T[o])
    report('rows of R, full inner', R, inner='full')
    report('rows of R1^T, full inner', R1.T, inner='full')
