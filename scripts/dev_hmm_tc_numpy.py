"""Dev prototype (NumPy): the fragment algebra of the fp64 tensor-core HMM kernel (csrc/hmm_tc.cu) with a warp emulated as
32 lanes and mma.m8n8k4 as an explicit function -- pins the index maps (state relabelling that makes the accumulator layout
of one step the A-operand layout of the next, the P / P' fragment tables, the time-contraction of the transition counts)
before any CUDA is written.  Fragment layout (PTX ISA; the same convention csrc/slds_glue.cu runs on the GPU):
g = lane >> 2, q = lane & 3;  A[g][q], B[q][g], C[g][2q], C[g][2q+1]."""
import numpy as np

LANES = np.arange(32)
G, Q = LANES >> 2, LANES & 3


def dmma(c, a, b):
    """c (32,2) += A (8x4) @ B (4x8) with lane (g,q) holding A[g][q] = a[lane], B[q][g] = b[lane], C[g][2q+e] = c[lane][e]."""
    A = np.zeros((8, 4)); Bm = np.zeros((4, 8))
    A[G, Q] = a
    Bm[Q, G] = b
    C = A @ Bm
    c[:, 0] += C[G, 2 * Q]
    c[:, 1] += C[G, 2 * Q + 1]


def state_of(nt, e):           # vector layout: lane (g,q) holds, for sequence g, states 8 nt + 2 q + e
    return 8 * nt + 2 * Q + e


def k_state(s):                # A-operand of slab s = vector element [s >> 1][s & 1]: state 8 (s>>1) + 2 q + (s&1)
    return 8 * (s >> 1) + 2 * Q + (s & 1)


def to_vec(X):                 # X (8 sequences, 32 states) -> v[nt][e][lane]
    return np.array([[X[G, state_of(nt, e)] for e in range(2)] for nt in range(4)])


def from_vec(v):
    X = np.zeros((8, 32))
    for nt in range(4):
        for e in range(2):
            X[G, state_of(nt, e)] = v[nt][e]
    return X


def matvec(v, Pm):
    """out[g][j] = sum_i X[g][i] Pm[i][j] in the vector layout, 32 mma."""
    out = np.zeros((4, 32, 2))
    for nt in range(4):
        for s in range(8):
            frag = Pm[k_state(s), 8 * nt + G]          # B[q][g] = Pm[state_k(s,q)][8 nt + g]
            dmma(out[nt], v[s >> 1][s & 1], frag)
    return np.array([[out[nt][:, e] for e in range(2)] for nt in range(4)])


def etrans_time_contraction(Ah, Wh):
    """sum_t Ah[t] (x) Wh[t] for ONE sequence by mma with the contraction over 4 time steps: (T,32),(T,32) -> (32,32)."""
    T = Ah.shape[0]
    acc = np.zeros((4, 4, 32, 2))
    for t0 in range(0, T, 4):
        tt = t0 + Q
        ok = tt < T
        tc = np.minimum(tt, T - 1)
        for it in range(4):
            a = np.where(ok, Ah[tc, 8 * it + G], 0.0)            # A[g][q] = Ah[t0+q][8 it + g]
            for jt in range(4):
                b = np.where(ok, Wh[tc, 8 * jt + G], 0.0)        # B[q][g] = Wh[t0+q][8 jt + g]
                dmma(acc[it][jt], a, b)
    O = np.zeros((32, 32))
    for it in range(4):
        for jt in range(4):
            for e in range(2):
                O[8 * it + G, 8 * jt + 2 * Q + e] = acc[it][jt][:, e]
    return O


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    X, P = rng.standard_normal((8, 32)), rng.standard_normal((32, 32))
    assert np.allclose(from_vec(to_vec(X)), X)
    assert np.allclose(from_vec(matvec(to_vec(X), P)), X @ P), "forward matvec"
    assert np.allclose(from_vec(matvec(to_vec(X), P.T)), X @ P.T), "backward matvec (P' fragments)"
    # two steps chained without any re-layout
    assert np.allclose(from_vec(matvec(matvec(to_vec(X), P), P)), X @ P @ P)
    for T in (1, 3, 4, 9):
        Ah, Wh = rng.standard_normal((T, 32)), rng.standard_normal((T, 32))
        assert np.allclose(etrans_time_contraction(Ah, Wh), Ah.T @ Wh), T
    print("fragment algebra ok")


# ---------------------------------------------------------------------------------------------------------------------
# the whole sweep of csrc/hmm_tc.cu for one warp (8 sequences), statement by statement, against the oracle
# ---------------------------------------------------------------------------------------------------------------------
def sweep_warp(log_pi, log_A, node):
    """node (8,T,K), K <= 32.  Returns logZ (8), Ez (8,T,K), Ez0, Etrans (8,K,K) like hmm_tc_sweep_kernel + hmm_tc_trans_kernel."""
    Bw, T, K = node.shape
    assert Bw == 8
    gmax = float(np.float32(log_A.max()))
    Pp = np.zeros((32, 32)); Pp[:K, :K] = np.exp(log_A - gmax)
    inm = np.array([[state_of(nt, e) < K for e in range(2)] for nt in range(4)])          # (4,2,32)

    def load_row(arr, t, fill):      # arr (8,T,K) -> vector layout
        X = np.full((8, 32), fill); X[:, :K] = arr[:, t]
        return to_vec(X)

    def row_sum(v):                  # per-lane value = the row's total (quad reduction)
        tot = from_vec(v).sum(1)
        return tot[G]

    def row_maxf(v):
        return np.float32(from_vec(v)).max(1).astype(np.float64)[G]

    wsA, wsW = np.zeros((8, T, 32)), np.zeros((8, T, 32))
    Ez = np.zeros((8, T, 32))

    def store_row(dst, t, v):
        X = from_vec(v)
        dst[:, t, :K] = X[:, :K]

    x = load_row(node, 0, -np.inf)
    lp = np.full(32, -np.inf); lp[:K] = log_pi
    x = x + to_vec(np.tile(lp, (8, 1)))
    m = row_maxf(x)
    e_ = np.where(inm, np.exp(x - m), 0.0)
    logZ = m.copy()
    cprod = np.ones(32)
    for t in range(1, T):
        xn1 = load_row(node, t, -np.inf)
        d = matvec(e_, Pp)
        c = row_sum(e_); invc = 1.0 / c
        m = row_maxf(xn1)
        ah = e_ * invc
        e_ = np.where(inm, d * invc * np.exp(xn1 - m), 0.0)
        store_row(wsA, t - 1, ah)
        logZ += m
        cprod *= c
        if (t & 7) == 7:
            logZ += np.log(cprod); cprod[:] = 1.0
    c = row_sum(e_)
    cprod *= c
    logZ += np.log(cprod) + (T - 1) * gmax
    e_ = e_ / c
    store_row(Ez, T - 1, e_)
    if T > 1:
        bet = np.where(inm, 1.0, 0.0)
        xn = load_row(node, T - 1, -np.inf)
        an = to_vec(wsA[:, T - 2])
        for t in range(T - 2, -1, -1):
            if t > 0:
                xp = load_row(node, t, -np.inf)
                ap = to_vec(wsA[:, t - 1])
            m = row_maxf(xn)
            w = np.where(inm, np.exp(xn - m) * bet, 0.0)
            bu = matvec(w, Pp.T)
            prd = an * bu
            s = row_sum(prd); invs = 1.0 / s
            prd = prd * invs; w = w * invs; bet = bu * invs
            store_row(Ez, t, prd)
            store_row(wsW, t, w)
            if t > 0:
                xn, an = xp, ap
    Et = np.zeros((8, K, K))
    for r in range(8):
        O = etrans_time_contraction(wsA[r, :T - 1], wsW[r, :T - 1]) if T > 1 else np.zeros((32, 32))
        Et[r] = Pp[:K, :K] * O[:K, :K]
    lane0 = np.arange(8) * 4
    return logZ[lane0], Ez[:, :, :K], Ez[:, 0, :K], Et


if __name__ == "__main__":
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import hmm_numpy as HO
    rng = np.random.default_rng(1)
    for K, T in ((32, 21), (17, 9), (25, 1), (32, 2), (20, 8)):
        lp, lA = HO.rand_hmm_natparam(K, rng)
        lA = lA + rng.standard_normal((K, 1))
        node = 3.0 * rng.standard_normal((8, T, K))
        w_logZ, (w0, wt, wz) = HO.hmm_estep((lp, lA, node))
        logZ, Ez, Ez0, Et = sweep_warp(lp, lA, node)
        err = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
        assert err(logZ, w_logZ) < 1e-12 and err(Ez, wz) < 1e-11 and err(Ez0, w0) < 1e-11, (K, T)
        if T > 1:
            assert err(Et, wt) < 1e-11, (K, T)
    print("sweep emulation ok")


# ---------------------------------------------------------------------------------------------------------------------
# v2: FOUR warps per group of 8 sequences (warp w owns n-tile w: states 8 w + 2 q + e), the vector exchanged through shared
# memory once per step, every row normaliser applied one step LATE so that a step needs a single barrier.
# Mirrors hmm_tc_sweep4_kernel statement by statement; "smem" arrays are written in one phase and read in the next.
# ---------------------------------------------------------------------------------------------------------------------
def sweep_cta4(log_pi, log_A, node):
    Bw, T, K = node.shape
    assert Bw == 8
    gmax = float(np.float32(log_A.max()))
    Pp = np.zeros((32, 32)); Pp[:K, :K] = np.exp(log_A - gmax)
    W4 = range(4)
    st = lambda w, e: 8 * w + 2 * Q + e                       # (32,) states of warp w, element e
    inm = [[st(w, e) < K for e in range(2)] for w in W4]

    def tile_mv(vall, Pm, w):                                 # own n-tile of vall @ Pm, two accumulator chains
        c = np.zeros((2, 32, 2))
        for s in range(8):
            dmma(c[s & 1], vall[s >> 1][s & 1], Pm[k_state(s), 8 * w + G])
        return (c[0] + c[1]).T                                # (2, 32): element e, lane

    def ld(arr, t, w, fill):                                  # own tile of row t: (2, 32)
        X = np.full((8, 32), fill); X[:, :K] = arr[:, t]
        return np.array([X[G, st(w, e)] for e in range(2)])

    def quad_tot(x):                                          # (32,) -> per-lane total over the lane's quad
        return x.reshape(8, 4).sum(1)[G]

    def quad_max(x):
        return x.reshape(8, 4).max(1)[G]

    wsA, wsW, Ez = np.zeros((8, T, 32)), np.zeros((8, T, 32)), np.zeros((8, T, 32))

    def st_row(dst, t, w, v):                                 # own tile store
        for e in range(2):
            ok = inm[w][e]
            dst[G[ok], t, st(w, e)[ok]] = v[e][ok]

    lp = np.full(32, -np.inf); lp[:K] = log_pi
    # ---- forward ----
    vec = np.zeros((2, 4, 2, 32)); psum = np.zeros((2, 4, 32)); pmax = np.full((2, 4, 32), -np.inf)
    x0 = [ld(node, 0, w, -np.inf) + np.array([lp[st(w, e)] for e in range(2)]) for w in W4]
    pm0 = np.array([quad_max(np.float32(np.maximum(x0[w][0], x0[w][1])).astype(np.float64)) for w in W4])   # phase A
    m0 = pm0.max(0)                                                                                            # after barrier
    e_own = [np.where(inm[w], np.exp(x0[w] - m0), 0.0) for w in W4]
    logZ = m0.copy()
    cprod = np.ones(32)
    for w in W4:                                              # publish message 0 (+ partial sums, + partial max of node 1)
        vec[1, w] = e_own[w]
        psum[1, w] = quad_tot(e_own[w][0] + e_own[w][1])
        if T > 1:
            xn = ld(node, 1, w, -np.inf)
            pmax[1, w] = quad_max(np.float32(np.maximum(xn[0], xn[1])).astype(np.float64))
    for t in range(1, T):
        buf = t & 1
        new_e = []
        c = psum[buf].sum(0); m = pmax[buf].max(0)            # after the barrier: everybody's partials
        invc = 1.0 / c
        vall = vec[buf]                                       # (4,2,32)
        for w in W4:
            d = tile_mv(vall, Pp, w)
            xn1 = ld(node, t, w, -np.inf)
            ah = e_own[w] * invc
            st_row(wsA, t - 1, w, ah)
            new_e.append(np.where(inm[w], d * invc * np.exp(xn1 - m), 0.0))
        logZ += m
        cprod *= c
        if (t & 7) == 7:
            logZ += np.log(cprod); cprod[:] = 1.0
        for w in W4:
            e_own[w] = new_e[w]
            vec[buf ^ 1, w] = e_own[w]
            psum[buf ^ 1, w] = quad_tot(e_own[w][0] + e_own[w][1])
            if t + 1 < T:
                xn = ld(node, t + 1, w, -np.inf)
                pmax[buf ^ 1, w] = quad_max(np.float32(np.maximum(xn[0], xn[1])).astype(np.float64))
    c = psum[T & 1].sum(0)
    cprod *= c
    logZ += np.log(cprod) + (T - 1) * gmax
    for w in W4:
        st_row(Ez, T - 1, w, e_own[w] / c)
    # ---- backward (lagged normaliser) ----
    if T > 1:
        U = np.zeros((2, 4, 2, 32)); ps = np.zeros((2, 4, 32)); pmx = np.full((2, 4, 32), -np.inf)
        # prologue: U_{T-1} = exp(x_{T-1} - m) * 1, s_{T-1} := 1
        xl = [ld(node, T - 1, w, -np.inf) for w in W4]
        mm = np.array([quad_max(np.float32(np.maximum(xl[w][0], xl[w][1])).astype(np.float64)) for w in W4]).max(0)
        b0 = (T - 2) & 1
        for w in W4:
            U[b0, w] = np.where(inm[w], np.exp(xl[w] - mm), 0.0)
            ps[b0, w] = 0.25
            xn = ld(node, T - 2, w, -np.inf)
            pmx[b0, w] = quad_max(np.float32(np.maximum(xn[0], xn[1])).astype(np.float64))
        prd_prev = [None] * 4; uw_prev = [None] * 4
        for t in range(T - 2, -1, -1):
            buf = t & 1
            s_next = ps[buf].sum(0); inv = 1.0 / s_next       # s_{t+1}
            m = pmx[buf].max(0)                               # max of node_t
            Uall = U[buf]
            newU = []
            for w in W4:
                if t < T - 2:                                 # finish step t+1 now that s_{t+1} is known
                    st_row(Ez, t + 1, w, prd_prev[w] * inv)
                    st_row(wsW, t + 1, w, uw_prev[w] * inv)   # W_{t+1} = U_{t+2} / s_{t+2} / s_{t+1}
                Bt = tile_mv(Uall, Pp.T, w) * inv             # bu_t (own tile)
                an = np.array([np.where(inm[w][e], wsA[G, t, st(w, e)], 0.0) for e in range(2)])
                prd_prev[w] = an * Bt
                uw_prev[w] = Uall[w] * inv                    # w_{t+1} = U_{t+1} / s_{t+1}, own tile
                xt = ld(node, t, w, -np.inf)
                newU.append(np.where(inm[w], np.exp(xt - m) * Bt, 0.0))
            for w in W4:
                U[buf ^ 1, w] = newU[w]
                ps[buf ^ 1, w] = quad_tot(prd_prev[w][0] + prd_prev[w][1])
                if t > 0:
                    xn = ld(node, t - 1, w, -np.inf)
                    pmx[buf ^ 1, w] = quad_max(np.float32(np.maximum(xn[0], xn[1])).astype(np.float64))
        s0 = ps[1].sum(0); inv0 = 1.0 / s0                    # buffer (0 ^ 1)
        for w in W4:
            st_row(Ez, 0, w, prd_prev[w] * inv0)
            st_row(wsW, 0, w, uw_prev[w] * inv0)
    Et = np.zeros((8, K, K))
    for r in range(8):
        O = etrans_time_contraction(wsA[r, :T - 1], wsW[r, :T - 1]) if T > 1 else np.zeros((32, 32))
        Et[r] = Pp[:K, :K] * O[:K, :K]
    lane0 = np.arange(8) * 4
    return logZ[lane0], Ez[:, :, :K], Ez[:, 0, :K], Et


if __name__ == "__main__":
    rng = np.random.default_rng(2)
    for K, T in ((32, 21), (17, 9), (25, 1), (32, 2), (20, 8), (32, 3)):
        lp, lA = HO.rand_hmm_natparam(K, rng)
        lA = lA + rng.standard_normal((K, 1))
        node = 3.0 * rng.standard_normal((8, T, K))
        w_logZ, (w0, wt, wz) = HO.hmm_estep((lp, lA, node))
        logZ, Ez, Ez0, Et = sweep_cta4(lp, lA, node)
        err = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
        assert err(logZ, w_logZ) < 1e-12 and err(Ez, wz) < 1e-11 and err(Ez0, w0) < 1e-11, (K, T, err(logZ, w_logZ), err(Ez, wz))
        if T > 1:
            assert err(Et, wt) < 1e-11, (K, T, err(Et, wt))
    print("4-warp sweep emulation ok")


# ---------------------------------------------------------------------------------------------------------------------
# v3: EIGHT warps per group of 8 sequences -- warps 2w and 2w+1 both form n-tile w and keep element e = warp & 1 of it (one
# state per lane).  Only the partial sums / maxima change (8 per row instead of 4); checked here against the 4-warp statement.
# ---------------------------------------------------------------------------------------------------------------------
def partials_8_equal_4(rng):
    v = rng.standard_normal((4, 2, 32))                       # a published vector [tile][element][lane]
    quad = lambda x: x.reshape(8, 4).sum(1)[G]
    p4 = np.array([quad(v[w][0] + v[w][1]) for w in range(4)]).sum(0)
    p8 = np.array([quad(v[wi >> 1][wi & 1]) for wi in range(8)]).sum(0)
    assert np.allclose(p4, p8)
    # and the per-lane state of warp wi is 8 (wi >> 1) + 2 q + (wi & 1): every state of the 8 x 32 block exactly once
    seen = np.zeros((8, 32), dtype=int)
    for wi in range(8):
        seen[G, 8 * (wi >> 1) + 2 * Q + (wi & 1)] += 1
    assert (seen == 1).all()


if __name__ == "__main__":
    partials_8_equal_4(np.random.default_rng(3))
    print("8-warp partials ok")
