"""Developer prototype (numpy): candidate interior-point variants on the oracle's assembled QP, in the sparse form the
oracle solves and in the condensed form the CUDA kernel solves.  Used to design the solver changes of round 2
(static regularisation of the cone block, refinement against the unregularised system, infeasibility certificate)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from common import wl


def get_qp(cfg_name, state, ee, steps=0):
    o = common.make_oracle(cfg_name, state)
    for _ in range(steps):
        o.solve(state, 0.0, ee, real_time=True)
    o.assemble(state, 0.0, ee)
    return o.qp()


def condense(qp, N):
    A = qp["A"].toarray(); P = qp["P"].toarray(); q = qp["q"]; b = qp["ub"]; eq = qp["is_eq"]
    n = A.shape[1]; nx = 12 * (N + 1); nu = n - nx
    dyn = np.arange(nx)                      # dynamics rows are the first 12(N+1) rows
    Ax, Au = A[dyn][:, :nx], A[dyn][:, nx:]
    Phi = -np.linalg.solve(Ax, Au); phi = np.linalg.solve(Ax, b[dyn])
    T = np.vstack([Phi, np.eye(nu)]); t0 = np.concatenate([phi, np.zeros(nu)])
    H = T.T @ P @ T; g = T.T @ (P @ t0 + q); c0 = 0.5 * t0 @ P @ t0 + q @ t0
    rest = np.arange(nx, A.shape[0])
    Ar = A[rest]; Cfull = Ar @ T; dfull = b[rest] - Ar @ t0; eqr = eq[rest]
    nzrow = np.abs(Ar).sum(1) > 0
    ine = (~eqr) & nzrow
    return dict(H=0.5 * (H + H.T), g=g, C=Cfull[ine], d=dfull[ine], E=Cfull[eqr], e=dfull[eqr], c0=c0, T=T, t0=t0)


def ipm_condensed(Q, eps=1e-8, delta=1e-8, max_iter=50, tol=1e-8, refine=2, verbose=False, aug_refine=True):
    H, g, C, d, E, e = Q["H"], Q["g"], Q["C"], Q["d"], Q["E"], Q["e"]
    nu, m, me = H.shape[0], C.shape[0], E.shape[0]
    EtE = E.T @ E / delta
    def factor(w):
        K = H + C.T @ (w[:, None] * C) + EtE
        try:
            return np.linalg.cholesky(K)
        except np.linalg.LinAlgError:
            return None
    def csolve(L, r):
        y = np.linalg.solve(L, r)
        return np.linalg.solve(L.T, y)
    L = factor(np.ones(m))
    u = csolve(L, -g + C.T @ d + E.T @ e / delta)
    s = d - C @ u
    shift = max(0.0, -1.5 * s.min())
    v = np.maximum(s + shift, 1e-2)
    s = v.copy(); lam = v.copy()
    xi = (v * v).sum(); s = s + 0.5 * xi / v.sum(); lam = lam + 0.5 * xi / s.sum()
    nu_eq = np.zeros(me)
    nrm_q = max(1.0, np.abs(g).max()); nrm_d = max(1.0, np.abs(d).max(), np.abs(e).max() if me else 0)
    status = "MaxIter"; hist = []
    for it in range(max_iter + 1):
        Hu = H @ u
        pobj = u @ (0.5 * Hu + g)
        rd = Hu + g + C.T @ lam + E.T @ nu_eq
        rp = C @ u + s - d
        re = E @ u - e
        n_rd, n_rp, n_re = np.abs(rd).max(), np.abs(rp).max(), (np.abs(re).max() if me else 0.0)
        sdl = s @ lam; mu = sdl / m
        gs = max(1.0, abs(pobj + Q["c0"]))
        # Farkas certificate
        by = d @ lam + e @ nu_eq
        aty = np.abs(C.T @ lam + E.T @ nu_eq).max()
        hist.append((n_rd, n_rp, n_re, sdl, by, aty))
        if verbose:
            print(it, f"rd {n_rd:.2e} rp {n_rp:.2e} re {n_re:.2e} gap {sdl:.2e} by {by:.3e} aty {aty:.2e} lam {lam.max():.2e}")
        if n_rd <= tol * nrm_q and n_rp <= tol * nrm_d and n_re <= tol * nrm_d and sdl <= tol * gs:
            status = "Solved"; break
        if by < -1e-8 and aty <= 1e-8 * max(1.0, np.abs(lam).max()) * (-by) / max(1.0, nrm_d):
            status = "PrimalInfeasible"; break
        if it == max_iter:
            break
        D = s / lam
        w = 1.0 / (D + eps)
        L = factor(w)
        if L is None:
            status = "Other"; break
        def newton(rc):
            # unreduced system in (du, dl, dnu):  H du + C'dl + E'dnu = -rd ; C du - D dl = -rp + rc/lam ; E du - delta dnu = -re
            r1 = -rd; r2 = -rp + rc / lam; r3 = -re
            def solve_reg(a1, a2, a3):
                rhs = a1 + C.T @ (w * a2) + E.T @ a3 / delta     # C'(D+eps)^-1 a2 ... sign: dl = w (C du - a2)
                # from C du - (D+eps) dl = a2 -> dl = w (C du - a2); H du + C' w (C du - a2) + E' (E du - a3)/delta = a1
                rhs = a1 + C.T @ (w * a2) + E.T @ a3 / delta
                du = csolve(L, rhs)
                dl = w * (C @ du - a2)
                dn = (E @ du - a3) / delta
                return du, dl, dn
            du, dl, dn = solve_reg(r1, r2, r3)
            for _ in range(refine):
                e1 = r1 - (H @ du + C.T @ dl + E.T @ dn)
                e2 = r2 - (C @ du - (D if aug_refine else D + eps) * dl)
                e3 = r3 - (E @ du - delta * dn)
                cu, cl, cn = solve_reg(e1, e2, e3)
                du += cu; dl += cl; dn += cn
            ds = (-rc - s * dl) / lam
            return du, dl, dn, ds
        def max_step(ds, dl):
            a = 1e300
            neg = ds < 0
            if neg.any(): a = min(a, (-s[neg] / ds[neg]).min())
            neg = dl < 0
            if neg.any(): a = min(a, (-lam[neg] / dl[neg]).min())
            return a
        du, dl, dn, ds = newton(s * lam)
        a_aff = min(1.0, max_step(ds, dl))
        mu_aff = ((s + a_aff * ds) @ (lam + a_aff * dl)) / m
        sig = (mu_aff / mu) ** 3
        du, dl, dn, ds = newton(s * lam + ds * dl - sig * mu)
        alpha = min(1.0, 0.99 * max_step(ds, dl))
        u = u + alpha * du; s = s + alpha * ds; lam = lam + alpha * dl; nu_eq = nu_eq + alpha * dn
    return dict(status=status, iters=it, u=u, hist=hist)


if __name__ == "__main__":
    cfg_name = "a1_configuration"; cfg = wl.CONFIGS[cfg_name]
    B = int(os.environ.get("B", 64))
    states, t0, ee = wl.batched_trot_inputs(cfg, 4096, seed=0)
    import collections
    cnt = collections.Counter(); cnt_o = collections.Counter(); pairs = collections.Counter()
    for b in range(B):
        qp = get_qp(cfg_name, states[b], ee[b])
        Q = condense(qp, cfg["num_nodes"])
        r = ipm_condensed(Q, eps=float(os.environ.get("EPS", 1e-8)), refine=int(os.environ.get("REFINE", 2)))
        o = common.make_oracle(cfg_name, states[b]); so = o.solve(states[b], 0.0, ee[b], real_time=True)
        cnt[r["status"]] += 1; cnt_o[so] += 1; pairs[(r["status"], so)] += 1
        if os.environ.get("V"): print(b, r["status"], r["iters"], "oracle", so, o.qp_solution()["iters"])
    print(cnt, cnt_o); print(pairs)


def hsde_condensed(Q, eps=1e-8, delta=1e-8, max_iter=60, tol_feas=1e-8, tol_gap=1e-8, refine=2, verbose=False, tol_inf=1e-8):
    """Clarabel-style homogeneous self-dual embedding on the condensed QP; equality rows E u = e tau keep the proximal
    (static-regularisation) treatment: they are rows of the KKT system with -delta on the diagonal."""
    H, g, C, d, E, e = Q["H"], Q["g"], Q["C"], Q["d"], Q["E"], Q["e"]
    nu, m, me = H.shape[0], C.shape[0], E.shape[0]
    EtE = E.T @ E / delta
    def factor(w):
        K = H + eps * np.eye(nu) + C.T @ (w[:, None] * C) + EtE
        try:
            return np.linalg.cholesky(K)
        except np.linalg.LinAlgError:
            return None
    def csolve(L, r):
        return np.linalg.solve(L.T, np.linalg.solve(L, r))
    def kkt_solve(L, w, D, a1, a2, a3):
        # H du + C'dl + E'dn = a1 ; C du - D dl = a2 ; E du - delta dn = a3     (D = s/lam), regularised solve + refinement
        def reg(b1, b2, b3):
            du = csolve(L, b1 + C.T @ (w * b2) + E.T @ b3 / delta)
            return du, w * (C @ du - b2), (E @ du - b3) / delta
        du, dl, dn = reg(a1, a2, a3)
        for _ in range(refine):
            e1 = a1 - (H @ du + C.T @ dl + E.T @ dn)
            e2 = a2 - (C @ du - D * dl)
            e3 = a3 - (E @ du)
            cu, cl, cn = reg(e1, e2, e3)
            du += cu; dl += cl; dn += cn
        return du, dl, dn
    # initial point
    one = np.ones(m)
    L = factor(one)
    u, z, y = kkt_solve(L, one, one, -g, d, e)
    s = -z
    def shift(v):
        a = v.min()
        return v + (1.0 - a) if a < 1e-8 else v
    s = shift(s); z = shift(z)
    tau = 1.0; kap = 1.0
    nb = max(np.abs(d).max(), np.abs(e).max() if me else 0.0); nq = np.abs(g).max()
    status = "MaxIter"
    for it in range(max_iter + 1):
        Hu = H @ u
        uHu = u @ Hu
        rx = Hu + C.T @ z + E.T @ y + g * tau
        rz = C @ u + s - d * tau
        re = E @ u - e * tau
        rt = kap + g @ u + d @ z + e @ y + uHu / tau
        mu = (s @ z + tau * kap) / (m + 1)
        # termination on the de-homogenised point
        xs, zs, ss = np.abs(u).max() / tau, max(np.abs(z).max(), np.abs(y).max() if me else 0) / tau, np.abs(s).max() / tau
        pc = (0.5 * uHu / tau + g @ u) / tau
        dc = (-(d @ z + e @ y) - 0.5 * uHu / tau) / tau
        res_p = max(np.abs(rz).max(), np.abs(re).max() if me else 0) / tau / max(1.0, nb + xs + ss)
        res_d = np.abs(rx).max() / tau / max(1.0, nq + xs + zs)
        gap_abs = abs(pc - dc); gap_rel = gap_abs / max(1.0, min(abs(pc + Q["c0"]), abs(dc + Q["c0"])))
        bz = d @ z + e @ y
        aty = np.abs(C.T @ z + E.T @ y).max()
        if verbose:
            print(it, f"resp {res_p:.2e} resd {res_d:.2e} gap {gap_abs:.2e}/{gap_rel:.2e} tau {tau:.2e} kap {kap:.2e} mu {mu:.2e} bz {bz:.2e} aty {aty:.2e}")
        if res_p < tol_feas and res_d < tol_feas and (gap_abs < tol_gap or gap_rel < tol_gap):
            status = "Solved"; break
        if bz < -tol_inf and aty < -tol_inf * max(1.0, xs * tau + zs * tau) * bz:
            status = "PrimalInfeasible"; break
        if it == max_iter:
            break
        D = s / z
        w = 1.0 / (D + eps)
        L = factor(w)
        if L is None:
            status = "Other"; break
        x1, z1, y1 = kkt_solve(L, w, D, -g, d, e)
        xi = u / tau
        Hxi = Hu / tau
        den = kap / tau - g @ x1 - d @ z1 - e @ y1 + (x1 - xi) @ (H @ (x1 - xi)) - x1 @ (H @ x1)
        def step(dx_, dz_, de_, dt_, dk_, ds_):
            x2, z2, y2 = kkt_solve(L, w, D, -dx_, -dz_ + ds_ / z, -de_)
            dtau = (dt_ - dk_ / tau + (2 * Hxi + g) @ x2 + d @ z2 + e @ y2) / den
            du = x2 + dtau * x1; dz = z2 + dtau * z1; dy = y2 + dtau * y1
            dsv = (-ds_ - s * dz) / z
            dkap = (-dk_ - kap * dtau) / tau
            return du, dz, dy, dsv, dtau, dkap
        def max_step(dsv, dz, dtau, dkap):
            a = 1.0
            for v, dv in ((s, dsv), (z, dz)):
                neg = dv < 0
                if neg.any(): a = min(a, (-v[neg] / dv[neg]).min())
            if dtau < 0: a = min(a, -tau / dtau)
            if dkap < 0: a = min(a, -kap / dkap)
            return a
        du, dz, dy, dsv, dtau, dkap = step(rx, rz, re, rt, kap * tau, s * z)
        a_aff = max_step(dsv, dz, dtau, dkap)
        sig = (1 - a_aff) ** 3
        du, dz, dy, dsv, dtau, dkap = step((1 - sig) * rx, (1 - sig) * rz, (1 - sig) * re, (1 - sig) * rt,
                                           kap * tau + dkap * dtau - sig * mu, s * z + dsv * dz - sig * mu)
        alpha = 0.99 * max_step(dsv, dz, dtau, dkap)
        if verbose:
            lin2 = C @ du + dsv - d * dtau + (1 - sig) * rz
            lin3 = E @ du - e * dtau + (1 - sig) * re
            print("    alpha", alpha, "sig", sig, "eq2 err", np.abs(lin2).max(), "eq3 err", np.abs(lin3).max() if me else 0)
        u = u + alpha * du; z = z + alpha * dz; y = y + alpha * dy; s = s + alpha * dsv; tau += alpha * dtau; kap += alpha * dkap
    return dict(status=status, iters=it, u=u / tau)


def equilibrate(Q):
    """u = Du ut with Du = 1/sqrt(diag H); rows of C, E scaled to unit infinity norm."""
    H, g, C, d, E, e = Q["H"], Q["g"], Q["C"], Q["d"], Q["E"], Q["e"]
    Du = 1.0 / np.sqrt(np.diag(H))
    Ht = Du[:, None] * H * Du[None, :]
    Ct = C * Du[None, :]; Et = E * Du[None, :]
    rc = 1.0 / np.abs(Ct).max(1); re_ = 1.0 / np.abs(Et).max(1) if E.shape[0] else np.zeros(0)
    return dict(H=Ht, g=Du * g, C=rc[:, None] * Ct, d=rc * d, E=re_[:, None] * Et, e=re_ * e, c0=Q["c0"]), Du, rc, re_


def hsde_kernel_form(Q, eps=1e-10, delta=1e-10, max_iter=50, tol_feas=1e-8, tol_gap=1e-8, refine=0, verbose=False, tol_inf=1e-8):
    """The iteration exactly as csrc/bgg_ipm.cu and oracle/qp_ipm.cpp run it: unrefined regularised solves for the constant
    and the step right-hand sides, d tau from those, the TOTAL direction refined against the unregularised system with the
    analytic residuals (-eps dz, -delta dy) when refine > 0 (default 0: see oracle/qp_ipm.cpp), ds from the complementarity equation."""
    H, g, C, d, E, e = Q["H"], Q["g"], Q["C"], Q["d"], Q["E"], Q["e"]
    nu, m, me = H.shape[0], C.shape[0], E.shape[0]
    EtE = E.T @ E / delta
    def factor(w):
        return np.linalg.cholesky(H + eps * np.eye(nu) + C.T @ (w[:, None] * C) + EtE)
    def csolve(L, r):
        return np.linalg.solve(L.T, np.linalg.solve(L, r))
    def reg(L, w, b1, b2, b3):
        du = csolve(L, b1 + C.T @ (w * b2) + E.T @ b3 / delta)
        return du, w * (C @ du - b2), (E @ du - b3) / delta
    w = np.full(m, 1.0 / (1.0 + eps))
    L = factor(w)
    u, z, y = reg(L, w, -g, d, e)
    s = -z
    def shift(v):
        a = v.min()
        return v + (1.0 - a) if a < 1e-8 else v
    s = shift(s); z = shift(z)
    tau = 1.0; kap = 1.0
    nb = max(1.0, np.abs(d).max(), np.abs(e).max() if me else 0.0); nq = max(1.0, Q.get("nrm_q", np.abs(g).max()))
    status = "MaxIter"
    for it in range(max_iter + 1):
        Hu = H @ u; uHu = u @ Hu
        aty_v = C.T @ z + E.T @ y
        rx = Hu + aty_v + g * tau
        rz = C @ u + s - d * tau
        re = E @ u - e * tau
        bz = d @ z + e @ y
        rt = kap + g @ u + bz + uHu / tau
        mu = (s @ z + tau * kap) / (m + 1)
        pc = (0.5 * uHu / tau + g @ u) / tau; dc = (-bz - 0.5 * uHu / tau) / tau
        res_p = max(np.abs(rz).max(), np.abs(re).max() if me else 0) / tau; res_d = np.abs(rx).max() / tau
        gap = abs(pc - dc); gs = max(1.0, min(abs(pc + Q["c0"]), abs(dc + Q["c0"])))
        aty = np.abs(aty_v).max(); zn = max(1.0, np.abs(z).max())
        if verbose:
            print(it, f"resp {res_p:.2e} (rz {np.abs(rz).max():.2e} @{np.abs(rz).argmax()} re {np.abs(re).max() if me else 0:.2e}) resd {res_d:.2e} gap {gap:.2e} tau {tau:.2e} kap {kap:.2e} mu {mu:.2e} zmax {z.max():.2e} ymax {np.abs(y).max():.2e} smin {s.min():.1e}")
        if res_p <= tol_feas * nb and res_d <= tol_feas * nq and gap <= tol_gap * gs:
            status = "Solved"; break
        if bz < -tol_inf and aty <= tol_inf * zn * (-bz):
            status = "PrimalInfeasible"; break
        if it == max_iter:
            break
        D = s / z; w = 1.0 / (D + eps)
        L = factor(w)
        x1, z1, y1 = reg(L, w, -g, d, e)
        den = kap / tau - g @ x1 - d @ z1 - e @ y1 + uHu / tau ** 2 - 2 * (Hu @ x1) / tau
        def step(scale, d_s, d_kap, do_refine):
            a1 = -scale * rx; a2 = -scale * rz + d_s / z; a3 = -scale * re
            x2, z2, y2 = reg(L, w, a1, a2, a3)
            dtau = (scale * rt - d_kap / tau + g @ x2 + d @ z2 + e @ y2 + 2 * (Hu @ x2) / tau) / den
            du = x2 + dtau * x1; dz = z2 + dtau * z1; dy = y2 + dtau * y1
            if verbose and do_refine:
                print("      pre-refine eq2", np.abs(C @ du - D * dz - (a2 + dtau * d)).max(), "x2-only", np.abs(C @ x2 - D * z2 - a2).max(), "x1-only", np.abs(C @ x1 - D * z1 - d).max(), "eq1", np.abs((a1 - dtau * g) - (H @ du + C.T @ dz + E.T @ dy)).max(), "|du|", np.abs(du).max(), "|dz|", np.abs(dz).max())
            for _ in range(refine if do_refine else 0):
                e1 = (a1 - dtau * g) - (H @ du + C.T @ dz + E.T @ dy)
                cu, cz, cy = reg(L, w, e1, -eps * dz, -delta * dy)
                du += cu; dz += cz; dy += cy
                if verbose: print("      post-refine eq2", np.abs(C @ du - D * dz - (a2 + dtau * d)).max(), "|cz|", np.abs(cz).max(), "|cu|", np.abs(cu).max(), "eq1", np.abs((a1 - dtau * g) - (H @ du + C.T @ dz + E.T @ dy)).max())
            dsv = (-d_s - s * dz) / z      # from the complementarity equation: keeps the relative accuracy of tiny slacks
            dkap = (-d_kap - kap * dtau) / tau
            return du, dz, dy, dsv, dtau, dkap
        def max_step(dsv, dz, dtau, dkap):
            a = 1.0
            for v, dv in ((s, dsv), (z, dz)):
                neg = dv < 0
                if neg.any(): a = min(a, (-v[neg] / dv[neg]).min())
            if dtau < 0: a = min(a, -tau / dtau)
            if dkap < 0: a = min(a, -kap / dkap)
            return a
        du, dz, dy, dsv, dtau, dkap = step(1.0, s * z, kap * tau, False)
        a_aff = max_step(dsv, dz, dtau, dkap)
        sig = (1 - a_aff) ** 3
        du, dz, dy, dsv, dtau, dkap = step(1 - sig, s * z + dsv * dz - sig * mu, kap * tau + dkap * dtau - sig * mu, True)
        alpha = 0.99 * max_step(dsv, dz, dtau, dkap)
        if verbose:
            lin2 = C @ du + dsv - d * dtau + (1 - sig) * rz
            lin3 = E @ du - e * dtau + (1 - sig) * re
            print("    alpha", alpha, "sig", sig, "eq2 err", np.abs(lin2).max(), "eq3 err", np.abs(lin3).max() if me else 0)
        u = u + alpha * du; z = z + alpha * dz; y = y + alpha * dy; s = s + alpha * dsv; tau += alpha * dtau; kap += alpha * dkap
    return dict(status=status, iters=it, u=u / tau)
