"""Algorithm prototype for the MPC-CBF kernel (development aid, not shipped, not a test oracle).

Condensed (single-shooting) primal-dual interior-point method with exact Lagrangian Hessian,
derivatives from torch.autograd so that only the *algorithm* (step rule, barrier schedule,
regularisation) is under study here.  The CUDA kernel (scb_mpc.cuh) implements the same
iteration with hand-structured derivatives.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.mpc_cbf import OracleMPCCBF  # noqa: E402


def ipm(o, x_init, goal, u_prev, obs, max_iter=200, tol=1e-8, verbose=False, mu0=0.1):
    H, nu = o.H, o.nu
    n = H * nu
    z = torch.tensor(np.tile(np.asarray(u_prev, float), H))

    def fg(zz):
        return o.condensed(x_init, goal, u_prev, obs, zz)

    def derivs(zz, lam):
        zt = zz.clone().requires_grad_(True)
        J, g = fg(zt)
        gradJ = torch.autograd.grad(J, zt, retain_graph=True)[0]
        Jac = torch.autograd.functional.jacobian(lambda q: fg(q)[1], zz, vectorize=True)
        L = lambda q: fg(q)[0] - (lam * fg(q)[1]).sum()
        W = torch.autograd.functional.hessian(L, zz)
        return float(J), gradJ, g.detach(), Jac, W

    with torch.no_grad():
        J, g = fg(z)
    m = g.numel()
    s = torch.clamp(g, min=1e-2)
    mu = mu0
    lam = mu / s
    it_hist = []
    for it in range(max_iter):
        J, gradJ, g, Jac, W = derivs(z, lam)
        r_d = gradJ - Jac.T @ lam
        r_p = g - s
        comp = lam * s
        err0 = max(float(r_d.abs().max()), float(r_p.abs().max()), float(comp.abs().max()))
        errmu = max(float(r_d.abs().max()), float(r_p.abs().max()), float((comp - mu).abs().max()))
        if verbose:
            print(f"it {it:3d} J={J:.6f} rd={float(r_d.abs().max()):.2e} rp={float(r_p.abs().max()):.2e} comp={float(comp.max()):.2e} mu={mu:.1e}")
        if err0 <= tol:
            break
        # barrier update (Fiacco-McCormick, IPOPT-like constants)
        while errmu <= 10.0 * mu and mu > tol / 10:
            mu = max(tol / 10, min(0.2 * mu, mu ** 1.5))
            errmu = max(float(r_d.abs().max()), float(r_p.abs().max()), float((comp - mu).abs().max()))
        Sig = lam / s
        r_c = comp - mu
        Hred = W + Jac.T @ (Sig[:, None] * Jac)
        rhs = -r_d - Jac.T @ (r_c / s + Sig * r_p)
        delta = 0.0
        while True:
            try:
                Lc = torch.linalg.cholesky(Hred + delta * torch.eye(n))
                break
            except Exception:
                delta = 1e-4 if delta == 0 else delta * 10
        dz = torch.cholesky_solve(rhs[:, None], Lc)[:, 0]
        ds = Jac @ dz + r_p
        dlam = -(r_c + lam * ds) / s
        tau = max(0.99, 1 - mu)

        def ftb(v, dv):
            neg = dv < 0
            if not neg.any():
                return 1.0
            return float(min(1.0, (tau * (-v[neg] / dv[neg])).min()))
        ap, ad = ftb(s, ds), ftb(lam, dlam)
        # merit line search on (z, s): phi = J - mu sum log s + nu |g - s|_1
        nu_pen = max(1.0, float(lam.abs().max()) * 1.1)

        def phi(zz, ss):
            with torch.no_grad():
                Jv, gv = fg(zz)
            return float(Jv) - mu * float(torch.log(ss).sum()) + nu_pen * float((gv - ss).abs().sum())
        phi0 = phi(z, s)
        dphi = float(gradJ @ dz) - mu * float((ds / s).sum()) - nu_pen * float(r_p.abs().sum())
        a = ap
        nbt = 0
        while nbt < 12:
            if phi(z + a * dz, s + a * ds) <= phi0 + 1e-4 * a * min(dphi, 0.0) + 1e-12 * abs(phi0):
                break
            a *= 0.5; nbt += 1
        z = z + a * dz
        s = s + a * ds
        lam = lam + ad * dlam
        it_hist.append((a, ad, delta, nbt))
    return z.numpy().reshape(H, nu), dict(iters=it, err=err0, hist=it_hist, J=J)


if __name__ == "__main__":
    from safe_control_b200 import scenes
    M, H = 16, 8
    for dense in (False, True):
        sc = scenes.make_scene("DynamicUnicycle2D", 10, M, seed=4321, dense=dense)
        o = OracleMPCCBF(sc["spec"], num_obs=M, horizon=H)
        for i in range(10):
            k = int(sc["nobs"][i])
            obs = sc["OBS"][i][:k]
            goal = sc["goal"][i] if not dense else sc["X"][i][:2] + 1.5 * np.array([np.cos(sc["X"][i][2]), np.sin(sc["X"][i][2])])
            t = time.time()
            u, info = o.solve(sc["X"][i], goal, sc["u_prev"][i], obs)
            t_or = time.time() - t
            t = time.time()
            U, st = ipm(o, sc["X"][i], goal, sc["u_prev"][i], obs, verbose="-v" in sys.argv)
            t_ip = time.time() - t
            nbts = sum(h[3] for h in st["hist"]); regs = sum(h[2] > 0 for h in st["hist"])
            print(f"dense={dense} i={i} oracle u0={u} ok={info['success']} nit={info['nit']} J={info['fun']:.6f} | ipm u0={U[0]} it={st['iters']} err={st['err']:.1e} J={st['J']:.6f} bt={nbts} reg={regs} | du={np.abs(U - info['u_pred']).max():.2e}")
