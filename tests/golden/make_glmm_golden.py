"""Generates tests/golden/glmm_*.json: known-answer vectors for the GLMM log-density and its
gradient, from an INDEPENDENT transcription of /root/reference/src/stan_files/continuous.stan
into torch (float64) with autograd doing the differentiation -- the same division of labour
as the reference (Stan program + reverse-mode AD, src/include/stan/model/gradient.hpp:21-35).

The reference itself cannot be executed in this image (needs R, Eigen, Boost, TBB), so these
vectors pin the oracle (oracle/oracle_glmm.c) and, through it, the CUDA path.  Every lpdf keeps
its normalising constants because the generated C++ hard-codes `<false>` (continuous.hpp:2460,
:853-909).  Run:  python tests/golden/make_glmm_golden.py
"""
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from stan4bart_b200.frontend import build_stan_data  # noqa: E402

torch.set_default_dtype(torch.float64)


def normal_lpdf(y, mu, sigma):
    # stan::math::normal_lpdf<false>: -0.5 z^2 - log(sigma) - 0.5 log(2 pi), summed
    z = (y - mu) / sigma
    n = y.numel()
    return (-0.5 * z * z).sum() - n * torch.log(sigma) - n * 0.5 * math.log(2 * math.pi)


def gamma_lpdf(y, alpha, beta=1.0):
    alpha = torch.as_tensor(alpha)
    return (alpha * math.log(beta) - torch.lgamma(alpha) + (alpha - 1) * torch.log(y) - beta * y).sum()


def beta_lpdf(y, a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return (torch.lgamma(a + b) - torch.lgamma(a) - torch.lgamma(b) + (a - 1) * torch.log(y) + (b - 1) * torch.log1p(-y)).sum()


def inv_gamma_lpdf(y, a, b):
    # stan::math::inv_gamma_lpdf<false>
    return (a * torch.log(b) - torch.lgamma(a) - (a + 1) * torch.log(y) - b / y).sum()


def student_t_lpdf(y, nu, mu, sigma):
    z = (y - mu) / sigma
    return (math.lgamma((nu + 1) / 2) - math.lgamma(nu / 2) - 0.5 * math.log(nu * math.pi) - math.log(sigma)
            - (nu + 1) / 2 * torch.log1p(z * z / nu)).sum()


def make_theta_L(p, dispersion, tau, scale, zeta, rho, z_T):
    """continuous.stan:2-59, including the scaled onion rows of blocks with more than two coefficients (note that the
    off-diagonal entries of row r + 1 are scaled with the standard deviation of row r, as the Stan program does)."""
    out = []
    zeta_mark = 0
    rho_mark = 0
    z_T_mark = 0
    for i, nc in enumerate(p):
        if nc == 1:
            out.append(tau[i] * scale[i] * dispersion)
        else:
            trace = (tau[i] * scale[i] * dispersion) ** 2 * nc
            pi = zeta[zeta_mark:zeta_mark + nc]
            pi = pi / pi.sum()
            zeta_mark += nc
            T = [[None] * nc for _ in range(nc)]
            std_dev = torch.sqrt(pi[0] * trace)
            T[0][0] = std_dev
            std_dev = torch.sqrt(pi[1] * trace)
            T21 = 2.0 * rho[rho_mark] - 1.0
            rho_mark += 1
            T[1][1] = std_dev * torch.sqrt(1.0 - T21 ** 2)
            T[1][0] = std_dev * T21
            for r in range(2, nc):          # stan r = 2 .. nc - 1, rp1 = r + 1 (1-based) -> row index r (0-based)
                T_row = z_T[z_T_mark:z_T_mark + r]
                scale_factor = torch.sqrt(rho[rho_mark] / (T_row * T_row).sum()) * std_dev
                z_T_mark += r
                std_dev = torch.sqrt(pi[r] * trace)
                for c in range(r):
                    T[r][c] = T_row[c] * scale_factor
                T[r][r] = torch.sqrt(1.0 - rho[rho_mark]) * std_dev
                rho_mark += 1
            for c in range(nc):             # vech, column major
                for r in range(c, nc):
                    out.append(T[r][c])
    return torch.stack(out) if out else torch.zeros(0)


def make_b(z_b, theta_L, p, l):
    b = []
    b_mark = 0
    th = 0
    for i, nc in enumerate(p):
        if nc == 1:
            b.append(theta_L[th] * z_b[b_mark:b_mark + l[i]])
            b_mark += l[i]
            th += 1
        else:
            rows_T = [[torch.zeros(()) for _ in range(nc)] for _ in range(nc)]
            for c in range(nc):
                for r in range(c, nc):
                    rows_T[r][c] = theta_L[th]
                    th += 1
            Tm = torch.stack([torch.stack(row) for row in rows_T])
            rows = []
            for j in range(l[i]):
                rows.append(Tm @ z_b[b_mark:b_mark + nc])
                b_mark += nc
            b.append(torch.cat(rows))
    return torch.cat(b) if b else torch.zeros(0)


def log_prob(sd, q, offset, y):
    """continuous.stan, default path, jacobian = true."""
    K, nq, t = sd.K, sd.q, sd.t
    p, l = list(sd.p), list(sd.l)
    len_rho = sum(p) - t
    len_conc = len(sd.concentration)
    pos = 0
    lp = torch.zeros(())
    nzb = sd.len_z_beta
    z_beta = q[pos:pos + nzb]; pos += nzb
    # parameters of the shrinkage priors, all <lower=0>: exp transform, Jacobian + u   (continuous.stan:265-270)
    hs = sd.hs
    def lower0(count):
        nonlocal pos, lp
        u = q[pos:pos + count]; pos += count
        lp = lp + u.sum()
        return torch.exp(u)
    glob = lower0(hs)
    local = [lower0(K) for _ in range(hs)]
    caux = lower0(1 if hs > 0 else 0)
    mix = lower0(K) if sd.prior_dist in (5, 6) else None
    ool = lower0(1) if sd.prior_dist == 6 else None
    z_b = q[pos:pos + nq]; pos += nq
    len_z_T = sum((pi - 2) * (pi - 1) for pi in p if pi > 2)        # continuous.stan:258
    z_T = q[pos:pos + len_z_T]; pos += len_z_T
    rho_u = q[pos:pos + len_rho]; pos += len_rho
    zeta_u = q[pos:pos + len_conc]; pos += len_conc
    tau_u = q[pos:pos + t]; pos += t
    rho = torch.sigmoid(rho_u)
    lp = lp + (torch.log(rho) + torch.log1p(-rho)).sum()          # lub_constrain(0, 1) Jacobian
    zeta = torch.exp(zeta_u); lp = lp + zeta_u.sum()
    tau = torch.exp(tau_u); lp = lp + tau_u.sum()
    if not sd.is_binary:
        aux_u = q[pos]
        aux_unscaled = torch.exp(aux_u); lp = lp + aux_u
        if sd.prior_dist_for_aux == 0:
            aux = aux_unscaled
        elif sd.prior_dist_for_aux <= 2:
            aux = sd.prior_scale_for_aux * aux_unscaled + sd.prior_mean_for_aux
        else:
            aux = sd.prior_scale_for_aux * aux_unscaled
        dispersion = aux
    else:
        aux = torch.ones(())
        dispersion = torch.ones(())
    pscale, pmean, pdf = torch.as_tensor(sd.prior_scale), torch.as_tensor(sd.prior_mean), torch.as_tensor(np.asarray(sd.prior_df, dtype=np.float64))
    if sd.prior_dist == 0:
        beta = z_beta
    elif sd.prior_dist == 1:
        beta = z_beta * pscale + pmean
    elif sd.prior_dist == 2:                                   # continuous.stan:146-158, :295-297
        z = z_beta; df = pdf
        z2 = z * z; z3 = z2 * z; z5 = z2 * z3; z7 = z2 * z5; z9 = z2 * z7
        df2 = df * df; df3 = df2 * df; df4 = df2 * df2
        cft = (z + (z3 + z) / (4 * df) + (5 * z5 + 16 * z3 + 3 * z) / (96 * df2)
               + (3 * z7 + 19 * z5 + 17 * z3 - 15 * z) / (384 * df3)
               + (79 * z9 + 776 * z7 + 1482 * z5 - 1920 * z3 - 945 * z) / (92160 * df4))
        beta = cft * pscale + pmean
    elif sd.prior_dist in (3, 4):                              # continuous.stan:123-143, :298-305
        c2 = sd.slab_scale ** 2 * caux[0]
        lam = local[0] * torch.sqrt(local[1])
        if sd.prior_dist == 4:
            lam = lam * (local[2] * torch.sqrt(local[3]))
        tau_g = glob[0] * torch.sqrt(glob[1]) * sd.global_prior_scale * aux
        lam2 = lam * lam
        lam_tilde = torch.sqrt(c2 * lam2 / (c2 + tau_g * tau_g * lam2))
        beta = z_beta * lam_tilde * tau_g
    elif sd.prior_dist == 5:
        beta = pmean + pscale * torch.sqrt(2 * mix) * z_beta
    elif sd.prior_dist == 6:
        beta = pmean + ool[0] * pscale * torch.sqrt(2 * mix) * z_beta
    else:                                                      # product_normal, continuous.stan:310-322
        parts, zp = [], 0
        for k in range(K):
            nn = int(sd.num_normals[k])
            parts.append(torch.prod(z_beta[zp:zp + nn]) * float(sd.prior_scale[k]) ** nn + float(sd.prior_mean[k]))
            zp += nn
        beta = torch.stack(parts) if parts else z_beta[:0]
    theta_L = make_theta_L(p, dispersion, tau, torch.as_tensor(sd.scale), zeta, rho, z_T)
    b = make_b(z_b, theta_L, p, l)
    # eta = offset + X beta + Z b   (CSR w, v, u)
    eta = torch.as_tensor(offset).clone()
    if K > 0:
        eta = eta + torch.as_tensor(np.ascontiguousarray(sd.X)) @ beta
    u = sd.u
    rows = np.repeat(np.arange(sd.N), np.diff(u))
    Zb = torch.zeros(sd.N).index_add(0, torch.as_tensor(rows), torch.as_tensor(sd.w) * b[torch.as_tensor(sd.v.astype(np.int64))])
    eta = eta + Zb
    actual_aux = aux if not sd.is_binary else torch.ones(())
    if getattr(sd, "weights", None) is None:
        lp = lp + normal_lpdf(torch.as_tensor(y), eta, actual_aux)
    else:
        # continuous.stan:358-366 (the sum of the log weights is dropped there)
        wt = torch.as_tensor(sd.weights)
        lp = lp + (-0.5 * sd.N * torch.log(6.283185307179586232 * actual_aux * actual_aux)
                   - 0.5 * (wt * (torch.as_tensor(y) - eta) ** 2).sum() / (actual_aux * actual_aux))
    if (not sd.is_binary) and sd.prior_dist_for_aux > 0 and sd.prior_scale_for_aux > 0:
        log_half = -0.693147180559945286
        if sd.prior_dist_for_aux == 1:
            lp = lp + normal_lpdf(aux_unscaled.reshape(1), torch.zeros(()), torch.ones(())) - log_half
        elif sd.prior_dist_for_aux == 2:
            lp = lp + student_t_lpdf(aux_unscaled.reshape(1), sd.prior_df_for_aux, 0.0, 1.0) - log_half
        else:
            lp = lp - aux_unscaled                                   # exponential_lpdf(. | 1)
    if sd.prior_dist >= 1:
        lp = lp + normal_lpdf(z_beta, torch.zeros(()), torch.ones(()))
    log_half = -0.693147180559945286
    def half_normal(x):
        return normal_lpdf(x, torch.zeros(()), torch.ones(())) - log_half
    if sd.prior_dist in (3, 4):                                # continuous.stan:381-401
        lp = lp + half_normal(local[0]) + inv_gamma_lpdf(local[1], 0.5 * pdf, 0.5 * pdf)
        if sd.prior_dist == 4:
            lp = lp + half_normal(local[2]) + inv_gamma_lpdf(local[3], 0.5 * pscale, 0.5 * pscale)
        lp = lp + half_normal(glob[0:1]) + inv_gamma_lpdf(glob[1:2], torch.tensor([0.5 * sd.global_prior_df]), torch.tensor([0.5 * sd.global_prior_df]))
        lp = lp + inv_gamma_lpdf(caux, torch.tensor([0.5 * sd.slab_df]), torch.tensor([0.5 * sd.slab_df]))
    elif sd.prior_dist in (5, 6):
        lp = lp - mix.sum()                                    # exponential_lpdf(. | 1)
        if sd.prior_dist == 6:
            nu = float(sd.prior_df[0])                         # chi_square_lpdf
            lp = lp + (-(nu / 2) * math.log(2.0) - math.lgamma(nu / 2) + (nu / 2 - 1) * torch.log(ool[0]) - ool[0] / 2)
    # decov_lp
    lp = lp + normal_lpdf(z_b, torch.zeros(()), torch.ones(()))
    if len_z_T:
        lp = lp + normal_lpdf(z_T, torch.zeros(()), torch.ones(()))
    pos_reg = 0
    pos_rho = 0
    for i in range(t):
        if p[i] > 1:
            nu = sd.regularization[pos_reg] + 0.5 * (p[i] - 2)
            pos_reg += 1
            shape1, shape2 = [nu], [nu]
            for j in range(2, p[i]):          # stan j = 2 .. p - 1
                nu -= 0.5
                shape1.append(0.5 * j)
                shape2.append(nu)
            lp = lp + beta_lpdf(rho[pos_rho:pos_rho + p[i] - 1], np.array(shape1), np.array(shape2))
            pos_rho += p[i] - 1
    delta = []
    for i in range(t):
        if p[i] > 1:
            delta += [sd.concentration[j] for j in range(p[i])]
    if len_conc:
        lp = lp + gamma_lpdf(zeta, torch.as_tensor(np.array(delta)))
    if t:
        lp = lp + gamma_lpdf(tau, torch.as_tensor(sd.shape))
    extras = torch.cat([glob] + local + [caux] + ([mix] if mix is not None else []) + ([ool] if ool is not None else []))
    return lp, dict(beta=beta, b=b, theta_L=theta_L, aux=aux, rho=rho, zeta=zeta, tau=tau, z_T=z_T, extras=extras)


def make_case(name, N, seed, binary, terms_spec, aux_prior=3, prior_dist=1, weighted=False, coef=None):
    rng = np.random.default_rng(seed)
    Xf = np.column_stack([rng.random(N), (rng.random(N) < 0.3).astype(float)])
    terms = []
    groups = []
    for (nlev, pc) in terms_spec:
        g = rng.integers(0, nlev, N)
        g[:nlev] = np.arange(nlev)          # every level present
        M = np.column_stack([np.ones(N)] + [rng.random(N) for _ in range(pc - 1)])
        terms.append((g, M))
        groups.append(dict(levels=int(nlev), g=g.tolist(), M=M.tolist()))
    y = rng.standard_normal(N) * 2 + 3 * Xf[:, 0]
    sd = build_stan_data(Xf, y, terms, is_binary=binary)
    if not binary:
        sd.prior_dist_for_aux = aux_prior
        if aux_prior in (1, 2):
            sd.prior_mean_for_aux = 0.4
            sd.prior_df_for_aux = 4.0
    sd.prior_dist = prior_dist
    if coef:
        for key, val in coef.items():
            setattr(sd, key, np.asarray(val) if isinstance(val, list) else val)
    offset = rng.standard_normal(N) * 0.5
    if binary:
        y = rng.standard_normal(N) + 0.3      # latents
        sd.y = y
    if weighted:
        sd.weights = rng.gamma(2.0, 0.5, N)
        sd.weights[:3] = [0.0, 1.0, 2.5]
    qs, lps, grads, was = [], [], [], []
    for k in range(3):
        q = torch.tensor(rng.uniform(-1.5, 1.5, sd.num_params), requires_grad=True)
        lp, tp = log_prob(sd, q, offset, y)
        lp.backward()
        qs.append(q.detach().numpy().tolist()); lps.append(float(lp)); grads.append(q.grad.numpy().tolist())
        wa = q.detach().numpy().tolist()
        # write_array order: params (constrained) then aux, beta, b, theta_L
        K, nq = sd.K, sd.q
        nzb = sd.len_z_beta
        qn = q.detach().numpy()
        ne = len(tp["extras"])
        cons = list(qn[:nzb]) + tp["extras"].tolist() + list(qn[nzb + ne:nzb + ne + nq]) + tp["z_T"].tolist() + tp["rho"].tolist() + tp["zeta"].tolist() + tp["tau"].tolist()
        if not binary:
            cons += [float(torch.exp(q[-1])), float(tp["aux"])]
        cons += tp["beta"].tolist() + tp["b"].tolist() + tp["theta_L"].tolist()
        was.append([float(v) for v in cons])
    case = dict(name=name, N=N, binary=binary, X_fixed=Xf.tolist(), y=np.asarray(y).tolist(), y_for_scaling=None, offset=offset.tolist(),
                groups=groups, aux_prior=aux_prior, prior_dist=prior_dist,
                **(dict(weights=sd.weights.tolist()) if weighted else {}), **(dict(coef=coef) if coef else {}),
                prior_scale=sd.prior_scale.tolist(), prior_scale_for_aux=sd.prior_scale_for_aux,
                prior_mean_for_aux=sd.prior_mean_for_aux, prior_df_for_aux=sd.prior_df_for_aux,
                q=qs, lp=lps, grad=grads, write_array=was)
    with open(os.path.join(HERE, f"glmm_{name}.json"), "w") as f:
        json.dump(case, f)
    print(name, "d =", sd.num_params, "lp =", lps)


if __name__ == "__main__":
    make_case("friedman_like", 60, 1, False, [(5, 2), (8, 1)])
    make_case("binary", 50, 2, True, [(4, 2), (6, 1)])
    make_case("intercepts_only", 40, 3, False, [(3, 1), (7, 1)], aux_prior=1)
    make_case("two_slopes_t_aux", 70, 4, False, [(4, 2), (5, 2), (3, 1)], aux_prior=2)
    make_case("no_ranef_flat_prior", 30, 5, False, [], aux_prior=0, prior_dist=0)
    make_case("three_coefficients", 80, 6, False, [(4, 3), (6, 1)])
    make_case("four_and_three_binary", 90, 7, True, [(3, 4), (5, 3), (4, 2)])
    make_case("weighted", 64, 8, False, [(5, 2), (6, 1)], weighted=True)
    make_case("weighted_binary", 55, 9, True, [(4, 3), (3, 1)], weighted=True)
    make_case("coef_student_t", 45, 10, False, [(4, 1)], prior_dist=2, coef=dict(prior_df=[3.0, 7.0]))
    make_case("coef_hs", 50, 11, False, [(4, 2)], prior_dist=3,
              coef=dict(prior_df=[1.0, 3.0], global_prior_df=1.0, global_prior_scale=0.05, slab_df=4.0, slab_scale=2.5))
    make_case("coef_hs_plus", 50, 12, False, [(5, 1)], prior_dist=4,
              coef=dict(prior_df=[1.0, 2.0], global_prior_df=2.0, global_prior_scale=0.1, slab_df=3.0, slab_scale=2.0))
    make_case("coef_laplace_binary", 48, 13, True, [(4, 1)], prior_dist=5)
    make_case("coef_lasso", 52, 14, False, [(3, 2)], prior_dist=6, coef=dict(prior_df=[2.0, 2.0]))
    make_case("coef_product_normal", 44, 15, False, [(4, 1)], prior_dist=7, coef=dict(num_normals=[2, 3]))
