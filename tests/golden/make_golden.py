"""Generates tests/golden/oracle_vectors.npz: small seeded input/output vectors of the float64 oracle.

These are ORACLE-generated regression vectors (the reference cannot be executed here — see oracle/__init__.py); they
pin the oracle against accidental edits and give the GPU tests a committed fixture to compare with.
Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dense as od, philox as op, sngp as og  # noqa: E402


def main():
    rng = np.random.default_rng(20261019)
    f32 = lambda a: np.float32(a)  # noqa: E731
    out = {}
    E, n, I, O = 2, 4, 16, 8
    B = E * n
    x, w = f32(rng.standard_normal((B, I))), f32(rng.standard_normal((I, O)) / 4)
    alpha, gamma, bias = f32(rng.standard_normal((E, I))), f32(rng.standard_normal((E, O))), f32(rng.standard_normal((E, O)))
    gy = f32(rng.standard_normal((B, O)))
    out.update(be_x=x, be_w=w, be_alpha=alpha, be_gamma=gamma, be_bias=bias, be_gy=gy)
    y, _ = od.batch_ensemble_fwd(x, w, alpha, gamma, bias, "relu")
    g = od.batch_ensemble_bwd(x, w, alpha, gamma, bias, "relu", gy)
    out.update(be_y=y, **{f"be_{k}": v for k, v in g.items()})

    mu, rho, eps = f32(rng.standard_normal((I, O)) / 4), f32(rng.standard_normal((I, O)) * 0.2 - 2), f32(rng.standard_normal((I, O)))
    b1 = f32(rng.standard_normal(O))
    s_in, s_out = f32(rng.choice([-1.0, 1.0], (B, I))), f32(rng.choice([-1.0, 1.0], (B, O)))
    out.update(bn_mu=mu, bn_rho=rho, bn_eps=eps, bn_bias=b1, bn_sign_in=s_in, bn_sign_out=s_out)
    y, _ = od.reparam_fwd(x, mu, rho, eps, b1, "relu")
    g = od.reparam_bwd(x, mu, rho, eps, b1, "relu", gy)
    out.update(rp_y=y, **{f"rp_{k}": v for k, v in g.items()})
    y, _ = od.flipout_fwd(x, mu, rho, eps, s_in, s_out, b1, "relu")
    g = od.flipout_bwd(x, mu, rho, eps, s_in, s_out, b1, "relu", gy)
    out.update(fl_y=y, **{f"fl_{k}": v for k, v in g.items()})
    out.update(kl=np.float64(od.normal_kl(mu, rho, 0.5, 0.7, 0.25)))

    mu_r, mu_s = f32(rng.choice([-1.0, 1.0], (E, I))), f32(rng.choice([-1.0, 1.0], (E, O)))
    rho_r, rho_s = f32(rng.standard_normal((E, I)) * 0.2 - 1), f32(rng.standard_normal((E, O)) * 0.2 - 1)
    eps_r, eps_s = f32(rng.standard_normal((B, I))), f32(rng.standard_normal((B, O)))
    eps_r[0, 0] = 90.0
    out.update(r1_mu_r=mu_r, r1_mu_s=mu_s, r1_rho_r=rho_r, r1_rho_s=rho_s, r1_eps_r=eps_r, r1_eps_s=eps_s)
    y, _, _, _ = od.rank1_fwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu")
    g = od.rank1_bwd(x, w, mu_r, rho_r, eps_r, mu_s, rho_s, eps_s, bias, "relu", gy)
    out.update(r1_y=y, **{f"r1_{k}": v for k, v in g.items()})

    d, D = 12, 24
    xs, W, bb = f32(rng.standard_normal((B, d))), f32(rng.standard_normal((d, D))), f32(rng.uniform(0, 2 * np.pi, D))
    phi, _ = og.rff(og.layer_norm(xs), W, bb, np.sqrt(2.0 / D))
    P = og.precision_update_tf(np.zeros((D, D)), phi, 0.9)
    S = og.covariance_tf(P, 1e-2)
    out.update(sn_x=xs, sn_w=W, sn_b=bb, sn_phi=phi, sn_precision=P, sn_cov=S,
               sn_var=og.predictive_variance(phi, S, 1e-2), sn_full=og.predictive_covariance_tf(phi, S, 1e-2))
    wsn, u, v = f32(rng.standard_normal((I, O))), f32(rng.standard_normal(O)), f32(rng.standard_normal(I))
    for nm, fn in (("tf", og.spectral_norm_tf), ("jax", og.spectral_norm_jax)):
        wn, uh, vh, sg = fn(wsn, u, v, 2, 0.95, True)
        out.update(**{f"sn_{nm}_w": wn, f"sn_{nm}_u": uh.reshape(-1), f"sn_{nm}_v": vh.reshape(-1), f"sn_{nm}_sigma": np.float64(sg)})
    out.update(sn_w_in=wsn, sn_u_in=u, sn_v_in=v)
    out.update(philox_normal=op.normal(2026, 3, 64), philox_sign=op.sign(2026, 3, 64), philox_uniform=op.uniform(2026, 3, 64))
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
