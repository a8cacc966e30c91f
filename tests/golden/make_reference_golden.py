"""Generates tests/golden/reference_vectors.npz by EXECUTING THE UNMODIFIED REFERENCE SOURCE FILES

    /root/reference/edward2/jax/nn/{dense,random_feature,normalization,utils,heteroscedastic_lib}.py

under the NumPy-backed stand-ins of jax / flax in tests/golden/refstub/ (neither library is installed here; see
refstub/README.md).  Every arithmetic statement that runs is the reference's; the stand-ins evaluate it in float64.
The fixture pins `oracle/` (tests/test_oracle.py::test_reference_*) and the CUDA path (tests/test_gpu_reference_golden.py)
to the reference itself rather than to the oracle's own output.

Run here (the GPU box has no /root/reference):  python tests/golden/make_reference_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/edward2/jax/nn"
OUT = os.path.join(HERE, "reference_vectors.npz")


def load_reference():
    """Import the four reference files by path (their package __init__ pulls in unrelated modules)."""
    sys.path.insert(0, os.path.join(HERE, "refstub"))
    import flax  # noqa: F401  (stand-in)
    import jax  # noqa: F401  (stand-in)

    for name in ("edward2", "edward2.jax", "edward2.jax.nn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    mods = {}
    for name in ("utils", "dense", "normalization", "random_feature", "heteroscedastic_lib"):
        full = f"edward2.jax.nn.{name}"
        spec = importlib.util.spec_from_file_location(full, os.path.join(REF, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
        setattr(sys.modules["edward2.jax.nn"], name, m)
        mods[name] = m
    return mods


def main():
    mods = load_reference()
    import flax.linen as nn
    import jax

    dense, rf, norm, utils = mods["dense"], mods["random_feature"], mods["normalization"], mods["utils"]
    rng = np.random.default_rng(20261020)
    out = {}
    key = lambda i: {"params": jax.random.PRNGKey(i)}  # noqa: E731

    # ---- a1 DenseBatchEnsemble.__call__ (dense.py:238-272): 2-D and 3-D inputs, sign / N(1, 0.5) fast weights
    for tag, shape, sign in (("be2d", (12, 20), 0.5), ("be3d", (6, 5, 20), -0.5)):
        E, O = 3, 16
        layer = dense.DenseBatchEnsemble(features=O, ens_size=E, activation=nn.relu,
                                         alpha_init=utils.make_sign_initializer(sign),
                                         gamma_init=utils.make_sign_initializer(sign),
                                         bias_init=nn.initializers.normal(stddev=0.3))
        x = rng.standard_normal(shape)
        v = layer.init(key(1), x)
        y = layer.apply(v, x)
        p = v["params"]
        out.update({f"{tag}_x": x, f"{tag}_kernel": p["kernel"], f"{tag}_alpha": p["fast_weight_alpha"],
                    f"{tag}_gamma": p["fast_weight_gamma"], f"{tag}_bias": p["bias"], f"{tag}_y": y})

    # ---- a10 RandomFourierFeatures.__call__ (random_feature.py:183-214)
    for tag, scale in (("rff_unit", 1.0), ("rff_auto", None)):
        layer = rf.RandomFourierFeatures(features=48, feature_scale=scale)
        x = rng.standard_normal((10, 24)) * 2.0
        v = layer.init(key(2), x)
        out.update({f"{tag}_x": x, f"{tag}_kernel": v["random_features"]["kernel"],
                    f"{tag}_bias": v["random_features"]["bias"], f"{tag}_phi": layer.apply(v, x)})

    # ---- a11 LaplaceRandomFeatureCovariance (random_feature.py:255-404)
    D = 24
    phi1, phi2, phit = rng.standard_normal((16, D)) * 0.3, rng.standard_normal((16, D)) * 0.3, rng.standard_normal((9, D)) * 0.3
    lg1, lg2 = rng.standard_normal((16, 1)), rng.standard_normal((16, 1))
    out.update(cov_phi1=phi1, cov_phi2=phi2, cov_phit=phit, cov_logits1=lg1, cov_logits2=lg2)
    for tag, momentum, likelihood, ridge in (("cov_exact", None, "gaussian", 1.0), ("cov_mom", 0.9, "gaussian", 0.5),
                                             ("cov_logistic", None, "binary_logistic", 1.0),
                                             ("cov_poisson", 0.99, "poisson", 2.0)):
        layer = rf.LaplaceRandomFeatureCovariance(hidden_features=D, ridge_penalty=ridge, momentum=momentum,
                                                  likelihood=likelihood)
        v = layer.init(key(3), phi1, lg1)
        out[f"{tag}_p0"] = v["laplace_covariance"]["precision_matrix"]
        r, v1 = layer.apply(v, phi1, lg1, mutable=["laplace_covariance"])
        assert r is None  # random_feature_test.py:346: nothing is returned while the precision is being updated
        _, v2 = layer.apply(v1, phi2, lg2, mutable=["laplace_covariance"])
        out[f"{tag}_p1"] = v1["laplace_covariance"]["precision_matrix"]
        out[f"{tag}_p2"] = v2["laplace_covariance"]["precision_matrix"]
        out[f"{tag}_var"] = layer.apply(v2, phit, None, diagonal_only=True)
        out[f"{tag}_cov"] = layer.apply(v2, phit, None, diagonal_only=False)
        out[f"{tag}_ridge"] = np.float64(ridge)

    # ---- a10 + a11 RandomFeatureGaussianProcess end to end (random_feature.py:112-140): two training calls, one eval
    gp = rf.RandomFeatureGaussianProcess(features=3, hidden_features=32, hidden_kwargs=dict(feature_scale=None),
                                         covmat_kwargs=dict(ridge_penalty=1.0))
    xa, xb, xt = rng.standard_normal((20, 12)), rng.standard_normal((20, 12)), rng.standard_normal((7, 12))
    v = gp.init(key(4), xa)
    # give the output layer non-trivial values (init draws; bias init is zeros)
    v["params"]["output_layer"]["bias"] = rng.standard_normal(3) * 0.1
    state = {k: v[k] for k in v if k != "params"}
    (lg_a, cov_a), st1 = gp.apply({"params": v["params"], **state}, xa, mutable=["laplace_covariance"])
    assert cov_a is None
    (_, _), st2 = gp.apply({"params": v["params"], **{**state, **st1}}, xb, mutable=["laplace_covariance"])
    full = {"params": v["params"], **{**state, **st2}}
    logits, var, feats = gp.apply(full, xt, return_random_features=True)
    _, covfull = gp.apply(full, xt, return_full_covmat=True)
    out.update(gp_xa=xa, gp_xb=xb, gp_xt=xt, gp_rff_kernel=v["random_features"]["hidden_layer"]["kernel"],
               gp_rff_bias=v["random_features"]["hidden_layer"]["bias"],
               gp_out_kernel=v["params"]["output_layer"]["kernel"], gp_out_bias=v["params"]["output_layer"]["bias"],
               gp_logits_a=lg_a, gp_precision=st2["laplace_covariance"]["covmat_layer"]["precision_matrix"],
               gp_logits=logits, gp_var=var, gp_cov=covfull, gp_features=feats)

    # ---- a9 SpectralNormalization over nn.Dense (normalization.py:123-185): 1 and 3 power iterations, then eval
    for tag, iters, c, wscale in (("sn1", 1, 0.95, 1.0), ("sn3", 3, 0.6, 1.0), ("sn_small", 2, 0.95, 0.02)):
        layer = norm.SpectralNormalization(nn.Dense(14), iteration=iters, norm_multiplier=c)
        x = rng.standard_normal((9, 10))
        v = layer.init(key(5), x)
        v["params"]["Dense"]["kernel"] = v["params"]["Dense"]["kernel"] * wscale
        v["params"]["Dense"]["bias"] = rng.standard_normal(14) * 0.1
        out.update({f"{tag}_x": x, f"{tag}_w": v["params"]["Dense"]["kernel"], f"{tag}_b": v["params"]["Dense"]["bias"],
                    f"{tag}_u0": v["spectral_stats"]["u"], f"{tag}_v0": v["spectral_stats"]["v"],
                    f"{tag}_iters": np.int64(iters), f"{tag}_c": np.float64(c)})
        y, st = layer.apply(v, x, training=True, mutable=["spectral_stats", "intermediates"])
        out.update({f"{tag}_y": y, f"{tag}_u1": st["spectral_stats"]["u"], f"{tag}_v1": st["spectral_stats"]["v"],
                    f"{tag}_what": st["intermediates"]["w"][0]})
        v2 = {"params": v["params"], "spectral_stats": st["spectral_stats"]}
        y_eval, st_e = layer.apply(v2, x, training=False, mutable=["intermediates"])
        out.update({f"{tag}_y_eval": y_eval, f"{tag}_what_eval": st_e["intermediates"]["w"][0]})

    # ---- a12 mean_field_logits (utils.py:54-101)
    lg, var = rng.standard_normal((11, 5)) * 3, rng.uniform(0.0, 4.0, 11)
    out.update(mf_logits=lg, mf_var=var)
    for lik in ("logistic", "binary_logistic", "poisson"):
        out[f"mf_{lik}"] = utils.mean_field_logits(lg, var, mean_field_factor=0.7, likelihood=lik)
    out["mf_nocov"] = utils.mean_field_logits(lg, None, mean_field_factor=3.0)
    out["mf_negative"] = utils.mean_field_logits(lg, var, mean_field_factor=-1.0)

    # ---- §8f-2 MCSoftmaxDenseFA.__call__ (heteroscedastic_lib.py:43-336): Monte-Carlo softmax with factor-analysis noise.
    # The stand-in jax.random.normal is wrapped so that the draws the reference makes are stored next to its outputs.
    het = mods["heteroscedastic_lib"]
    import jax.random as jr

    real_normal = jr.normal
    for tag, share, temp in (("het", False, 1.0), ("het_shared", True, 0.6)):
        drawn = []

        def recording_normal(key, shape=(), dtype=None, _d=drawn):
            v = real_normal(key, shape, dtype)
            _d.append(np.array(v))
            return v

        jr.normal = jax.random.normal = recording_normal
        try:
            B, Din, K, Fn, S = 9, 7, 5, 3, 64
            layer = het.MCSoftmaxDenseFA(num_classes=K, num_factors=Fn, temperature=temp, train_mc_samples=S,
                                         test_mc_samples=S, share_samples_across_batch=share)
            x = rng.standard_normal((B, Din))
            rngs = {"params": jax.random.PRNGKey(6), "diag_noise_samples": jax.random.PRNGKey(7),
                    "standard_norm_noise_samples": jax.random.PRNGKey(8)}
            v = layer.init(rngs, x)
            for name in ("loc_layer", "scale_layer", "diag_layer"):  # non-trivial biases (init: zeros)
                v["params"][name]["bias"] = rng.standard_normal(v["params"][name]["bias"].shape) * 0.2
            del drawn[:]
            logits, log_probs, probs = layer.apply(v, x, training=True, rngs=rngs)
        finally:
            jr.normal = jax.random.normal = real_normal
        diag_n, fac_n = [d for d in drawn if d.ndim == 3 and d.shape[-1] == 1][-1], [d for d in drawn if d.ndim == 3 and d.shape[-1] == Fn][-1]
        out.update({f"{tag}_x": x, f"{tag}_logits": logits, f"{tag}_log_probs": log_probs, f"{tag}_probs": probs,
                    f"{tag}_diag_noise": diag_n, f"{tag}_factor_noise": fac_n, f"{tag}_temperature": np.float64(temp)})
        for name in ("loc_layer", "scale_layer", "diag_layer"):
            out[f"{tag}_{name}_kernel"] = v["params"][name]["kernel"]
            out[f"{tag}_{name}_bias"] = v["params"][name]["bias"]

    # ---- §8f-2 MCSigmoidDenseFA.__call__ (heteroscedastic_lib.py:338-594)
    drawn = []

    def recording_normal2(key, shape=(), dtype=None, _d=drawn):
        v = real_normal(key, shape, dtype)
        _d.append(np.array(v))
        return v

    jr.normal = jax.random.normal = recording_normal2
    try:
        B, Din, K, Fn, S = 8, 6, 4, 2, 48
        layer = het.MCSigmoidDenseFA(num_outputs=K, num_factors=Fn, temperature=0.8, train_mc_samples=S, test_mc_samples=S)
        x = rng.standard_normal((B, Din))
        rngs = {"params": jax.random.PRNGKey(9), "diag_noise_samples": jax.random.PRNGKey(10),
                "standard_norm_noise_samples": jax.random.PRNGKey(11)}
        v = layer.init(rngs, x)
        for name in ("loc_layer", "scale_layer", "diag_layer"):
            v["params"][name]["bias"] = rng.standard_normal(v["params"][name]["bias"].shape) * 0.2
        del drawn[:]
        logits, log_probs, probs = layer.apply(v, x, training=True, rngs=rngs)
    finally:
        jr.normal = jax.random.normal = real_normal
    out.update(hsig_x=x, hsig_logits=logits, hsig_log_probs=log_probs, hsig_probs=probs, hsig_temperature=np.float64(0.8),
               hsig_diag_noise=[d for d in drawn if d.ndim == 3 and d.shape[-1] == 1][-1],
               hsig_factor_noise=[d for d in drawn if d.ndim == 3 and d.shape[-1] == Fn][-1])
    for name in ("loc_layer", "scale_layer", "diag_layer"):
        out[f"hsig_{name}_kernel"] = v["params"][name]["kernel"]
        out[f"hsig_{name}_bias"] = v["params"][name]["bias"]

    # ---- §8f-2 MCSigmoidDenseFASNGP / MCSigmoidDenseFASNGPBE (random_feature.py:407-735): one training call (the
    # precision matrix is updated), then an evaluation call that mixes the GP variance into the diagonal scale
    for tag, cls, extra in (("hsngp", rf.MCSigmoidDenseFASNGP, {}), ("hsngpbe", rf.MCSigmoidDenseFASNGPBE, dict(ens_size=2))):
        drawn = []

        def recording_normal3(key, shape=(), dtype=None, _d=drawn):
            v = real_normal(key, shape, dtype)
            _d.append(np.array(v))
            return v

        jr.normal = jax.random.normal = recording_normal3
        try:
            B, Din, K, Fn, S, D = 8, 6, 4, 2, 48, 32
            layer = cls(num_outputs=K, num_factors=Fn, temperature=0.8, train_mc_samples=S, test_mc_samples=S,
                        hidden_features=D, het_var_weight=0.7, sngp_var_weight=1.3,
                        covmat_kwargs=dict(ridge_penalty=1.0), **extra)
            x, xt = rng.standard_normal((B, Din)), rng.standard_normal((B, Din))
            rngs = {"params": jax.random.PRNGKey(12), "diag_noise_samples": jax.random.PRNGKey(13),
                    "standard_norm_noise_samples": jax.random.PRNGKey(14)}
            v = layer.init(rngs, x)
            for name in ("scale_layer", "diag_layer"):
                v["params"][name]["bias"] = rng.standard_normal(v["params"][name]["bias"].shape) * 0.2
                if extra:  # BatchEnsemble fast weights away from their all-ones init
                    for fw in ("fast_weight_alpha", "fast_weight_gamma"):
                        v["params"][name][fw] = 1.0 + 0.3 * rng.standard_normal(v["params"][name][fw].shape)
            v["params"]["loc_layer"]["output_layer"]["bias"] = rng.standard_normal(K) * 0.2
            del drawn[:]
            (lg, cov, lp, pr), st = layer.apply(v, x, training=True, rngs=rngs, mutable=["laplace_covariance"])
            assert cov is None
            train_noise = list(drawn)
            v2 = {**{k: v[k] for k in v if k != "laplace_covariance"}, **st}
            del drawn[:]
            lg_e, cov_e, lp_e, pr_e = layer.apply(v2, xt, training=False, rngs=rngs)
            eval_noise = list(drawn)
        finally:
            jr.normal = jax.random.normal = real_normal
        pick = lambda ds, last: [d for d in ds if d.ndim == 3 and d.shape[-1] == last][-1]  # noqa: E731
        out.update({f"{tag}_x": x, f"{tag}_xt": xt, f"{tag}_logits": lg, f"{tag}_log_probs": lp, f"{tag}_probs": pr,
                    f"{tag}_diag_noise": pick(train_noise, 1), f"{tag}_factor_noise": pick(train_noise, Fn),
                    f"{tag}_eval_logits": lg_e, f"{tag}_eval_var": cov_e, f"{tag}_eval_log_probs": lp_e,
                    f"{tag}_eval_probs": pr_e, f"{tag}_eval_diag_noise": pick(eval_noise, 1),
                    f"{tag}_eval_factor_noise": pick(eval_noise, Fn),
                    f"{tag}_precision": st["laplace_covariance"]["loc_layer"]["covmat_layer"]["precision_matrix"],
                    f"{tag}_rff_kernel": v["random_features"]["loc_layer"]["hidden_layer"]["kernel"],
                    f"{tag}_rff_bias": v["random_features"]["loc_layer"]["hidden_layer"]["bias"],
                    f"{tag}_out_kernel": v["params"]["loc_layer"]["output_layer"]["kernel"],
                    f"{tag}_out_bias": v["params"]["loc_layer"]["output_layer"]["bias"]})
        for name in ("scale_layer", "diag_layer"):
            for leaf, val in v["params"][name].items():
                out[f"{tag}_{name}_{leaf}"] = val

    np.savez_compressed(OUT, **{k: np.asarray(a) for k, a in out.items()})
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
