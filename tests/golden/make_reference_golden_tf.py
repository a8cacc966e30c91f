"""Generates tests/golden/reference_vectors_tf.npz by EXECUTING THE UNMODIFIED REFERENCE SOURCE FILES

    /root/reference/edward2/tensorflow/layers/dense.py        (DenseReparameterization, DenseFlipout,
                                                                DenseBatchEnsemble, DenseRank1)
    /root/reference/edward2/tensorflow/layers/utils.py        (the add_weight decorator)
    /root/reference/edward2/tensorflow/{initializers,regularizers,constraints}.py
                                                               (TrainableNormal, NormalKLDivergence, Softplus,
                                                                OrthogonalRandomFeatures)
    /root/reference/edward2/tensorflow/layers/normalization.py (SpectralNormalization)
    /root/reference/edward2/tensorflow/layers/random_feature.py (RandomFeatureGaussianProcess,
                                                                LaplaceRandomFeatureCovariance)

under the NumPy-backed stand-ins of tensorflow / tensorflow_probability / ed.RandomVariable in tests/golden/refstub_tf/
(TensorFlow and TFP are not installable here; see refstub_tf/README.md for what the stand-ins restate).  Every statement
of the layers' build / call / regulariser code that runs is the reference's; the stand-ins evaluate it in float64 and
record every random draw (standard normals of Normal.sample, tf.random.uniform) next to the outputs.  The fixture pins
`oracle/dense.py` to the reference's TF classes (tests/test_oracle.py::test_reference_tf_*); the CUDA path is checked
against that oracle on injected noise (tests/test_gpu_dense_ops.py, tests/test_gpu_layers.py).

Run here (the GPU box has no /root/reference):  python tests/golden/make_reference_golden_tf.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/edward2/tensorflow"
OUT = os.path.join(HERE, "reference_vectors_tf.npz")


def load_reference():
    sys.path.insert(0, os.path.join(HERE, "refstub_tf"))
    import ed_random_variable as rv  # stand-in
    import tensorflow as tf  # noqa: F401  (stand-in)
    import tensorflow_probability as tfp  # noqa: F401  (stand-in)

    for name in ("edward2", "edward2.tensorflow", "edward2.tensorflow.layers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    pkg = sys.modules["edward2.tensorflow"]
    for name in ("random_variable", "generated_random_variables"):
        sys.modules[f"edward2.tensorflow.{name}"] = rv
        setattr(pkg, name, rv)
    sys.modules["edward2.tensorflow.transformed_random_variable"] = types.ModuleType("transformed_random_variable")
    pkg.transformed_random_variable = sys.modules["edward2.tensorflow.transformed_random_variable"]
    mods = {}
    for full, rel in (("edward2.tensorflow.constraints", "constraints.py"),
                      ("edward2.tensorflow.regularizers", "regularizers.py"),
                      ("edward2.tensorflow.initializers", "initializers.py"),
                      ("edward2.tensorflow.layers.utils", "layers/utils.py"),
                      ("edward2.tensorflow.layers.dense", "layers/dense.py"),
                      ("edward2.tensorflow.layers.normalization", "layers/normalization.py"),
                      ("edward2.tensorflow.layers.random_feature", "layers/random_feature.py"),
                      ("edward2.tensorflow.layers.convolutional", "layers/convolutional.py"),
                      ("edward2.tensorflow.layers.recurrent", "layers/recurrent.py"),
                      ("edward2.tensorflow.layers.heteroscedastic", "layers/heteroscedastic.py"),
                      ("edward2.tensorflow.layers.hetsngp", "layers/hetsngp.py")):
        spec = importlib.util.spec_from_file_location(full, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
        parent, leaf = full.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, m)
        mods[leaf] = m
    return mods


def main():
    mods = load_reference()
    import tensorflow as tf
    import tensorflow_probability as tfp

    dense = mods["dense"]
    rng = np.random.default_rng(20261021)
    tf.random._rng = np.random.default_rng(7)
    out = {}

    uniforms = []
    real_uniform = tf.random.uniform

    def recording_uniform(*args, **kwargs):
        v = real_uniform(*args, **kwargs)
        uniforms.append(np.array(v, dtype=np.float64))
        return v

    tf.random.uniform = recording_uniform

    def set_normal(init, mean, rho):
        """Overwrite the variables of a built TrainableNormal initializer (mean, pre-softplus stddev)."""
        init.mean, init.stddev = tf._t(mean), tf._t(rho)

    # ---- a2 DenseReparameterization (dense.py:33-83) + a5-a7: add_weight, TrainableNormal, Softplus, NormalKLDivergence
    B, I, O = 6, 10, 7
    layer = dense.DenseReparameterization(O, activation="relu", bias_initializer="zeros")
    x = rng.standard_normal((B, I))
    layer(x)  # builds: kernel is an ed.RandomVariable made by the trainable initializer
    mu, rho = rng.standard_normal((I, O)) * 0.3, rng.standard_normal((I, O)) * 0.5 - 2.0
    set_normal(layer.kernel_initializer, mu, rho)
    layer.bias = tf._t(rng.standard_normal(O) * 0.2)
    del tfp.recorded[:]
    y = layer(x)
    eps = tfp.recorded[-1]  # the draw behind the kernel that call() used (call_weights re-samples, dense.py:75-78)
    kl = layer.losses
    assert len(kl) == 1
    out.update(rp_x=x, rp_mu=mu, rho_rp=rho, rp_eps=eps, rp_bias=np.asarray(layer.bias), rp_y=y, rp_kl=kl[0],
               rp_kernel=np.asarray(layer.kernel.value), rp_stddev=np.asarray(layer.kernel.distribution.stddev()))
    # the regulariser with a non-default prior and scale_factor (regularizers.py:166-186)
    reg = mods["regularizers"].NormalKLDivergence(mean=0.3, stddev=0.7, scale_factor=1.0 / 50000.0)
    out.update(rp_kl_prior=reg(layer.kernel), rp_kl_prior_args=np.array([0.3, 0.7, 1.0 / 50000.0]))
    # 3-D inputs go through Keras Dense's tensordot branch
    x3 = rng.standard_normal((2, 3, I))
    del tfp.recorded[:]
    y3 = layer(x3)
    out.update(rp_x3=x3, rp_eps3=tfp.recorded[-1], rp_y3=y3)

    # ---- a3 DenseFlipout (dense.py:227-285): int32 sign_input, FLOAT sign_output (Uniform[-1, 3) for float inputs)
    layer = dense.DenseFlipout(O, activation="relu")
    x = tf._t(rng.standard_normal((B, I)))
    layer(x)
    mu, rho = rng.standard_normal((I, O)) * 0.3, rng.standard_normal((I, O)) * 0.5 - 2.0
    set_normal(layer.kernel_initializer, mu, rho)
    layer.bias = tf._t(rng.standard_normal(O) * 0.2)
    del tfp.recorded[:], uniforms[:]
    y = layer(x)
    assert len(uniforms) == 2
    out.update(fo_x=np.asarray(x), fo_mu=mu, fo_rho=rho, fo_eps=tfp.recorded[-1], fo_bias=np.asarray(layer.bias),
               fo_sign_in=2 * uniforms[0] - 1, fo_sign_out=2 * uniforms[1] - 1, fo_y=y)
    x3 = tf._t(rng.standard_normal((2, 3, I)))
    del tfp.recorded[:], uniforms[:]
    y3 = layer(x3)
    out.update(fo_x3=np.asarray(x3), fo_eps3=tfp.recorded[-1], fo_sign_in3=2 * uniforms[0] - 1,
               fo_sign_out3=2 * uniforms[1] - 1, fo_y3=y3)

    # ---- a1 DenseBatchEnsemble, TF class (dense.py:470-607): rank 1 and rank 3
    E, n = 3, 4
    for tag, rank in (("be", 1), ("be_r3", 3)):
        layer = dense.DenseBatchEnsemble(O, rank=rank, ensemble_size=E, activation="relu")
        x = rng.standard_normal((E * n, I))
        layer(x)
        lead = (rank,) if rank > 1 else ()
        layer.alpha = tf._t(1.0 + 0.4 * rng.standard_normal(lead + (E, I)))
        layer.gamma = tf._t(1.0 + 0.4 * rng.standard_normal(lead + (E, O)))
        layer.ensemble_bias = tf._t(0.2 * rng.standard_normal((E, O)))
        y = layer(x)
        out.update({f"{tag}_x": x, f"{tag}_kernel": np.asarray(layer.kernel), f"{tag}_alpha": np.asarray(layer.alpha),
                    f"{tag}_gamma": np.asarray(layer.gamma), f"{tag}_bias": np.asarray(layer.ensemble_bias), f"{tag}_y": y})

    # ---- a4 DenseRank1 (dense.py:991-1175): per-example sampled fast weights [n, E, dim] -> transposed to [E, n, dim],
    # clip bounds, multiplicative / additive perturbations, deterministic / sampled ensemble bias, KL of the fast weights
    for tag, additive, sampled_bias, lo, hi in (("r1", False, False, 0.6, 1.5), ("r1_add", True, False, -10, 10),
                                                ("r1_bias", False, True, -10, 10)):
        layer = dense.DenseRank1(O, activation="relu", ensemble_size=E, use_additive_perturbation=additive,
                                 min_perturbation_value=lo, max_perturbation_value=hi,
                                 bias_initializer="trainable_normal" if sampled_bias else "zeros")
        x = rng.standard_normal((E * n, I))
        layer(x)
        mu_r, rho_r = 1.0 + 0.3 * rng.standard_normal((E, I)), rng.standard_normal((E, I)) * 0.3 - 1.5
        mu_s, rho_s = 1.0 + 0.3 * rng.standard_normal((E, O)), rng.standard_normal((E, O)) * 0.3 - 1.5
        set_normal(layer.alpha_initializer, mu_r, rho_r)
        set_normal(layer.gamma_initializer, mu_s, rho_s)
        if sampled_bias:
            mu_b, rho_b = 0.2 * rng.standard_normal((E, O)), rng.standard_normal((E, O)) * 0.3 - 2.0
            set_normal(layer.ensemble_bias_initializer, mu_b, rho_b)
            out.update({f"{tag}_rho_b": rho_b})
        else:
            mu_b = 0.2 * rng.standard_normal((E, O))
            layer.ensemble_bias = tf._t(mu_b)
        del tfp.recorded[:]
        y = layer(x)
        # Normal.sample(examples_per_model) draws are the [n, E, dim] ones; each initializer call also makes one [E, dim] draw
        per_example = [r for r in tfp.recorded if r.ndim == 3]
        assert len(per_example) == (3 if sampled_bias else 2)
        # reference layout eps[i, e, :] -> member-major rows b = e * n + i
        to_rows = lambda a: np.transpose(a, (1, 0, 2)).reshape(E * n, -1)  # noqa: E731
        out.update({f"{tag}_x": x, f"{tag}_kernel": np.asarray(layer.kernel), f"{tag}_mu_r": mu_r, f"{tag}_rho_r": rho_r,
                    f"{tag}_mu_s": mu_s, f"{tag}_rho_s": rho_s, f"{tag}_bias": mu_b, f"{tag}_eps_r": to_rows(per_example[0]),
                    f"{tag}_eps_s": to_rows(per_example[1]), f"{tag}_y": y, f"{tag}_clip": np.array([lo, hi], dtype=np.float64),
                    f"{tag}_kl": np.array([float(np.asarray(v)) for v in layer.losses])})
        if sampled_bias:
            out[f"{tag}_eps_b"] = to_rows(per_example[2])

    # ---- a9 SpectralNormalization, TF class (normalization.py:284-395): power iteration, sigma, IN-PLACE kernel update
    # (self.w aliases self.layer.kernel, so restore_weights is a no-op and the kernel stays normalised)
    norm = mods["normalization"]
    for tag, iters, c, wscale in (("sn1", 1, 0.95, 1.0), ("sn3", 3, 0.6, 1.0), ("sn_small", 2, 0.95, 0.02)):
        inner = tf.keras.layers.Dense(14)
        sn = norm.SpectralNormalization(inner, iteration=iters, norm_multiplier=c)
        x = rng.standard_normal((9, 10))
        sn(x)  # builds (and already runs update_weights once, normalization.py:352)
        w0 = rng.standard_normal((10, 14)) * wscale
        u0, v0 = rng.standard_normal((1, 14)), rng.standard_normal((1, 10))
        inner.kernel.assign(w0); sn.u.assign(u0); sn.v.assign(v0)  # noqa: E702
        inner.bias.assign(rng.standard_normal(14) * 0.1)
        y = sn(x, training=True)
        out.update({f"{tag}_x": x, f"{tag}_w0": w0, f"{tag}_u0": u0, f"{tag}_v0": v0, f"{tag}_b": np.array(inner.bias),
                    f"{tag}_iters": iters, f"{tag}_c": c, f"{tag}_y": y, f"{tag}_u1": np.array(sn.u), f"{tag}_v1": np.array(sn.v),
                    f"{tag}_w1": np.array(inner.kernel)})
        y_eval = sn(x, training=False)  # no power iteration: sigma from the stored u, v and the already normalised kernel
        out.update({f"{tag}_y_eval": y_eval, f"{tag}_w2": np.array(inner.kernel)})

    # ---- a11 LaplaceRandomFeatureCovariance, TF class (random_feature.py:284-549): momentum / exact updates, likelihoods,
    # the covariance inverse on the first evaluation call, the update flag
    rfm = mods["random_feature"]
    D = 24
    phi1, phi2, phit = rng.standard_normal((16, D)) * 0.3, rng.standard_normal((16, D)) * 0.3, rng.standard_normal((9, D)) * 0.3
    lg1, lg2 = rng.standard_normal((16, 1)), rng.standard_normal((16, 1))
    out.update(cov_phi1=phi1, cov_phi2=phi2, cov_phit=phit, cov_logits1=lg1, cov_logits2=lg2)
    for tag, momentum, likelihood, ridge in (("cov_exact", -1.0, "gaussian", 1.0), ("cov_mom", 0.9, "gaussian", 0.5),
                                             ("cov_logistic", -1.0, "binary_logistic", 1.0),
                                             ("cov_poisson", 0.99, "poisson", 2.0)):
        layer = rfm.LaplaceRandomFeatureCovariance(momentum=momentum, ridge_penalty=ridge, likelihood=likelihood)
        r1 = layer(tf._t(phi1), tf._t(lg1), training=True)
        assert np.array_equal(np.asarray(r1), np.eye(16))  # random_feature.py:531: identity while training
        p1 = np.array(layer.precision_matrix)
        layer(tf._t(phi2), tf._t(lg2), training=True)
        p2 = np.array(layer.precision_matrix)
        assert bool(np.asarray(layer.if_update_covariance))
        cov = layer(tf._t(phit), None, training=False)
        assert not bool(np.asarray(layer.if_update_covariance))
        out.update({f"{tag}_p1": p1, f"{tag}_p2": p2, f"{tag}_sigma": np.array(layer.covariance_matrix), f"{tag}_cov": cov,
                    f"{tag}_args": np.array([momentum, ridge])})

    # ---- a10 RandomFeatureGaussianProcess, TF class (random_feature.py:33-281) end to end: LayerNormalization ->
    # Dense(cos) with OrthogonalRandomFeatures / Uniform(0, 2 pi) bias -> * sqrt(2 / D) -> output weights + bias -> covariance
    gp = rfm.RandomFeatureGaussianProcess(units=3, num_inducing=32, gp_cov_momentum=-1, gp_cov_ridge_penalty=1.0,
                                          return_random_features=True)
    xa, xt = rng.standard_normal((20, 12)), rng.standard_normal((7, 12))
    la, cova, fa = gp(tf._t(xa), training=True)
    gp._gp_output_bias.assign(rng.standard_normal(3) * 0.1)
    la, cova, fa = gp(tf._t(xa), training=True)   # precision now holds two exact-sum updates of the same batch
    lt, covt, ft = gp(tf._t(xt), training=False)
    rf_w = np.array(gp._random_feature.kernel)
    out.update(gp_xa=xa, gp_xt=xt, gp_rff_kernel=rf_w, gp_rff_bias=np.array(gp._random_feature.bias),
               gp_out_kernel=np.array(gp._gp_output_layer.kernel), gp_out_bias=np.array(gp._gp_output_bias),
               gp_ln_gamma=np.array(gp._input_norm_layer.gamma), gp_ln_beta=np.array(gp._input_norm_layer.beta),
               gp_logits_a=la, gp_features_a=fa, gp_precision=np.array(gp._gp_cov_layer.precision_matrix),
               gp_logits=lt, gp_cov=covt, gp_features=ft)
    # OrthogonalRandomFeatures (initializers.py:759-832): orthogonal directions, rows < cols -> concatenated square blocks
    out["orf_gram"] = rf_w[:, :12].T @ rf_w[:, :12]

    # ---- a12 mean_field_logits, TF function (layers/utils.py:379-424)
    utils = mods["utils"]
    lg, covm = rng.standard_normal((11, 5)) * 3, np.diag(rng.uniform(0.0, 4.0, 11)) + 0.01
    out.update(mf_logits=lg, mf_cov=covm)
    for lik in ("logistic", "binary_logistic", "poisson"):
        out[f"mf_{lik}"] = utils.mean_field_logits(tf._t(lg), tf._t(covm), mean_field_factor=0.7, likelihood=lik)
    out["mf_nocov"] = utils.mean_field_logits(tf._t(lg), None, mean_field_factor=3.0)
    out["mf_negative"] = utils.mean_field_logits(tf._t(lg), tf._t(covm), mean_field_factor=-1.0)

    # ---- §8f-1 Conv2D variants (layers/convolutional.py:32,167,560,2056), NHWC
    conv = mods["convolutional"]
    Bc, Hc, Wc, Ci, Co = 4, 6, 7, 3, 5
    # Conv2DReparameterization (:88-98): 'same' padding, stride 1
    layer = conv.Conv2DReparameterization(Co, 3, padding="same", activation="relu")
    x = rng.standard_normal((Bc, Hc, Wc, Ci))
    layer(tf._t(x))
    mu, rho = rng.standard_normal((3, 3, Ci, Co)) * 0.3, rng.standard_normal((3, 3, Ci, Co)) * 0.5 - 2.0
    set_normal(layer.kernel_initializer, mu, rho)
    layer.bias = tf._t(rng.standard_normal(Co) * 0.2)
    del tfp.recorded[:]
    y = layer(tf._t(x))
    out.update(crp_x=x, crp_mu=mu, crp_rho=rho, crp_eps=tfp.recorded[-1], crp_bias=np.asarray(layer.bias), crp_y=y,
               crp_kl=layer.losses[0])
    # Conv2DFlipout (:225-249): 'valid' padding, stride 2; both sign tensors int32 of shape [B, 1, 1, C]
    layer = conv.Conv2DFlipout(Co, 3, strides=2, padding="valid", activation="relu")
    layer(tf._t(x))
    mu, rho = rng.standard_normal((3, 3, Ci, Co)) * 0.3, rng.standard_normal((3, 3, Ci, Co)) * 0.5 - 2.0
    set_normal(layer.kernel_initializer, mu, rho)
    layer.bias = tf._t(rng.standard_normal(Co) * 0.2)
    del tfp.recorded[:], uniforms[:]
    y = layer(tf._t(x))
    assert len(uniforms) == 2 and uniforms[0].shape == (Bc, 1, 1, Ci) and uniforms[1].shape == (Bc, 1, 1, Co)
    out.update(cfo_x=x, cfo_mu=mu, cfo_rho=rho, cfo_eps=tfp.recorded[-1], cfo_bias=np.asarray(layer.bias),
               cfo_sign_in=(2 * uniforms[0] - 1).reshape(Bc, Ci), cfo_sign_out=(2 * uniforms[1] - 1).reshape(Bc, Co), cfo_y=y)
    # Conv2DBatchEnsemble (:681-699), rank 1, two members
    Ec, nc = 2, 2
    layer = conv.Conv2DBatchEnsemble(Co, 3, ensemble_size=Ec, padding="same", activation="relu")
    layer(tf._t(x))
    layer.alpha = tf._t(1.0 + 0.4 * rng.standard_normal((Ec, Ci)))
    layer.gamma = tf._t(1.0 + 0.4 * rng.standard_normal((Ec, Co)))
    layer.ensemble_bias = tf._t(0.2 * rng.standard_normal((Ec, Co)))
    y = layer(tf._t(x))
    out.update(cbe_x=x, cbe_kernel=np.asarray(layer.kernel), cbe_alpha=np.asarray(layer.alpha),
               cbe_gamma=np.asarray(layer.gamma), cbe_bias=np.asarray(layer.ensemble_bias), cbe_y=y)
    # Conv2DRank1 (:1968-2020): per-example fast weights [n, E, C] -> [E, n, C] -> rows, broadcast over the spatial axes
    layer = conv.Conv2DRank1(Co, 3, ensemble_size=Ec, padding="same", strides=(2, 1), activation="relu")
    layer(tf._t(x))
    mu_r, rho_r = 1.0 + 0.3 * rng.standard_normal((Ec, Ci)), rng.standard_normal((Ec, Ci)) * 0.3 - 1.5
    mu_s, rho_s = 1.0 + 0.3 * rng.standard_normal((Ec, Co)), rng.standard_normal((Ec, Co)) * 0.3 - 1.5
    set_normal(layer.alpha_initializer, mu_r, rho_r)
    set_normal(layer.gamma_initializer, mu_s, rho_s)
    mu_b = 0.2 * rng.standard_normal((Ec, Co))
    layer.ensemble_bias = tf._t(mu_b)
    del tfp.recorded[:]
    y = layer(tf._t(x))
    per_example = [r_ for r_ in tfp.recorded if r_.ndim == 3]
    assert len(per_example) == 2
    rows = lambda a: np.transpose(a, (1, 0, 2)).reshape(Ec * nc, -1)  # noqa: E731
    out.update(cr1_x=x, cr1_kernel=np.asarray(layer.kernel), cr1_mu_r=mu_r, cr1_rho_r=rho_r, cr1_mu_s=mu_s, cr1_rho_s=rho_s,
               cr1_bias=mu_b, cr1_eps_r=rows(per_example[0]), cr1_eps_s=rows(per_example[1]), cr1_y=y)

    # ---- §8f-3 LSTM cells (layers/recurrent.py:30,185,360): weights / signs / fast weights drawn once per sequence by
    # get_initial_state, reused by every step
    rec = mods["recurrent"]
    Hh, Ii, T = 5, 4, 3

    def unroll(cell, xs, states):
        hs = []
        for t in range(xs.shape[0]):
            h, states = cell(tf._t(xs[t]), states)
            hs.append(np.array(h))
        return np.stack(hs), np.array(states[1])

    for tag, cls, impl in (("lrp", rec.LSTMCellReparameterization, 2), ("lrp1", rec.LSTMCellReparameterization, 1),
                           ("lfo", rec.LSTMCellFlipout, 2), ("lfo1", rec.LSTMCellFlipout, 1)):
        Bl = 3
        cell = cls(Hh, implementation=impl)
        xs = rng.standard_normal((T, Bl, Ii))
        cell(tf._t(xs[0]), cell.get_initial_state(batch_size=Bl, dtype=tf.float32))  # builds
        mu_k, rho_k = rng.standard_normal((Ii, 4 * Hh)) * 0.4, rng.standard_normal((Ii, 4 * Hh)) * 0.5 - 2.0
        mu_u, rho_u = rng.standard_normal((Hh, 4 * Hh)) * 0.4, rng.standard_normal((Hh, 4 * Hh)) * 0.5 - 2.0
        set_normal(cell.kernel_initializer, mu_k, rho_k)
        set_normal(cell.recurrent_initializer, mu_u, rho_u)
        cell.bias = tf._t(rng.standard_normal(4 * Hh) * 0.3)
        del tfp.recorded[:], uniforms[:]
        states = cell.get_initial_state(batch_size=Bl, dtype=tf.float32)  # samples the weights (and the Flipout signs)
        draws = [d for d in tfp.recorded]
        assert [d.shape for d in draws] == [(Ii, 4 * Hh)] * 2 + [(Hh, 4 * Hh)] * 2
        hs, c_last = unroll(cell, xs, states)
        assert len(tfp.recorded) == 4  # nothing is re-drawn by the steps
        out.update({f"{tag}_xs": xs, f"{tag}_mu_k": mu_k, f"{tag}_rho_k": rho_k, f"{tag}_mu_u": mu_u, f"{tag}_rho_u": rho_u,
                    f"{tag}_bias": np.array(cell.bias), f"{tag}_eps_k": draws[1], f"{tag}_eps_u": draws[3], f"{tag}_hs": hs,
                    f"{tag}_c": c_last, f"{tag}_kl": np.array([float(np.asarray(v)) for v in cell.losses])})
        if cls is rec.LSTMCellFlipout:
            assert [u.shape for u in uniforms] == [(Bl, Ii), (Bl, 4 * Hh), (Bl, Hh), (Bl, 4 * Hh)]
            out.update({f"{tag}_sign_in": 2 * uniforms[0] - 1, f"{tag}_sign_out": 2 * uniforms[1] - 1,
                        f"{tag}_rsign_in": 2 * uniforms[2] - 1, f"{tag}_rsign_out": 2 * uniforms[3] - 1})

    El, nl = 2, 3
    for tag, additive in (("lr1", False), ("lr1_add", True)):
        cell = rec.LSTMCellRank1(Hh, ensemble_size=El, use_additive_perturbation=additive)
        xs = rng.standard_normal((T, El * nl, Ii))
        cell(tf._t(xs[0]), cell.get_initial_state(batch_size=El * nl, dtype=tf.float32))  # builds
        params = {}
        for name, dim in (("alpha", Ii), ("gamma", 4 * Hh), ("recurrent_alpha", Hh), ("recurrent_gamma", 4 * Hh)):
            mu, rho = 1.0 + 0.3 * rng.standard_normal((El, dim)), rng.standard_normal((El, dim)) * 0.3 - 1.5
            set_normal(getattr(cell, name + "_initializer"), mu, rho)
            params[name] = (mu, rho)
        cell.kernel = tf._t(rng.standard_normal((Ii, 4 * Hh)) * 0.4)
        cell.recurrent_kernel = tf._t(rng.standard_normal((Hh, 4 * Hh)) * 0.4)
        cell.bias = tf._t(rng.standard_normal((El, 4 * Hh)) * 0.3)
        del tfp.recorded[:]
        states = cell.get_initial_state(batch_size=El * nl, dtype=tf.float32)
        per_example = [d for d in tfp.recorded if d.ndim == 3]
        assert len(per_example) == 4
        hs, c_last = unroll(cell, xs, states)
        rows = lambda a: np.transpose(a, (1, 0, 2)).reshape(El * nl, -1)  # noqa: E731
        out.update({f"{tag}_xs": xs, f"{tag}_kernel": np.array(cell.kernel), f"{tag}_rkernel": np.array(cell.recurrent_kernel),
                    f"{tag}_bias": np.array(cell.bias), f"{tag}_hs": hs, f"{tag}_c": c_last})
        for (name, (mu, rho)), eps in zip(params.items(), per_example):
            out.update({f"{tag}_{name}_mu": mu, f"{tag}_{name}_rho": rho, f"{tag}_{name}_eps": rows(eps)})

    # ---- §8f-2 heteroscedastic heads, TF classes (layers/heteroscedastic.py:508 MCSoftmaxDenseFA, :1382 MCSigmoidDenseFA;
    # layers/hetsngp.py:25 HeteroscedasticSNGPLayer): ONE diagonal draw per CLASS (the JAX classes draw one per sample)
    het, hsn = mods["heteroscedastic"], mods["hetsngp"]
    Bh, Dh, Kh, Fh, Sh = 6, 7, 5, 3, 24

    def fill_dense(layer_, scale=0.5):
        layer_.kernel.assign(rng.standard_normal(layer_.kernel.shape) * scale)
        layer_.bias.assign(rng.standard_normal(layer_.bias.shape) * 0.2)

    def head_params(layer_, tag):
        for name in ("_loc_layer", "_scale_layer", "_diag_layer"):
            sub = getattr(layer_, name)
            out[f"{tag}{name}_kernel"], out[f"{tag}{name}_bias"] = np.array(sub.kernel), np.array(sub.bias)

    def mc_draws():
        """(diag [B or 1, S, K], factors [B or 1, S, F]) of the last predictive-mean pass, as [rows, S, dim]."""
        d = [a for a in tfp.recorded if a.ndim == 3]
        return np.transpose(d[0], (1, 0, 2)), np.transpose(d[1], (1, 0, 2))

    for tag, cls, share in (("hsm", het.MCSoftmaxDenseFA, False), ("hsm_shared", het.MCSoftmaxDenseFA, True),
                            ("hsg", het.MCSigmoidDenseFA, False)):
        kw = dict(num_factors=Fh, temperature=0.7, train_mc_samples=Sh, test_mc_samples=Sh, compute_pred_variance=True,
                  share_samples_across_batch=share)
        layer = cls(Kh, **kw)
        x = rng.standard_normal((Bh, Dh))
        layer(tf._t(x), training=True, seed=1)
        for name in ("_loc_layer", "_scale_layer", "_diag_layer"):
            fill_dense(getattr(layer, name))
        del tfp.recorded[:]
        logits, log_probs, probs, var = layer(tf._t(x), training=True, seed=11)
        e, z = mc_draws()
        # the variance pass re-draws with the same seed: identical samples (TF semantics, see the tfp stand-in)
        d3 = [a for a in tfp.recorded if a.ndim == 3]
        assert len(d3) == 4 and np.array_equal(d3[0], d3[2]) and np.array_equal(d3[1], d3[3])
        head_params(layer, tag)
        out.update({f"{tag}_x": x, f"{tag}_logits": logits, f"{tag}_log_probs": log_probs, f"{tag}_probs": probs,
                    f"{tag}_var": var, f"{tag}_diag_noise": e, f"{tag}_factor_noise": z})

    layer = hsn.HeteroscedasticSNGPLayer(Kh, num_factors=Fh, temperature=0.8, train_mc_samples=Sh, test_mc_samples=Sh,
                                         logits_only=False, num_inducing=16, gp_cov_momentum=-1.0, gp_cov_ridge_penalty=1.0,
                                         normalize_input=True, sngp_var_weight=1.3, het_var_weight=0.7)
    x, xt = rng.standard_normal((Bh, Dh)), rng.standard_normal((Bh, Dh))
    # hetsngp.py:160 calls the SNGP layer without `training`: it follows tf.keras.backend.learning_phase()
    tf.keras.backend.learning_phase = lambda: 1
    layer(tf._t(x), training=True, seed=2)  # builds; first exact-sum precision update
    for name in ("_scale_layer", "_diag_layer"):
        fill_dense(getattr(layer, name))
    sn = layer.sngp_layer
    sn._gp_output_layer.kernel.assign(rng.standard_normal(sn._gp_output_layer.kernel.shape) * 0.5)
    sn._gp_output_bias.assign(rng.standard_normal(Kh) * 0.2)
    sn._gp_cov_layer.precision_matrix.assign(np.zeros((16, 16)))
    del tfp.recorded[:]
    logits, log_probs, probs, _ = layer(tf._t(x), training=True, seed=12)
    e, z = mc_draws()
    out.update(hsn_x=x, hsn_xt=xt, hsn_logits=logits, hsn_log_probs=log_probs, hsn_probs=probs, hsn_diag_noise=e,
               hsn_factor_noise=z, hsn_precision=np.array(sn._gp_cov_layer.precision_matrix),
               hsn_rff_kernel=np.array(sn._random_feature.kernel), hsn_rff_bias=np.array(sn._random_feature.bias),
               hsn_out_kernel=np.array(sn._gp_output_layer.kernel), hsn_out_bias=np.array(sn._gp_output_bias),
               hsn_ln_gamma=np.array(sn._input_norm_layer.gamma), hsn_ln_beta=np.array(sn._input_norm_layer.beta),
               hsn_weights=np.array([0.7, 1.3, 0.8]))
    for name in ("_scale_layer", "_diag_layer"):
        sub = getattr(layer, name)
        out[f"hsn{name}_kernel"], out[f"hsn{name}_bias"] = np.array(sub.kernel), np.array(sub.bias)
    tf.keras.backend.learning_phase = lambda: 0
    del tfp.recorded[:]
    logits, log_probs, probs, _ = layer(tf._t(xt), training=False, seed=13)
    e, z = mc_draws()
    out.update(hsn_eval_logits=logits, hsn_eval_log_probs=log_probs, hsn_eval_probs=probs, hsn_eval_diag_noise=e,
               hsn_eval_factor_noise=z, hsn_eval_sigma=np.array(sn._gp_cov_layer.covariance_matrix))

    np.savez_compressed(OUT, **{k: np.asarray(a, dtype=np.float64) for k, a in out.items()})
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
