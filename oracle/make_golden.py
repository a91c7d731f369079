"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the LIVE reference.

Run in the build container (where /root/reference exists):

    PYTHONHASHSEED=0 python oracle/make_golden.py

For every case the candidate/evaluated trees are produced by the reference's own generators
(``KernelGrammarCandidateGenerator``), the dense feature matrices by the reference's own
``internal_transform_X`` / ``create_hash_array_mapping`` / ``feature_matrix``
(bosot/kernels/kernel_kernel_grammar_tree.py:164-230), both unmodified and imported through
``oracle/ref_shim.py``.  Trees are stored as canonical name strings (``str(expr)``); raw Python
hash keys are never stored (they are salted per process).  Distances / Gram / posterior / EI are
computed from those reference feature matrices with the restated TF/GPflow arithmetic of
``oracle/ref_numpy.py`` (TensorFlow and GPflow cannot be installed here).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_numpy as rn  # noqa: E402
from oracle import ref_shim  # noqa: E402

CASES = [
    # name, search space class, input dim, E, N, hyper
    ("c1_toy_ckswithrq_d2", "CKSWithRQSearchSpace", 2, 24, 100, "H0"),
    ("c2_airline_ckswithrq_d1", "CKSWithRQSearchSpace", 1, 12, 100, "H1"),
    ("c3_powerplant_ckshighdim_d4", "CKSHighDimSearchSpace", 4, 24, 100, "H0"),
    ("c4_airfoil_ckshighdim_d5", "CKSHighDimSearchSpace", 5, 50, 160, "H1"),
    ("cks_d3", "CompositionalKernelSearchSpace", 3, 18, 60, "H1"),
    ("ndimfull_d2", "NDimFullKernelsSearchSpace", 2, 12, 60, "H0"),
    # big trees: long random walks (up to ~20 leaves) and C5-style random depth<=5 trees (up to 32 leaves)
    ("walks_ckswithrq_d2", "CKSWithRQSearchSpace", 2, 16, 64, "H1"),
    ("c5_stress_ckshighdim_d5", "CKSHighDimSearchSpace", 5, 40, 120, "H0"),
]


def main() -> None:
    ref_shim.install()
    import bosot.kernels.kernel_grammar.kernel_grammar_search_spaces as kgss
    from bosot.configs.kernels.grammar_tree_kernel_kernel_configs import OTWeightedDimsExtendedGrammarKernelConfig
    from bosot.kernels.kernel_grammar.kernel_grammar_candidate_generator import KernelGrammarCandidateGenerator
    from bosot.kernels.kernel_kernel_grammar_tree import FeatureType, OptimalTransportKernelKernel

    ft_map = {
        rn.DIM_WISE: FeatureType.DIM_WISE_WEIGHTED_ELEMENTARY_COUNT,
        rn.PATHS: FeatureType.REDUCED_ELEMENTARY_PATHS,
        rn.SUBTREES: FeatureType.SUBTREES,
    }
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, space, d, E, N, hyp in CASES:
        ss = getattr(kgss, space)(d)
        gen = KernelGrammarCandidateGenerator(ss, 5, 15, 1.0 / 3.0, 50, 5, False)
        B = ss.get_num_base_kernels()
        if name.startswith("walks"):
            np.random.seed(1)
            roots = ss.get_root_expressions()
            x_eval = [ss.random_walk(24, roots[i % B])[-1] for i in range(E)]
            np.random.seed(3)
            x_cand = [ss.random_walk(6 + (i % 28), roots[i % B])[-1] for i in range(N)]
        elif name.startswith("c5_stress"):
            from bosot.kernels.kernel_factory import KernelFactory
            from bosot.kernels.kernel_grammar.kernel_grammar import (
                ElementaryKernelGrammarExpression,
                KernelGrammarExpression,
                KernelGrammarOperator,
            )

            names = [str(r) for r in ss.get_root_expressions()]

            def build(t):
                if isinstance(t, str):
                    e = ElementaryKernelGrammarExpression(KernelFactory.build(ss.base_kernel_config_list[names.index(t)]))
                else:
                    op = KernelGrammarOperator.ADD if t[0] == "ADD" else KernelGrammarOperator.MULTIPLY
                    e = KernelGrammarExpression(build(t[1]), build(t[2]), op)
                e.set_generator_name(ss.name)
                return e

            rng_t = np.random.default_rng(1234)
            x_eval = [build(rn.random_stress_tree(rng_t, names)) for _ in range(E)]
            x_cand = [build(rn.random_stress_tree(rng_t, names)) for _ in range(N)]
        else:
            np.random.seed(1)
            x_eval = gen.get_random_canditates(max(B, (E // B) * B))  # KGCG:48-60: a multiple of B
            np.random.seed(2)
            extra = gen.get_dataset_recursivly_generated(E + B, 1)
            x_eval = (x_eval + extra[B:])[:E]
            np.random.seed(3)
            x_cand = gen.get_dataset_recursivly_generated(N, 1)
        # exact copies of evaluated trees among the candidates (distance 0 / variance ~ noise)
        x_cand[5] = x_eval[-1]
        x_cand[6] = x_eval[len(x_eval) // 2]
        kern = OptimalTransportKernelKernel(**OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=d).dict())
        assert [ft_map[f] for f in rn.DEFAULT_FEATURE_TYPES] == list(kern.feature_type_list)

        arrays = {}
        D_cross, D_self = [], []
        for f in rn.DEFAULT_FEATURE_TYPES:
            de = kern.internal_transform_X(x_eval, ft_map[f])
            dc = kern.internal_transform_X(x_cand, ft_map[f])
            norm = kern.normalize_features(ft_map[f])
            # K(X_eval, X_cand): index seeded by the evaluated trees first (gpflow asks kernel(X_train, X_new))
            idx = kern.create_hash_array_mapping(de + dc)
            FE = kern.feature_matrix(de, idx, norm)
            FC = kern.feature_matrix(dc, idx, norm)
            idx_e = kern.create_hash_array_mapping(de)
            FEE = kern.feature_matrix(de, idx_e, norm)
            arrays["F_eval_joint_" + f] = FE
            arrays["F_cand_joint_" + f] = FC
            arrays["F_eval_self_" + f] = FEE
            D_cross.append(rn.manhatten_distance(FE, FC))
            D_self.append(rn.manhatten_distance(FEE, FEE))
        hyper = rn.Hyper(**getattr(rn, hyp))
        w = (hyper.alphas / np.sum(hyper.alphas))[:, None, None]
        D_cross, D_self = np.stack(D_cross), np.stack(D_self)
        d_cross = np.sum(w * D_cross, axis=0)
        d_self = np.sum(w * D_self, axis=0)
        K_EN = hyper.variance * np.exp(-1 * d_cross / np.power(hyper.lengthscale, 2.0))
        K_EE = hyper.variance * np.exp(-1 * d_self / np.power(hyper.lengthscale, 2.0))
        rng = np.random.default_rng(4)
        y = rng.standard_normal((len(x_eval), 1))
        mu_e, sig_e = rn.gp_posterior_from_grams(K_EE, K_EE, np.full(len(x_eval), hyper.variance), y, hyper)
        cur = np.max(mu_e)
        mu, sig = rn.gp_posterior_from_grams(K_EE, K_EN, np.full(len(x_cand), hyper.variance), y, hyper)
        ei = rn.expected_improvement(mu, sig, cur)
        ucb = mu + np.sqrt(3.0) * sig
        arrays.update(
            eval_names=np.array([str(x) for x in x_eval]),
            cand_names=np.array([str(x) for x in x_cand]),
            generator_name=np.array(ss.name),
            input_dimension=np.array(d),
            base_kernel_names=np.array([str(r) for r in ss.get_root_expressions()]),
            hyper_name=np.array(hyp),
            y=y,
            Df_eval_cand=D_cross,
            Df_eval_eval=D_self,
            d_eval_cand=d_cross,
            K_eval_cand=K_EN,
            K_eval_eval=K_EE,
            mu_eval=mu_e,
            sigma_eval=sig_e,
            current_max=np.array(cur),
            mu=mu,
            sigma=sig,
            ei=ei,
            ucb=ucb,
        )
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(
            name,
            "E=%d N=%d" % (len(x_eval), len(x_cand)),
            "F_sub=%d" % arrays["F_cand_joint_" + rn.SUBTREES].shape[1],
            "max leaves %d" % max(x.count_elementary_expressions() for x in x_cand),
            "%.1f KB" % (os.path.getsize(path) / 1024),
        )


if __name__ == "__main__":
    main()
