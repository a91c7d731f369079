"""CPU, build container only: restatement vs the LIVE reference featuriser (skipped where
/root/reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

from oracle import ref_numpy as rn
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.mark.parametrize("space,d", [("CKSWithRQSearchSpace", 2), ("CKSHighDimSearchSpace", 5), ("CompositionalKernelSearchSpace", 3), ("NDimFullKernelsSearchSpace", 2)])
def test_dense_features_match_live_reference(space, d):
    ref_shim.install()
    import bosot.kernels.kernel_grammar.kernel_grammar_search_spaces as kgss
    from bosot.configs.kernels.grammar_tree_kernel_kernel_configs import OTWeightedDimsExtendedGrammarKernelConfig
    from bosot.kernels.kernel_grammar.kernel_grammar_candidate_generator import KernelGrammarCandidateGenerator
    from bosot.kernels.kernel_kernel_grammar_tree import FeatureType, OptimalTransportKernelKernel

    ft = {rn.DIM_WISE: FeatureType.DIM_WISE_WEIGHTED_ELEMENTARY_COUNT, rn.PATHS: FeatureType.REDUCED_ELEMENTARY_PATHS, rn.SUBTREES: FeatureType.SUBTREES}
    ss = getattr(kgss, space)(d)
    gen = KernelGrammarCandidateGenerator(ss, 5, 15, 1 / 3, 50, 5, False)
    np.random.seed(11)
    X = gen.get_dataset_recursivly_generated(120, 1)
    X2 = [ss.random_walk(12, r)[-1] for r in ss.get_root_expressions()]
    k = OptimalTransportKernelKernel(**OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=d).dict())
    TX, TX2 = [rn.parse_name(str(x)) for x in X], [rn.parse_name(str(x)) for x in X2]
    for f in rn.DEFAULT_FEATURE_TYPES:
        dA, dB = k.internal_transform_X(X, ft[f]), k.internal_transform_X(X2, ft[f])
        idx = k.create_hash_array_mapping(dA + dB)
        FA, FB = rn.dense_features(TX, TX2, f, ss.name, d)
        assert np.array_equal(k.feature_matrix(dA, idx, k.normalize_features(ft[f])), FA)
        assert np.array_equal(k.feature_matrix(dB, idx, k.normalize_features(ft[f])), FB)
