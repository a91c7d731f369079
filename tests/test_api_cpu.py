import os
"""CPU: the host-side mirror of the reference API (expression trees, search spaces, candidate
generators, configs / factories).  The golden fixtures hold trees produced by the REFERENCE's own
generators under fixed np.random seeds, so reproducing their name strings pins enumeration order
and RNG consumption of the mirror (needed for "identical BO kernel selections")."""
import numpy as np
import pytest

from bosot_b200.configs.kernels.kernel_grammar_generators.generator_configs import (
    CKSHighDimGeneratorConfig,
    CKSWithRQGeneratorConfig,
    CompositionalKernelSearchGeneratorConfig,
    NDimFullKernelsGrammarGeneratorConfig,
)
from bosot_b200.encoding import Grammar, decode_to_name, encode
from bosot_b200.kernels.kernel_grammar.generator_factory import GeneratorFactory
from bosot_b200.kernels.kernel_grammar.kernel_grammar import (
    ElementaryKernelGrammarExpression,
    KernelGrammarExpression,
    KernelGrammarExpressionTransformer,
    KernelGrammarOperator,
    SymbolicKernel,
)

CFG = {
    "CKSWithRQSearchSpace": CKSWithRQGeneratorConfig,
    "CKSHighDimSearchSpace": CKSHighDimGeneratorConfig,
    "CompositionalKernelSearchSpace": CompositionalKernelSearchGeneratorConfig,
    "NDimFullKernelsSearchSpace": NDimFullKernelsGrammarGeneratorConfig,
}


def test_generators_reproduce_reference_trees(golden):
    if golden["case"].startswith(("walks", "c5_stress")):
        pytest.skip("fixture trees not produced by get_dataset_recursivly_generated")
    space = str(golden["generator_name"])
    gen = GeneratorFactory.build(CFG[space](input_dimension=int(golden["input_dimension"])))
    assert [str(r) for r in gen.search_space.get_root_expressions()] == [str(s) for s in golden["base_kernel_names"]]
    np.random.seed(3)
    mine = [str(x) for x in gen.get_dataset_recursivly_generated(len(golden["cand_names"]), 1)]
    ref = [str(s) for s in golden["cand_names"]]
    assert mine[:5] == ref[:5] and mine[7:] == ref[7:]  # rows 5 and 6 of the fixture were overwritten
    B = gen.search_space.get_num_base_kernels()
    E = len(golden["eval_names"])
    np.random.seed(1)
    first = [str(x) for x in gen.get_random_canditates(max(B, (E // B) * B))]
    n = min(len(first), E)
    assert first[:n] == [str(s) for s in golden["eval_names"]][:n]
    for x in gen.search_space.get_root_expressions():
        assert x.get_generator_name() == space


def test_walk_fixture_reproduced():
    from conftest import load_golden

    g = load_golden("walks_ckswithrq_d2")
    gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=2))
    ss = gen.search_space
    B = ss.get_num_base_kernels()
    np.random.seed(1)
    roots = ss.get_root_expressions()
    ev = [str(ss.random_walk(24, roots[i % B])[-1]) for i in range(len(g["eval_names"]))]
    assert ev == [str(s) for s in g["eval_names"]]


def _leaf(name, d=2):
    return ElementaryKernelGrammarExpression(SymbolicKernel(name, d))


def test_expression_tree_api():
    a, b, c = _leaf("RBFWithPrior_on_0"), _leaf("LinearWithPrior_on_1"), _leaf("RQWithPrior_on_0")
    e = KernelGrammarExpression(KernelGrammarExpression(a, b, KernelGrammarOperator.ADD), c, KernelGrammarOperator.MULTIPLY)
    assert str(e) == "((RBFWithPrior_on_0 ADD LinearWithPrior_on_1) MULTIPLY RQWithPrior_on_0)"
    assert e.count_elementary_expressions() == 3 and e.count_operators() == 2 and e.get_input_dimension() == 2
    assert e.get_indexes_of_subexpression() == [[-1], [0], [0, 0], [0, 1], [1]]
    assert e.get_indexes_of_elementary_expressions() == [[0, 0], [0, 1], [1]]
    assert e.get_expression_at_index([0, 1]) is b and e.get_expression_at_index([-1]) is e
    assert a.get_indexes_of_subexpression() == [[-1]] and a.get_indexes_of_elementary_expressions() == [[-1]]
    cp = e.deep_copy()
    assert str(cp) == str(e) and cp is not e and cp.get_generator_name() is None
    cp.change_elementary_expression_at_index([0, 0], SymbolicKernel("PeriodicKernelWithPrior_on_1", 2))
    cp.extend_sub_expression_at_index([-1], KernelGrammarOperator.ADD, SymbolicKernel("RQWithPrior_on_1", 2))
    assert str(cp) == "(((PeriodicKernelWithPrior_on_1 ADD LinearWithPrior_on_1) MULTIPLY RQWithPrior_on_0) ADD RQWithPrior_on_1)"
    assert str(e) == "((RBFWithPrior_on_0 ADD LinearWithPrior_on_1) MULTIPLY RQWithPrior_on_0)"
    e.set_generator_name("CKSWithRQSearchSpace")
    nf = KernelGrammarExpressionTransformer.transform_to_normal_form(e)
    assert str(nf) == "((RBFWithPrior_on_0 MULTIPLY RQWithPrior_on_0) ADD (LinearWithPrior_on_1 MULTIPLY RQWithPrior_on_0))"
    assert nf.get_generator_name() == "CKSWithRQSearchSpace"
    g = Grammar.get("CKSWithRQSearchSpace", 2)
    assert decode_to_name(encode([e], g)[0], g) == str(e)


def test_neighbourhood_enumeration():
    gen = GeneratorFactory.build(CKSHighDimGeneratorConfig(input_dimension=2))
    ss = gen.search_space
    root = ss.get_root_expressions()[1]
    assert ss.get_num_neighbour_expressions(root) == 4 + 2 * 4
    nb = ss.get_neighbour_expressions(root)
    assert [str(x) for x in nb[:4]] == [str(r) for r in ss.get_root_expressions()]
    assert str(nb[4]) == "(RBFWithPrior_on_1 ADD RBFWithPrior_on_0)" and str(nb[-1]) == "(RBFWithPrior_on_1 MULTIPLY RQWithPrior_on_1)"
    tree = nb[5]
    assert ss.get_num_neighbour_expressions(tree) == 2 * 4 + 3 * 2 * 4
    assert all(x.get_generator_name() == ss.name for x in ss.get_neighbour_expressions(tree))
    swapped = KernelGrammarExpression(tree.expression2, tree.expression1, tree.operator)
    assert ss.check_expression_equality(tree, swapped) and not ss.check_expression_equality(tree, nb[6])


def test_configs_behave_like_settings_objects():
    from bosot_b200.configs.bayesian_optimization.bayesian_optimizer_objects_configs import ObjectBOExpectedImprovementEAConfig
    from bosot_b200.configs.kernels.grammar_tree_kernel_kernel_configs import OTWeightedDimsExtendedGrammarKernelConfig
    from bosot_b200.kernels.kernel_kernel_grammar_tree import FeatureType

    k = OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=3)
    d = k.dict()
    assert d["feature_type_list"] == [FeatureType.DIM_WISE_WEIGHTED_ELEMENTARY_COUNT, FeatureType.REDUCED_ELEMENTARY_PATHS, FeatureType.SUBTREES]
    assert d["base_alpha"] == 0.5 and d["transform_to_normal"] is False and d["name"] == "OTWeightedDimsExtendedGrammarKernel"
    bo = ObjectBOExpectedImprovementEAConfig(n_steps_evolutionary=3)
    assert bo.dict()["num_offspring_evolutionary"] == 4 and bo.dict()["n_steps_evolutionary"] == 3
    with pytest.raises(TypeError):
        CKSWithRQGeneratorConfig()  # input_dimension is required
    with pytest.raises(TypeError):
        CKSWithRQGeneratorConfig(input_dimension=1, bogus=2)


def test_bosot_alias_never_imports_a_module_twice():
    """``import bosot.<anything>`` after ``install_as_bosot()`` yields THE module object of ``bosot_b200.<anything>``,
    also for modules outside the explicit alias list: the factories dispatch with isinstance, so a second copy of a
    config module would break them."""
    import importlib
    import subprocess
    import sys

    code = (
        "import bosot_b200.compat as c; c.install_as_bosot()\n"
        "import importlib\n"
        "names = ['configs.kernels.kernel_grammar_generators.generator_configs', 'oracles.structural_oracle',\n"
        "         'kernels.kernel_kernel_hellinger', 'configs.kernels.hellinger_kernel_kernel_configs', 'flat_search']\n"
        "for n in names:\n"
        "    assert importlib.import_module('bosot.' + n) is importlib.import_module('bosot_b200.' + n), n\n"
        "from bosot.configs.kernels.kernel_grammar_generators.generator_configs import CKSHighDimGeneratorConfig\n"
        "from bosot.kernels.kernel_grammar.generator_factory import GeneratorFactory\n"
        "assert type(GeneratorFactory.build(CKSHighDimGeneratorConfig(input_dimension=3))).__name__ == 'KernelGrammarCandidateGenerator'\n"
        "try:\n"
        "    import bosot.no_such_module\n"
        "    raise SystemExit(2)\n"
        "except ImportError:\n"
        "    pass\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_bic_mean_counts_parameters_like_the_reference():
    """BICMean (reference object_mean_functions.py:51-74): m(x) = c - n_params(x) log(n) / n with n_params = sum of
    tf.size over the trainable variables of the composed kernel: RBF 1 + nd, Linear 2, Periodic 2 + nd, RQ 2 + nd."""
    from bosot_b200.kernels.kernel_grammar.kernel_grammar import (
        ElementaryKernelGrammarExpression, KernelGrammarExpression, KernelGrammarOperator, SymbolicKernel)
    from bosot_b200.models.object_mean_functions import BICMean, elementary_parameter_count

    assert [elementary_parameter_count(n, 3) for n in ("RBFWithPrior_on_0", "LinearWithPrior_on_2", "PeriodicKernelWithPrior_on_1", "RQWithPrior_on_1")] == [2, 2, 3, 3]
    assert [elementary_parameter_count(n, 3) for n in ("RBFWithPrior", "LinearWithPrior", "PeriodicKernelWithPrior")] == [4, 2, 5]

    def leaf(name):
        e = ElementaryKernelGrammarExpression(SymbolicKernel(name, 2))
        e.set_generator_name("CKSWithRQSearchSpace")
        return e

    def node(op, a, b):
        e = KernelGrammarExpression(a, b, KernelGrammarOperator[op])
        e.set_generator_name("CKSWithRQSearchSpace")
        return e

    x = [leaf("LinearWithPrior_on_1"), node("ADD", node("MULTIPLY", leaf("RBFWithPrior_on_0"), leaf("RQWithPrior_on_1")), leaf("PeriodicKernelWithPrior_on_0"))]
    m = BICMean(150, base_c=0.25)
    np.testing.assert_allclose(m(x)[:, 0], 0.25 - np.array([2, 2 + 3 + 3]) * np.log(150.0) / 150.0, rtol=1e-15)
    m2 = BICMean(150, divide_with_n_data=False)
    np.testing.assert_allclose(m2(x)[:, 0], -np.array([2, 8]) * np.log(150.0), rtol=1e-15)
    assert m.get_number_of_trainable_parameters(x[1].get_kernel()) == 8


def test_config_picker_resolves_the_hot_path_configs():
    from bosot_b200.configs.config_picker import ConfigPicker

    assert ConfigPicker.pick_kernel_config("OTWeightedDimsExtendedGrammarKernelConfig")(input_dimension=2).input_dimension == 2
    assert ConfigPicker.pick_model_config("BasicObjectGPModelConfig").__name__ == "BasicObjectGPModelConfig"
    assert ConfigPicker.pick_bayesian_optimization_config("ObjectBOExpectedImprovementEAConfig")().population_evolutionary == 100
    assert ConfigPicker.pick_kernel_grammar_generator_config("CKSHighDimGeneratorConfig")(input_dimension=5).input_dimension == 5
    with pytest.raises(KeyError):
        ConfigPicker.pick_kernel_config("NoSuchConfig")
