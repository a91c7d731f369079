"""GPU: the reference-facing Python API (kernel-kernel class, GP over kernel space, BO loop) against
the golden fixtures and the CPU oracle.  Reads like the reference's call sites:
``kernel(X, X2)``, ``model.infer(x, y)``, ``model.predictive_dist(x)``, ``bo.acquisition_function(x)``."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from bosot_b200.bayesian_optimization.bayesian_optimizer_objects import BayesianOptimizerObjects  # noqa: E402
from bosot_b200.bayesian_optimization.enums import AcquisitionFunctionType  # noqa: E402
from bosot_b200.configs.bayesian_optimization.bayesian_optimizer_objects_configs import (  # noqa: E402
    ObjectBOExpectedImprovementEAConfig,
    ObjectBOGPUCBConfig,
)
from bosot_b200.configs.kernels.grammar_tree_kernel_kernel_configs import OTWeightedDimsExtendedGrammarKernelConfig  # noqa: E402
from bosot_b200.configs.kernels.kernel_grammar_generators.generator_configs import CKSHighDimGeneratorConfig, CKSWithRQGeneratorConfig  # noqa: E402
from bosot_b200.configs.models.object_gp_model_config import BasicObjectGPModelConfig  # noqa: E402
from bosot_b200.kernels.kernel_factory import KernelFactory  # noqa: E402
from bosot_b200.kernels.kernel_grammar.generator_factory import GeneratorFactory  # noqa: E402
from bosot_b200.kernels.kernel_kernel_grammar_tree import FeatureType, OptimalTransportKernelKernel  # noqa: E402
from bosot_b200.models.gp_model import PredictionQuantity  # noqa: E402
from bosot_b200.models.model_factory import ModelFactory  # noqa: E402
from bosot_b200.models.object_mean_functions import ObjectConstant  # noqa: E402
from bosot_b200.oracles.structural_oracle import StructuralOracle  # noqa: E402
from oracle import ref_numpy as rn  # noqa: E402

pytestmark = pytest.mark.gpu


def _expressions(names, generator_name, d):
    """Name strings -> expression objects of the mirror API (what a user of the reference holds)."""
    from bosot_b200.kernels.kernel_grammar.kernel_grammar import (
        ElementaryKernelGrammarExpression, KernelGrammarExpression, KernelGrammarOperator, SymbolicKernel)

    def build(t):
        if isinstance(t, str):
            e = ElementaryKernelGrammarExpression(SymbolicKernel(t, d))
        else:
            e = KernelGrammarExpression(build(t[1]), build(t[2]), KernelGrammarOperator[t[0]])
        e.set_generator_name(generator_name)
        return e

    return [build(rn.parse_name(str(s))) for s in names]


def _kernel(g, hyper):
    k = KernelFactory.build(OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=int(g["input_dimension"])))
    assert isinstance(k, OptimalTransportKernelKernel)
    k.alphas.assign(hyper.alphas)
    k.lengthscale.assign(hyper.lengthscale)
    k.variance.assign(hyper.variance)
    return k


def test_kernel_object_api(golden):
    gen, d = str(golden["generator_name"]), int(golden["input_dimension"])
    h = rn.Hyper(**getattr(rn, str(golden["hyper_name"])))
    xe, xc = _expressions(golden["eval_names"], gen, d), _expressions(golden["cand_names"], gen, d)
    k = _kernel(golden, h)
    np.testing.assert_allclose(k(xe, xc).cpu().numpy(), golden["K_eval_cand"], rtol=1e-12)       # kernel(X_train, X_new)
    np.testing.assert_allclose(k(xe).cpu().numpy(), golden["K_eval_eval"], rtol=1e-12)           # kernel(X_train)
    assert np.array_equal(k(xc, full_cov=False).cpu().numpy(), np.full(len(xc), h.variance))     # kernel(X_new, full_cov=False)
    with pytest.raises(ValueError):
        k(xe, xc, full_cov=False)                                                                # base_object_kernel.py:37
    np.testing.assert_allclose(k.get_distance_matrix(xe, xc).cpu().numpy(), golden["d_eval_cand"], atol=1e-12)
    # the reference's own dense-matrix route, through the mirrored methods
    for ft, ref in ((FeatureType.DIM_WISE_WEIGHTED_ELEMENTARY_COUNT, rn.DIM_WISE), (FeatureType.REDUCED_ELEMENTARY_PATHS, rn.PATHS), (FeatureType.SUBTREES, rn.SUBTREES)):
        de, dc = k.internal_transform_X(xe, ft), k.internal_transform_X(xc, ft)
        index = k.create_hash_array_mapping(de + dc)
        assert np.array_equal(k.feature_matrix(de, index, k.normalize_features(ft)), golden["F_eval_joint_" + ref])
        assert np.array_equal(k.feature_matrix(dc, index, k.normalize_features(ft)), golden["F_cand_joint_" + ref])


def test_object_gp_model_and_acquisition(golden):
    gen, d = str(golden["generator_name"]), int(golden["input_dimension"])
    h = rn.Hyper(**getattr(rn, str(golden["hyper_name"])))
    xe, xc = _expressions(golden["eval_names"], gen, d), _expressions(golden["cand_names"], gen, d)
    cfg = BasicObjectGPModelConfig(kernel_config=OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=d), optimize_hps=False,
                                   observation_noise=float(np.sqrt(h.noise_variance)), prediction_quantity=PredictionQuantity.PREDICT_F)
    model = ModelFactory.build(cfg)
    model.kernel.alphas.assign(h.alphas)
    model.kernel.lengthscale.assign(h.lengthscale)
    model.kernel.variance.assign(h.variance)
    model.set_mean_function(ObjectConstant(base_c=h.mean_c))
    model.infer(xe, golden["y"])
    mu, sigma = model.predictive_dist(xc)
    assert mu.shape == (len(xc),) and isinstance(mu, np.ndarray)
    np.testing.assert_allclose(mu, golden["mu"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(sigma, golden["sigma"], rtol=1e-6, atol=1e-12)
    for cfg_cls, key in ((ObjectBOExpectedImprovementEAConfig, "ei"), (ObjectBOGPUCBConfig, "ucb")):
        bo = BayesianOptimizerObjects(**cfg_cls().dict())
        bo.set_model(model)
        bo.set_train_set(xe, golden["y"], [0.0] * len(xe))
        acq = bo.acquisition_function(xc)
        assert isinstance(acq, np.ndarray) and acq.shape == (len(xc),)
        np.testing.assert_allclose(acq, golden[key], rtol=1e-6, atol=1e-12)
    # PREDICT_Y: sigma carries the likelihood variance (gp_model.py:303-304)
    model.prediction_quantity = PredictionQuantity.PREDICT_Y
    _, sig_y = model.predictive_dist(xc)
    np.testing.assert_allclose(sig_y, np.sqrt(golden["sigma"] ** 2 + h.noise_variance), rtol=1e-6)
    bo = BayesianOptimizerObjects(**ObjectBOExpectedImprovementEAConfig().dict())
    bo.set_model(model)
    bo.set_train_set(xe, golden["y"], [0.0] * len(xe))
    mu_e = golden["mu_eval"]
    dd = golden["mu"] - np.max(mu_e) - 0.01
    from scipy.stats import norm
    np.testing.assert_allclose(bo.acquisition_function(xc), dd * norm.cdf(dd / sig_y) + sig_y * norm.pdf(dd / sig_y), rtol=1e-6, atol=1e-12)


def test_marginal_likelihood_and_gradient():
    gen = GeneratorFactory.build(CKSHighDimGeneratorConfig(input_dimension=3))
    np.random.seed(5)
    x = gen.get_dataset_recursivly_generated(30, 1)
    oracle = StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=1)
    y = np.array([[oracle.query(e)[0]] for e in x])  # same tree -> same value: a well-posed regression problem
    cfg = BasicObjectGPModelConfig(kernel_config=OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=3), optimize_hps=False,
                                   prediction_quantity=PredictionQuantity.PREDICT_F)
    model = ModelFactory.build(cfg)
    model.set_mean_function(ObjectConstant(base_c=0.2))
    model.infer(x, y)
    # value against the NumPy restatement of GPR.log_marginal_likelihood
    trees = [rn.parse_name(str(e)) for e in x]
    hyp = rn.Hyper(noise_variance=1e-4, mean_c=0.2)
    K_EE = rn.K(trees, None, hyp, "CKSHighDimSearchSpace", 3)
    lml = rn.log_marginal_likelihood(K_EE, y, hyp)  # Cholesky route, as GPflow computes it
    np.testing.assert_allclose(lml, rn.log_marginal_likelihood_direct(K_EE, y, hyp), rtol=1e-10)  # solve + slogdet route
    np.testing.assert_allclose(model.estimate_model_evidence(), lml, rtol=1e-9)
    # analytic (autograd) gradient against central differences in unconstrained space
    layout, theta, _ = model._pack()
    assert [n for n, _ in layout] == ["alphas", "lengthscale", "variance", "noise", "c"] and len(theta) == 7
    f0, g = model._loss_and_grad(theta, layout)
    for i in range(len(theta)):
        e = np.zeros_like(theta)
        e[i] = 1e-5
        num = (model._loss_and_grad(theta + e, layout)[0] - model._loss_and_grad(theta - e, layout)[0]) / 2e-5
        np.testing.assert_allclose(g[i], num, rtol=1e-3, atol=1e-5 * max(1.0, abs(f0), np.abs(g).max()))
    # fused kernel vs torch autograd: same loss, same gradient
    assert model.use_fused_fit
    model.use_fused_fit = False
    f_ref, g_ref = model._loss_and_grad(theta, layout)
    model.use_fused_fit = True
    np.testing.assert_allclose(f0, f_ref, rtol=1e-11)
    np.testing.assert_allclose(g, g_ref, rtol=1e-7, atol=1e-9 * max(1.0, np.abs(g_ref).max()))
    for trial in range(3):  # away from the defaults too, incl. the exponential prior on the noise
        th = theta + np.random.default_rng(trial).normal(0, 0.4, size=theta.shape)
        model.set_prior_on_observation_noise = trial == 2
        fa, ga = model._loss_and_grad_fused(th, layout)
        fb, gb = model._loss_and_grad_autograd(th, layout)
        np.testing.assert_allclose(fa, fb, rtol=1e-10)
        np.testing.assert_allclose(ga, gb, rtol=1e-6, atol=1e-9 * max(1.0, np.abs(gb).max()))
    model.set_prior_on_observation_noise = False
    # the fit lowers the loss and keeps the parameters in their domains
    model.optimize_hps = True
    np.random.seed(0)
    before = model.training_loss()
    model.optimize(True)
    assert model.training_loss() < before
    assert np.all((model.kernel.alphas.numpy() > 0) & (model.kernel.alphas.numpy() < 1))
    assert float(model.kernel.lengthscale) > 0 and float(model.kernel.variance) > 0 and float(model.likelihood_variance) >= 1e-6
    model.reset_model()
    assert np.array_equal(model.kernel.alphas.numpy(), [0.5, 0.5, 0.5]) and model.model is None


class _OracleAcquisitionBO(BayesianOptimizerObjects):
    """Same loop, acquisition_function replaced by the CPU oracle's (reference formulas, NumPy fp64)."""

    def __init__(self, hyper, gen_name, d, **kw):
        super().__init__(**kw)
        self._h, self._g, self._d = hyper, gen_name, d

    def update(self, step):
        from bosot_b200.bayesian_optimization.evolutionary_optimizer_objects import EvolutionaryOptimizerObjects

        opt = EvolutionaryOptimizerObjects(self.population_evolutionary, self.num_offspring_evolutionary, self.candidate_generator)
        q, _ = opt.maximize(self.acquisition_function, self.n_steps_evolutionary)
        return q, None

    def acquisition_function(self, x_grid):
        t = lambda xs: [rn.parse_name(str(x)) for x in xs]  # noqa: E731
        kind = "UCB" if self.acquisition_function_type == AcquisitionFunctionType.GP_UCB else "EI"
        return rn.acquisition_function(t(self.x_data), self.y_data, t(x_grid), self._h, self._g, self._d, kind=kind)["acq"]


class _CheckedBO(BayesianOptimizerObjects):
    """GPU loop that cross-checks every acquisition call against the CPU oracle."""

    def __init__(self, hyper, gen_name, d, **kw):
        super().__init__(**kw)
        self._h, self._g, self._d = hyper, gen_name, d
        self.calls = self.same_argmax = 0

    def acquisition_function(self, x_grid):
        gpu = super().acquisition_function(x_grid)
        t = lambda xs: [rn.parse_name(str(x)) for x in xs]  # noqa: E731
        cpu = rn.acquisition_function(t(self.x_data), self.y_data, t(x_grid), self._h, self._g, self._d)["acq"]
        np.testing.assert_allclose(gpu, cpu, rtol=1e-6, atol=1e-12)
        self.calls += 1
        # the maximiser agrees, or the two maximisers are tied within the tolerance (EI far tails cancel catastrophically)
        i, j = int(np.argmax(gpu)), int(np.argmax(cpu))
        assert i == j or abs(cpu[i] - cpu[j]) <= 1e-6 * abs(cpu[j]) + 1e-12
        self.same_argmax += int(i == j)
        return gpu


def _toy_model(hyp, d):
    # hyper-parameters go in through the config: BayesianOptimizerObjects.update() resets the kernel to
    # its INITIAL values before every fit (gp_model.py:92-98), exactly like the reference
    kcfg = OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=d, base_alpha=float(hyp.alphas[0]),
                                                     base_lengthscale=hyp.lengthscale, base_variance=hyp.variance)
    model = ModelFactory.build(BasicObjectGPModelConfig(
        kernel_config=kcfg, optimize_hps=False, observation_noise=float(np.sqrt(hyp.noise_variance)),
        prediction_quantity=PredictionQuantity.PREDICT_F))
    model.set_mean_function(ObjectConstant(base_c=hyp.mean_c))
    return model


def test_bo_loop_makes_the_same_selections_as_the_cpu_path():
    """Toy run (CKSWithRQ, d = 2, like bosot/examples/bayesian_optimization_kernel_space.py with a
    deterministic stand-in objective and fixed hyper-parameters).  With GP-UCB (no cancellation in
    the acquisition) the GPU-backed loop and the loop driven by the CPU oracle's acquisition values
    must query the same kernels in the same order; with EI every acquisition call of the GPU loop
    is checked against the CPU values and its maximiser."""
    d, space = 2, "CKSWithRQSearchSpace"
    hyp = rn.Hyper(alphas=(0.4, 0.4, 0.4), lengthscale=0.8, variance=1.3, noise_variance=1e-3, mean_c=-0.5)
    ucb_cfg = ObjectBOGPUCBConfig(n_steps_evolutionary=3, population_evolutionary=50, num_offspring_evolutionary=4).dict()
    from bosot_b200.bayesian_optimization.enums import AcquisitionOptimizationObjectBOType
    ucb_cfg["acquisiton_optimization_type"] = AcquisitionOptimizationObjectBOType.EVOLUTIONARY
    runs = []
    for which in ("gpu", "cpu"):
        gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=d))
        oracle = StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=3)
        if which == "gpu":
            bo = BayesianOptimizerObjects(**ucb_cfg)
            bo.set_model(_toy_model(hyp, d))
        else:
            bo = _OracleAcquisitionBO(hyp, space, d, **ucb_cfg)
        bo.set_oracle(oracle)
        bo.set_candidate_generator(gen)
        np.random.seed(42)
        bo.sample_train_set(16)
        metrics, queries, bests, *_ = bo.maximize(5)
        runs.append(([str(q) for q, _ in queries], [v for _, v in queries], str(bo.get_current_best()), metrics))
    assert runs[0][0] == runs[1][0], (runs[0][0], runs[1][0])
    np.testing.assert_allclose(runs[0][1], runs[1][1])
    assert runs[0][2] == runs[1][2]
    assert runs[0][3][-1] >= runs[0][3][0] and len(runs[0][3]) == 6

    gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=d))
    bo = _CheckedBO(hyp, space, d, **ObjectBOExpectedImprovementEAConfig(n_steps_evolutionary=3, population_evolutionary=50).dict())
    bo.use_flat_evolutionary = False  # the object EA calls acquisition_function, which is what this run cross-checks
    bo.set_model(_toy_model(hyp, d))
    bo.set_oracle(StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=3))
    bo.set_candidate_generator(gen)
    np.random.seed(42)
    bo.sample_train_set(16)
    metrics, queries, *_ = bo.maximize(5)
    assert bo.calls == 15 and bo.same_argmax >= 13 and len(queries) == 5


class _FittedOracleAcquisitionBO(BayesianOptimizerObjects):
    """The reference's loop with the ML-II fit ON (BOO:143-144): every step re-fits the hyper-parameters of the GP
    over kernels, reads (alphas, lengthscale, variance, noise, c) back from the model and hands them to the CPU
    oracle, whose acquisition values drive the object EA.  Only the fit is shared with the GPU loop (it consumes the
    global np.random stream through the perturbation, gp_model.py:263-282, so both loops must run it)."""

    def __init__(self, gen_name, d, **kw):
        super().__init__(**kw)
        self._g, self._d = gen_name, d
        self.hyper_trace, self.lml_checks = [], 0

    def update(self, step):
        from bosot_b200.bayesian_optimization.evolutionary_optimizer_objects import EvolutionaryOptimizerObjects

        self.model.reset_model()
        self.model.infer(self.x_data, self.y_data)
        k = self.model.kernel
        self._h = rn.Hyper(alphas=np.array(k.alphas.numpy(), copy=True), lengthscale=float(k.lengthscale), variance=float(k.variance),
                           noise_variance=float(self.model.likelihood_variance), mean_c=self.model.mean_function.value())
        self.hyper_trace.append((tuple(self._h.alphas), self._h.lengthscale, self._h.variance, self._h.noise_variance, self._h.mean_c))
        # the objective the fit minimised, at its optimum, against the oracle's log marginal likelihood (both routes)
        trees = [rn.parse_name(str(x)) for x in self.x_data]
        K_EE = rn.K(trees, None, self._h, self._g, self._d)
        lml = rn.log_marginal_likelihood(K_EE, self.y_data, self._h)
        np.testing.assert_allclose(lml, rn.log_marginal_likelihood_direct(K_EE, self.y_data, self._h), rtol=1e-9)
        np.testing.assert_allclose(-self.model.training_loss(), lml, rtol=1e-9, atol=1e-9)
        self.lml_checks += 1
        opt = EvolutionaryOptimizerObjects(self.population_evolutionary, self.num_offspring_evolutionary, self.candidate_generator)
        q, _ = opt.maximize(self.acquisition_function, self.n_steps_evolutionary)
        return q, None

    def acquisition_function(self, x_grid):
        t = lambda xs: [rn.parse_name(str(x)) for x in xs]  # noqa: E731
        kind = "UCB" if self.acquisition_function_type == AcquisitionFunctionType.GP_UCB else "EI"
        return rn.acquisition_function(t(self.x_data), self.y_data, t(x_grid), self._h, self._g, self._d, kind=kind)["acq"]


class _TracedBO(BayesianOptimizerObjects):
    """GPU loop that records the hyper-parameters every model update ends with."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.hyper_trace = []

    def update(self, step):
        out = super().update(step)
        k = self.model.kernel
        self.hyper_trace.append((tuple(np.array(k.alphas.numpy(), copy=True)), float(k.lengthscale), float(k.variance),
                                 float(self.model.likelihood_variance), self.model.mean_function.value()))
        return out


def _fitted_toy_model(d):
    """The toy example's model (bosot/examples/bayesian_optimization_kernel_space.py:87-95): ML-II fit with random
    perturbation, no multi-start, PREDICT_F, trainable constant mean."""
    model = ModelFactory.build(BasicObjectGPModelConfig(
        kernel_config=OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=d), optimize_hps=True,
        perform_multi_start_optimization=False, prediction_quantity=PredictionQuantity.PREDICT_F))
    model.set_mean_function(ObjectConstant())
    return model


@pytest.mark.parametrize("acq", ["UCB", "EI"])
def test_bo_loop_with_hyperparameter_fit_makes_the_same_selections_as_the_cpu_path(acq):
    """north_star: "identical BO kernel selections on the toy example" -- with the ML-II fit ON, as the reference
    runs it (reset_model(); infer() every step, BOO:143-144, gp_model.py:162-193).  Three loops from the same seed:
    GPU with the flat (device-resident) EA, GPU with the object EA, and the loop whose acquisition values come from
    the CPU oracle evaluated at the hyper-parameters each fit ended with.  All three must query the same kernels in
    the same order for 5 steps; the fits must end at the same hyper-parameters, and their objective must equal the
    oracle's log marginal likelihood there."""
    from bosot_b200.bayesian_optimization.enums import AcquisitionOptimizationObjectBOType

    d, space = 2, "CKSWithRQSearchSpace"
    cfg_cls = ObjectBOGPUCBConfig if acq == "UCB" else ObjectBOExpectedImprovementEAConfig
    cfg = cfg_cls(n_steps_evolutionary=3, population_evolutionary=50, num_offspring_evolutionary=4).dict()
    cfg["acquisiton_optimization_type"] = AcquisitionOptimizationObjectBOType.EVOLUTIONARY
    runs = {}
    for which in ("gpu_flat", "gpu_object", "cpu"):
        gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=d))
        if which == "cpu":
            bo = _FittedOracleAcquisitionBO(space, d, **cfg)
        else:
            bo = _TracedBO(**cfg)
            bo.use_flat_evolutionary = which == "gpu_flat"
        bo.set_model(_fitted_toy_model(d))
        bo.set_oracle(StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=3))
        bo.set_candidate_generator(gen)
        np.random.seed(42)
        bo.sample_train_set(16)
        _, queries, *_ = bo.maximize(5)
        runs[which] = ([str(q) for q, _ in queries], bo.hyper_trace)
        if which == "cpu":
            assert bo.lml_checks == 5
    assert runs["gpu_flat"][0] == runs["gpu_object"][0] == runs["cpu"][0], runs
    for a, b in zip(runs["gpu_flat"][1], runs["cpu"][1]):  # same fit, same np.random stream: the same optimum
        np.testing.assert_allclose(np.hstack([np.ravel(x) for x in a]), np.hstack([np.ravel(x) for x in b]), rtol=1e-9)
    trained = np.array([np.hstack([np.ravel(x) for x in h]) for h in runs["cpu"][1]])
    assert np.ptp(trained, axis=0).max() > 1e-3  # the fit really moved the hyper-parameters between steps


def test_flat_and_object_evolutionary_modes_agree():
    """BayesianOptimizerObjects proposes the same queries whether its EA runs on expression objects or on
    flat records resident on the GPU (same np.random consumption, same acquisition values)."""
    d = 2
    hyp = rn.Hyper(alphas=(0.4, 0.4, 0.4), lengthscale=0.8, variance=1.3, noise_variance=1e-3, mean_c=-0.5)
    runs = []
    for flat in (True, False):
        gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=d))
        bo = BayesianOptimizerObjects(**ObjectBOExpectedImprovementEAConfig(n_steps_evolutionary=3, population_evolutionary=50).dict())
        bo.use_flat_evolutionary = flat
        bo.set_model(_toy_model(hyp, d))
        bo.set_oracle(StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=3))
        bo.set_candidate_generator(gen)
        np.random.seed(42)
        bo.sample_train_set(16)
        _, queries, *_ = bo.maximize(4)
        runs.append([str(q) for q, _ in queries])
        assert all(q.get_generator_name() == "CKSWithRQSearchSpace" for q, _ in queries)
    assert runs[0] == runs[1], runs


def test_bic_mean_in_the_model_and_the_bo_loop():
    """A mean that depends on the expression (BICMean, reference object_mean_functions.py:51-74): the model subtracts
    m(X) from the targets and adds m(x*) to the posterior mean; against the NumPy posterior with that mean, and one BO
    step (which must leave the fused constant-mean acquisition path)."""
    from bosot_b200.models.object_mean_functions import BICMean

    gen_name, d = "CKSWithRQSearchSpace", 2
    gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=d))
    np.random.seed(11)
    x = gen.get_dataset_recursivly_generated(20, 1)
    xs = gen.get_dataset_recursivly_generated(60, 1)
    oracle = StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=2)
    y = np.array([[oracle.query(e)[0]] for e in x])
    hyp = rn.Hyper(alphas=(0.5, 0.5, 0.5), lengthscale=1.0, variance=1.0, noise_variance=1e-4, mean_c=0.3)
    mean = BICMean(150, base_c=hyp.mean_c)
    model = ModelFactory.build(BasicObjectGPModelConfig(kernel_config=OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=d),
                                                        optimize_hps=False, prediction_quantity=PredictionQuantity.PREDICT_F))
    model.set_mean_function(mean)
    model.infer(x, y)
    mu, sigma = model.predictive_dist(xs)
    t = lambda xs_: [rn.parse_name(str(e)) for e in xs_]  # noqa: E731
    K_EE, K_EN = rn.K(t(x), None, hyp, gen_name, d), rn.K(t(x), t(xs), hyp, gen_name, d)
    m_tr, m_te = mean(x)[:, 0], mean(xs)[:, 0]
    A = K_EE + hyp.noise_variance * np.eye(len(x))
    mu_ref = K_EN.T @ np.linalg.solve(A, y[:, 0] - m_tr) + m_te
    var_ref = hyp.variance - np.einsum("en,en->n", K_EN, np.linalg.solve(A, K_EN))
    np.testing.assert_allclose(mu, mu_ref, rtol=1e-6, atol=1e-9)
    ok = var_ref > 1e-9
    np.testing.assert_allclose(sigma[ok], np.sqrt(var_ref[ok]), rtol=1e-5)
    assert not model.has_constant_mean()
    bo = BayesianOptimizerObjects(**ObjectBOExpectedImprovementEAConfig(n_steps_evolutionary=2, population_evolutionary=30).dict())
    bo.set_model(model)
    bo.set_oracle(oracle)
    bo.set_candidate_generator(gen)
    bo.set_train_set(list(x), y, [0.0] * len(x))
    acq = bo.acquisition_function(xs)
    mu_d, _ = model.predictive_dist(x)
    dd = mu - np.max(mu_d) - 0.01
    from scipy.stats import norm as _norm
    np.testing.assert_allclose(acq[ok], (dd * _norm.cdf(dd / sigma) + sigma * _norm.pdf(dd / sigma))[ok], rtol=1e-9)
    _, queries, *_ = bo.maximize(1)
    assert len(queries) == 1


@pytest.mark.parametrize("n_eval", [1, 7, 16, 17, 50, 100, 160])
def test_fused_factorisation_matches_the_library_route(n_eval):
    """bosot_gp_factorize (one launch: Cholesky + inverse of the factor + packing + beta) against cuSOLVER / cuBLAS through
    torch.linalg on the same matrix: L^-1 and beta to 1e-9 relative of their scale; a matrix that is not positive definite
    raises like the library route (and like the reference's tf.linalg.cholesky, which the fit loop retries)."""
    from bosot_b200 import _lib, sot

    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(40 + n_eval)
    X = torch.rand((n_eval, 5), dtype=torch.float64, device=dev, generator=g)
    K = 1.3 * torch.exp(-torch.cdist(X, X, p=1.0) / 0.64)
    err = torch.randn(n_eval, dtype=torch.float64, device=dev, generator=g)
    noise = 1e-3
    Linv, beta = sot.gp_factorize(K, noise, err)  # fused route (n_eval <= bosot_gp_fit_max_eval())
    L = torch.linalg.cholesky(K + noise * torch.eye(n_eval, dtype=torch.float64, device=dev))
    Linv_ref = torch.linalg.solve_triangular(L, torch.eye(n_eval, dtype=torch.float64, device=dev), upper=False)
    beta_ref = torch.cholesky_solve(err.reshape(-1, 1), L).reshape(-1)
    ldl = (n_eval + 15) // 16 * 16
    packed_ref = torch.empty((ldl, ldl), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().bosot_pack_linv(Linv_ref.contiguous().data_ptr(), n_eval, n_eval, packed_ref.data_ptr(), ldl,
                                               torch.cuda.current_stream().cuda_stream), "pack_linv")
    scale = float(Linv_ref.abs().max())
    assert float((Linv - packed_ref).abs().max()) <= 1e-9 * scale
    assert float((beta - beta_ref).abs().max()) <= 1e-9 * float(beta_ref.abs().max() + 1.0)
    # the padding of the packed layout is zero (the posterior kernel multiplies it)
    assert float(Linv.abs().sum()) > 0 and torch.isfinite(Linv).all()
    bad = K.clone()
    bad[n_eval // 2, n_eval // 2] = -5.0
    with pytest.raises(RuntimeError, match="not positive definite"):
        sot.gp_factorize(bad, noise, err)
