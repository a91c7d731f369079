"""GPU: the Hellinger kernel-kernel behind the reference-facing API: ``KernelFactory`` ->
``KernelKernelHellinger`` -> ``ObjectGpModel`` -> ``BayesianOptimizerObjects``, against oracle/ref_hellinger.py and
the GP posterior of oracle/ref_numpy.py.  Reads like main.py:297-330 of the reference."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from bosot_b200.bayesian_optimization.bayesian_optimizer_objects import BayesianOptimizerObjects  # noqa: E402
from bosot_b200.configs.bayesian_optimization.bayesian_optimizer_objects_configs import ObjectBOExpectedImprovementEAConfig  # noqa: E402
from bosot_b200.configs.kernels.hellinger_kernel_kernel_configs import BasicHellingerKernelKernelConfig  # noqa: E402
from bosot_b200.configs.kernels.kernel_grammar_generators.generator_configs import CKSWithRQGeneratorConfig  # noqa: E402
from bosot_b200.configs.models.object_gp_model_config import BasicObjectGPModelConfig  # noqa: E402
from bosot_b200.kernels.kernel_factory import KernelFactory  # noqa: E402
from bosot_b200.kernels.kernel_grammar.generator_factory import GeneratorFactory  # noqa: E402
from bosot_b200.kernels.kernel_kernel_hellinger import KernelKernelHellinger  # noqa: E402
from bosot_b200.models.gp_model import PredictionQuantity  # noqa: E402
from bosot_b200.models.object_gp_model import ObjectGpModel  # noqa: E402
from bosot_b200.models.object_mean_functions import ObjectConstant  # noqa: E402
from bosot_b200.oracles.structural_oracle import StructuralOracle  # noqa: E402
from oracle import ref_hellinger as rh  # noqa: E402
from oracle import ref_numpy as rn  # noqa: E402

pytestmark = pytest.mark.gpu
D, S = 2, 24


def _setup(n=14, seed=5):
    gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=D))
    np.random.seed(seed)
    xs = gen.get_random_canditates(2 * 8, seed=seed, set_seed=True)[:n] + gen.search_space.random_walk(6, gen.search_space.get_root_expressions()[3])
    tuples = [rn.parse_name(str(x)) for x in xs]
    x_data = np.random.default_rng(seed).random((60, D))
    kernel = KernelFactory.build(BasicHellingerKernelKernelConfig(input_dimension=D, num_param_samples=S))
    assert isinstance(kernel, KernelKernelHellinger)
    state = np.random.get_state()
    kernel.set_virtual_x_from_dataset(x_data)
    np.random.set_state(state)
    assert np.array_equal(kernel.virtual_X, x_data[np.random.choice(60, 20, replace=False)])  # same draw as the reference (:255)
    return gen, xs, tuples, kernel


def test_kernel_object_api():
    _, xs, tuples, k = _setup()
    k.lengthscale.assign(0.7)
    k.variance.assign(1.4)
    X = k.virtual_X
    D_ref = rh.distance_matrix(tuples, None, X, S, D)
    ops = k.transform_X(xs)
    np.testing.assert_allclose(k.get_hellinger_distance_matrix(ops).cpu().numpy(), D_ref, rtol=0, atol=1e-7)
    K_ref = 1.4 * np.exp(-0.5 * (D_ref / 0.49))
    np.testing.assert_allclose(k(ops).cpu().numpy(), K_ref, rtol=1e-6)
    np.testing.assert_allclose(k(xs, xs[:5]).cpu().numpy(), K_ref[:, :5], rtol=1e-6)  # expression lists are accepted as well
    np.testing.assert_allclose(k(ops, full_cov=False).cpu().numpy(), np.full(len(xs), 1.4))
    with pytest.raises(ValueError):
        k(ops, ops, full_cov=False)
    sob = KernelFactory.build(BasicHellingerKernelKernelConfig(input_dimension=D, use_sobol_virtual_points=True, num_param_samples=8))
    assert np.array_equal(sob.virtual_X, rh.sobol_sample(D, 20))
    np.testing.assert_allclose(sob.get_hellinger_distance_matrix(xs[:6]).cpu().numpy(), rh.distance_matrix(tuples[:6], None, sob.virtual_X, 8, D),
                               rtol=0, atol=1e-7)


def test_object_gp_model_posterior():
    gen, xs, tuples, k = _setup()
    k.lengthscale.assign(0.4)
    k.variance.assign(0.9)
    y = StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=1)
    yv = np.array([y.query(x)[0] for x in xs]).reshape(-1, 1)
    E = 12
    model = ObjectGpModel(kernel=k, **BasicObjectGPModelConfig(kernel_config=None, optimize_hps=False, observation_noise=0.05,
                                                               prediction_quantity=PredictionQuantity.PREDICT_F).dict())
    model.set_mean_function(ObjectConstant())
    model.mean_function.c.assign(float(np.mean(yv[:E])))
    model.infer(xs[:E], yv[:E])
    mu, sigma = model.predictive_dist(xs)
    hyp = rn.Hyper(alphas=(1, 1, 1), lengthscale=0.4, variance=0.9, noise_variance=0.05 ** 2, mean_c=float(np.mean(yv[:E])))
    Kfull = rh.K(tuples, tuples[:E], k.virtual_X, S, D, 0.9, 0.4)
    mu_ref, sigma_ref = rn.gp_posterior_from_grams(Kfull[:E], Kfull.T, np.full(len(xs), 0.9), yv[:E, 0], hyp)
    np.testing.assert_allclose(mu, mu_ref, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(sigma, sigma_ref, rtol=1e-6, atol=1e-9)


def test_hyper_parameter_fit_fused_vs_autograd_and_hyperprior():
    _, xs, _, k = _setup()
    yv = np.sin(np.arange(len(xs)))[:, None]
    for use_prior in (False, True):
        kk = KernelFactory.build(BasicHellingerKernelKernelConfig(input_dimension=D, num_param_samples=S, use_hyperprior=use_prior))
        kk.virtual_X = k.virtual_X
        model = ObjectGpModel(kernel=kk, **BasicObjectGPModelConfig(kernel_config=None, optimize_hps=False, observation_noise=0.1).dict())
        model.set_mean_function(ObjectConstant())
        model.infer(xs, yv)
        layout, theta, _ = model._pack()
        assert [n for n, _ in layout] == ["lengthscale", "variance", "noise", "c"]
        theta = theta + np.array([0.3, -0.2, 0.1, 0.05])
        lf, gf = model._loss_and_grad_fused(theta, layout)
        la, ga = model._loss_and_grad_autograd(theta, layout)
        np.testing.assert_allclose(lf, la, rtol=1e-10)
        np.testing.assert_allclose(gf, ga, rtol=1e-6, atol=1e-9)
    # the fit itself lowers the loss and keeps the parameters positive
    model.optimize_hps = True
    before = model.training_loss()
    np.random.seed(0)
    model.optimize(False)
    assert model.training_loss() < before and float(model.kernel.lengthscale) > 0 and float(model.kernel.variance) > 0


def test_bayesian_optimisation_loop_runs_on_the_hellinger_kernel():
    gen, xs, _, k = _setup()
    model = ObjectGpModel(kernel=k, **BasicObjectGPModelConfig(kernel_config=None, perform_multi_start_optimization=False,
                                                               prediction_quantity=PredictionQuantity.PREDICT_F).dict())
    model.set_mean_function(ObjectConstant())
    runs = []
    for flat in (True, False):
        bo = BayesianOptimizerObjects(**ObjectBOExpectedImprovementEAConfig(n_steps_evolutionary=2, population_evolutionary=25).dict())
        bo.use_flat_evolutionary = flat
        bo.set_model(model)
        bo.set_oracle(StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=3))
        bo.set_candidate_generator(gen)
        np.random.seed(7)
        bo.sample_train_set(16)
        metrics, queries, *_ = bo.maximize(3)
        assert len(queries) == 3 and len(metrics) == 4 and metrics[-1] >= metrics[0]
        runs.append([str(q) for q, _ in queries])
    assert runs[0] == runs[1]  # flat (device-resident) and object evolutionary optimisers propose the same kernels
