"""Where does one BO step of the toy example spend its time? (fit vs acquisition optimisation)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bosot_b200.configs.kernels.grammar_tree_kernel_kernel_configs import OTWeightedDimsExtendedGrammarKernelConfig
from bosot_b200.configs.kernels.kernel_grammar_generators.generator_configs import CKSWithRQGeneratorConfig
from bosot_b200.configs.models.object_gp_model_config import BasicObjectGPModelConfig
from bosot_b200.configs.bayesian_optimization.bayesian_optimizer_objects_configs import ObjectBOExpectedImprovementEAConfig
from bosot_b200.kernels.kernel_grammar.generator_factory import GeneratorFactory
from bosot_b200.models.gp_model import PredictionQuantity
from bosot_b200.models.model_factory import ModelFactory
from bosot_b200.models.object_mean_functions import ObjectConstant
from bosot_b200.bayesian_optimization.bayesian_optimizer_objects import BayesianOptimizerObjects
from bosot_b200.oracles.structural_oracle import StructuralOracle

np.random.seed(0)
gen = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=2))
oracle = StructuralOracle([str(r) for r in gen.search_space.get_root_expressions()], seed=7, noise=0.01)
model = ModelFactory.build(BasicObjectGPModelConfig(kernel_config=OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=2),
                                                    prediction_quantity=PredictionQuantity.PREDICT_F, perform_multi_start_optimization=False))
model.set_mean_function(ObjectConstant())
for pop, flat in ((100, True), (100, False), (10000, True)):
    bo = BayesianOptimizerObjects(**ObjectBOExpectedImprovementEAConfig(n_steps_evolutionary=3, population_evolutionary=pop).dict())
    bo.use_flat_evolutionary = flat
    bo.set_model(model); bo.set_oracle(oracle); bo.set_candidate_generator(gen)
    bo.sample_train_set(24)
    n_eval = [0]
    orig = model._loss_and_grad
    def counted(*a, **k):
        n_eval[0] += 1
        return orig(*a, **k)
    model._loss_and_grad = counted
    for step in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        model.reset_model(); model.infer(bo.x_data, bo.y_data)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        q = bo._update_flat_evolutionary() if flat else None
        if not flat:
            from bosot_b200.bayesian_optimization.evolutionary_optimizer_objects import EvolutionaryOptimizerObjects
            q, _ = EvolutionaryOptimizerObjects(pop, 4, gen).maximize(bo.acquisition_function, 3)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("pop %5d flat=%s step %d: fit %.1f ms (%d loss evals), EA x3 generations %.1f ms" % (pop, flat, step, 1e3*(t1-t0), n_eval[0], 1e3*(t2-t1)))
        n_eval[0] = 0
        y, _ = oracle.query(q); bo.x_data.append(q); bo.y_data = np.vstack((bo.y_data, [y]))
    model._loss_and_grad = orig
