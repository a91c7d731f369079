"""Toy kernel search with BO over kernel space on the B200 path -- the counterpart of
bosot/examples/bayesian_optimization_kernel_space.py:51-120 (CKSWithRQ search space, d = 2, EI + EA).

The reference's objective (Laplace model evidence of a GP fitted to simulated data through GPflow) is
out of scope; a deterministic structural objective stands in for it.  Needs a CUDA device."""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bosot_b200.compat  # noqa: E402

bosot_b200.compat.install_as_bosot()

import numpy as np  # noqa: E402
from bosot.bayesian_optimization.bayesian_optimizer_objects import BayesianOptimizerObjects  # noqa: E402
from bosot.configs.bayesian_optimization.bayesian_optimizer_objects_configs import ObjectBOExpectedImprovementEAConfig  # noqa: E402
from bosot.configs.kernels.grammar_tree_kernel_kernel_configs import OTWeightedDimsExtendedGrammarKernelConfig  # noqa: E402
from bosot.configs.kernels.kernel_grammar_generators.cks_with_rq_generator_config import CKSWithRQGeneratorConfig  # noqa: E402
from bosot.configs.models.object_gp_model_config import BasicObjectGPModelConfig  # noqa: E402
from bosot.kernels.kernel_grammar.generator_factory import GeneratorFactory  # noqa: E402
from bosot.models.gp_model import PredictionQuantity  # noqa: E402
from bosot.models.model_factory import ModelFactory  # noqa: E402
from bosot.models.object_mean_functions import ObjectConstant  # noqa: E402
from bosot.oracles.structural_oracle import StructuralOracle  # noqa: E402

logging.basicConfig(level=logging.WARNING)
N_BO_STEPS = int(os.environ.get("N_BO_STEPS", "40"))
N_STEPS_EVOLUTIONARY_OPTIMIZER = 3
POPULATION_EV_OPTIMIZER = 100

np.random.seed(0)
kernel_grammar_generator = GeneratorFactory.build(CKSWithRQGeneratorConfig(input_dimension=2))
oracle = StructuralOracle([str(r) for r in kernel_grammar_generator.search_space.get_root_expressions()], seed=7, noise=0.01)
object_gp = ModelFactory.build(
    BasicObjectGPModelConfig(
        kernel_config=OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=2),
        prediction_quantity=PredictionQuantity.PREDICT_F,
        perform_multi_start_optimization=False,
    )
)
object_gp.set_mean_function(ObjectConstant())
optimizer = BayesianOptimizerObjects(
    **ObjectBOExpectedImprovementEAConfig(n_steps_evolutionary=N_STEPS_EVOLUTIONARY_OPTIMIZER, population_evolutionary=POPULATION_EV_OPTIMIZER).dict()
)
optimizer.set_model(object_gp)
optimizer.set_oracle(oracle)
optimizer.set_candidate_generator(kernel_grammar_generator)
optimizer.sample_train_set(24)
metrics, queries, bests, _, t_iter, t_oracle, t_acq = optimizer.maximize(N_BO_STEPS)

# the run's records, under the file names the reference writes (bosot/main.py:89-107,343-348)
out_dir = os.environ.get("BOSOT_OUTPUT_DIR", os.path.join(os.path.dirname(os.path.abspath(__file__)), "output"))
os.makedirs(out_dir, exist_ok=True)
for pairs, list_name in ((queries, "QueryList"), (bests, "BestList")):
    with open(os.path.join(out_dir, list_name + ".txt"), "w") as f:
        for element, output in pairs:
            f.write(str(element) + " value=" + str(output) + "\n")
np.savetxt(os.path.join(out_dir, "iteration_time.txt"), np.array(t_iter))
np.savetxt(os.path.join(out_dir, "oracle_time.txt"), np.array(t_oracle))
np.savetxt(os.path.join(out_dir, "acquisition_time.txt"), np.array(t_acq))
print("Best kernel:", optimizer.get_current_best())
print("max observed: %.4f -> %.4f" % (metrics[0], metrics[-1]))
print("acquisition time per BO step (median): %.1f ms (%d EI evaluations each)" % (1e3 * np.median(t_acq), N_STEPS_EVOLUTIONARY_OPTIMIZER * POPULATION_EV_OPTIMIZER))
