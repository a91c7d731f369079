"""BO over kernel space at the shapes of BASELINE.json's configurations 1-4 (search space, initial set, EA settings of
the reference's main.py:255-337 / example :55-63), with the deterministic structural objective in place of the
reference's model-evidence oracle (out of scope).  Prints the time per BO step and its split.

    python examples/bo_baseline_configs.py [--steps 50] [--configs c1 c2 c3 c4]      # needs a CUDA device
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bosot_b200.compat  # noqa: E402

bosot_b200.compat.install_as_bosot()

import numpy as np  # noqa: E402
import torch  # noqa: E402
from bosot.bayesian_optimization.bayesian_optimizer_objects import BayesianOptimizerObjects  # noqa: E402
from bosot.configs.bayesian_optimization.bayesian_optimizer_objects_configs import (  # noqa: E402
    ObjectBOExpectedImprovementEAConfig,
    ObjectBOExpectedImprovementEAFewerStepsConfig,
)
from bosot.configs.kernels.grammar_tree_kernel_kernel_configs import OTWeightedDimsExtendedGrammarKernelConfig  # noqa: E402
from bosot.configs.kernels.kernel_grammar_generators.generator_configs import CKSHighDimGeneratorConfig, CKSWithRQGeneratorConfig  # noqa: E402
from bosot.configs.models.object_gp_model_config import BasicObjectGPModelConfig  # noqa: E402
from bosot.kernels.kernel_grammar.generator_factory import GeneratorFactory  # noqa: E402
from bosot.models.gp_model import PredictionQuantity  # noqa: E402
from bosot.models.model_factory import ModelFactory  # noqa: E402
from bosot.models.object_mean_functions import ObjectConstant  # noqa: E402
from bosot.oracles.structural_oracle import StructuralOracle  # noqa: E402

# name: (generator config, input dimension, BO config, EA population, EA generations)
CONFIGS = {
    "c1": ("toy, CKSWithRQ d=2", CKSWithRQGeneratorConfig, 2, ObjectBOExpectedImprovementEAConfig, 100, 3),
    "c2": ("Airline-shaped, CKSWithRQ d=1", CKSWithRQGeneratorConfig, 1, ObjectBOExpectedImprovementEAFewerStepsConfig, 100, 6),
    "c3": ("Powerplant-shaped, CKSHighDim d=4", CKSHighDimGeneratorConfig, 4, ObjectBOExpectedImprovementEAConfig, 100, 10),
    "c4": ("Airfoil-shaped, CKSHighDim d=5, 10k-candidate EA pool", CKSHighDimGeneratorConfig, 5, ObjectBOExpectedImprovementEAConfig, 10_000, 3),
}


def run(name: str, n_steps: int) -> None:
    title, gen_cfg, d, bo_cfg, pop, generations = CONFIGS[name]
    np.random.seed(0)
    generator = GeneratorFactory.build(gen_cfg(input_dimension=d))
    n_base = len(generator.search_space.get_root_expressions())
    oracle = StructuralOracle([str(r) for r in generator.search_space.get_root_expressions()], seed=7, noise=0.01)
    model = ModelFactory.build(BasicObjectGPModelConfig(kernel_config=OTWeightedDimsExtendedGrammarKernelConfig(input_dimension=d),
                                                        prediction_quantity=PredictionQuantity.PREDICT_F,
                                                        perform_multi_start_optimization=False))
    model.set_mean_function(ObjectConstant())
    bo = BayesianOptimizerObjects(**bo_cfg(n_steps_evolutionary=generations, population_evolutionary=pop).dict())
    bo.set_model(model)
    bo.set_oracle(oracle)
    bo.set_candidate_generator(generator)
    bo.sample_train_set(3 * n_base)  # main.py:130,263: n_initial = 3 x number of base kernels
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    metrics, queries, bests, _, t_iter, t_oracle, t_acq = bo.maximize(n_steps)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("%s  %-52s B=%2d E=%d->%d  EA %5d x %2d  %d BO steps in %.2f s: %.1f ms per step (model update + EA %.1f ms)  best %.4f -> %.4f"
          % (name, title, n_base, 3 * n_base, 3 * n_base + n_steps, pop, generations, n_steps, wall, 1e3 * wall / n_steps,
             1e3 * float(np.mean(t_acq)), metrics[0], metrics[-1]))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--configs", nargs="*", default=list(CONFIGS))
    args = ap.parse_args()
    for c in args.configs:
        run(c, args.steps)
