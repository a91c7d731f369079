"""Bayesian optimisation over structured objects: drop-in for
bosot/bayesian_optimization/bayesian_optimizer_objects.py:40-280.

The loop, its bookkeeping (timers, query / best lists, validation metric) and the acquisition
optimisers are host code as in the reference; ``acquisition_function`` -- the hot path -- runs on the
GPU: one fused pass turns the candidate list into EI / UCB values (BOO:159-168), with the
evaluated-set state (index, Gram, Cholesky, current_max) prepared once per model update instead of
being recomputed inside every call.
"""
import logging
import random
import time
from typing import List

import numpy as np
from scipy.stats import norm

from .. import _lib
from ..models.gp_model import PredictionQuantity
from ..models.object_gp_model import ObjectGpModel
from .base_candidate_generator import CandidateGenerator
from .enums import AcquisitionFunctionType, AcquisitionOptimizationObjectBOType, ValidationType
from .evolutionary_optimizer_objects import EvolutionaryOptimizerObjects

logger = logging.getLogger(__name__)


class BayesianOptimizerObjects:
    def __init__(
        self,
        acquisition_function_type: AcquisitionFunctionType,
        validation_type: ValidationType,
        acquisiton_optimization_type: AcquisitionOptimizationObjectBOType,
        population_evolutionary: int,
        n_steps_evolutionary: int,
        num_offspring_evolutionary: int,
        n_prune_trailing: int,
        do_plotting: bool = False,
        **kwargs
    ):
        # acquisition set-up
        self.acquisition_function_type = acquisition_function_type
        self.acquisiton_optimization_type = acquisiton_optimization_type
        self.gp_ucb_beta, self.ei_xi = 3.0, 0.01
        self.population_evolutionary = population_evolutionary
        self.n_steps_evolutionary = n_steps_evolutionary
        self.num_offspring_evolutionary = num_offspring_evolutionary
        self.n_prune_trailing = n_prune_trailing
        # EA on flat records (GPU population, same np.random consumption => same proposals as the object EA);
        # switched off automatically when the set-up is not the plain kernel-grammar one
        self.use_flat_evolutionary = True
        # bookkeeping of a run (same attribute names as the reference, they are read by its scripts)
        self.validation_type = validation_type
        self.do_plotting = do_plotting
        self.additional_metrics_index_interval = 10
        for trace in ("current_bests", "query_list", "additional_metrics", "validation_metrics", "iteration_time_list",
                      "acquisition_time_list", "oracle_time_list", "query_durations_x_data"):
            setattr(self, trace, [])
        self.oracle = None

    # ---- wiring -----------------------------------------------------------------------------
    def set_oracle(self, oracle):
        self.oracle = oracle

    def set_model(self, model: ObjectGpModel):
        self.model = model

    def set_candidate_generator(self, generator: CandidateGenerator):
        self.candidate_generator = generator

    def set_duration_time_predictor(self, predictor):
        self.duration_time_predictor = predictor

    def set_evolutionary_opt_settings(self, population: int, n_steps: int, num_offspring: int):
        self.population_evolutionary = population
        self.n_steps_evolutionary = n_steps
        self.num_offspring_evolutionary = num_offspring

    def sample_train_set(self, n_data, seed=100, set_seed=False):
        self.x_data = self.candidate_generator.get_random_canditates(n_data, seed, set_seed)
        y_list = []
        for x in self.x_data:
            y, query_duration = self.oracle.query(x)
            y_list.append(y)
            self.query_durations_x_data.append(query_duration)
        self.y_data = np.expand_dims(np.array(y_list), axis=1)

    def set_train_set(self, x_train: List[object], y_train: np.ndarray, query_durations_x_train: List[float]):
        self.x_data = x_train
        self.y_data = y_train
        self.query_durations_x_data = query_durations_x_train
        assert len(self.query_durations_x_data) == len(self.x_data)

    # ---- one BO step --------------------------------------------------------------------------
    def update(self, step: int):
        if self.acquisition_function_type == AcquisitionFunctionType.RANDOM:
            return random.choice(self.candidates), None  # Python's generator, like the reference (BOO:136-137)
        if self.acquisition_function_type == AcquisitionFunctionType.EXPECTED_IMPROVEMENT_PER_SECOND:
            self.duration_time_predictor.fit(self.x_data, self.query_durations_x_data)
        self.model.reset_model()
        self.model.infer(self.x_data, self.y_data)
        if self.acquisiton_optimization_type == AcquisitionOptimizationObjectBOType.TRAILING_CANDIDATES:
            acquisition_values = self.acquisition_function(self.candidates)
            return self.candidates[int(np.argmax(acquisition_values))], acquisition_values
        if self._flat_evolutionary_possible():
            return self._update_flat_evolutionary(), None
        optimizer = EvolutionaryOptimizerObjects(self.population_evolutionary, self.num_offspring_evolutionary, self.candidate_generator)
        new_query, _ = optimizer.maximize(self.acquisition_function, self.n_steps_evolutionary)
        return new_query, None

    def _flat_evolutionary_possible(self) -> bool:
        return (
            self.use_flat_evolutionary
            and hasattr(self.candidate_generator, "search_space")
            and self.model.prediction_quantity == PredictionQuantity.PREDICT_F
            and not getattr(self.model.kernel, "transform_to_normal", False)
            and self.model.has_constant_mean()
            and self.acquisition_function_type in (AcquisitionFunctionType.GP_UCB, AcquisitionFunctionType.EXPECTED_IMPROVEMENT)
        )

    def _update_flat_evolutionary(self):
        """The EA of `update` with the whole population resident on the GPU as flat records: no Python
        object per candidate, acquisition values straight from the fused device pass."""
        from ..encoding import Grammar
        from ..flat_search import FlatCandidateGenerator, FlatEvolutionaryOptimizer, record_to_expression

        space = self.candidate_generator.search_space
        grammar = Grammar.get(space.name, space.input_dimension)
        acq_kind = _lib.ACQ_UCB if self.acquisition_function_type == AcquisitionFunctionType.GP_UCB else _lib.ACQ_EI
        engine = self.model.engine(acq_kind=acq_kind, xi=self.ei_xi, ucb_beta=self.gp_ucb_beta)
        ea = FlatEvolutionaryOptimizer(self.population_evolutionary, self.num_offspring_evolutionary, FlatCandidateGenerator(grammar), engine.device)
        best, _ = ea.maximize(lambda trees: engine.acquire_device(trees)[2], self.n_steps_evolutionary)
        return record_to_expression(best, space)

    def acquisition_function(self, x_grid: List[object]) -> np.ndarray:
        """EI / GP-UCB / EI-per-second for every candidate (BOO:159-178), numpy fp64 [N]."""
        kind = self.acquisition_function_type
        if self.model.prediction_quantity == PredictionQuantity.PREDICT_F and self.model.has_constant_mean() and kind in (
            AcquisitionFunctionType.GP_UCB, AcquisitionFunctionType.EXPECTED_IMPROVEMENT, AcquisitionFunctionType.EXPECTED_IMPROVEMENT_PER_SECOND
        ):
            # fused device path: pairs -> Gram -> posterior -> acquisition in one pass over the candidates
            acq_kind = _lib.ACQ_UCB if kind == AcquisitionFunctionType.GP_UCB else _lib.ACQ_EI
            engine = self.model.engine(acq_kind=acq_kind, xi=self.ei_xi, ucb_beta=self.gp_ucb_beta)
            flat = self.model.kernel.flatten(self.model.transform_input(x_grid))
            _, _, acq = engine.acquire_device(flat.trees)
            acquisition_function = acq.cpu().numpy()
            engine.raise_on_bad_trees()  # the stream is idle after the copy: reading the count costs one 4-byte D2H
        else:
            # PREDICT_Y: sigma carries the observation noise; same host formulas as the reference
            pred_mu_grid, pred_sigma_grid = self.model.predictive_dist(x_grid)
            if kind == AcquisitionFunctionType.GP_UCB:
                acquisition_function = pred_mu_grid + np.sqrt(self.gp_ucb_beta) * pred_sigma_grid
            else:
                pred_mu_data, _ = self.model.predictive_dist(self.x_data)
                d = pred_mu_grid - np.max(pred_mu_data) - self.ei_xi
                acquisition_function = d * norm.cdf(d / pred_sigma_grid) + pred_sigma_grid * norm.pdf(d / pred_sigma_grid)
        if kind == AcquisitionFunctionType.EXPECTED_IMPROVEMENT_PER_SECOND:
            duration_predictions = self.duration_time_predictor.predict(x_grid)
            assert len(acquisition_function) == len(duration_predictions)
            acquisition_function = acquisition_function / duration_predictions
        return acquisition_function

    def maximize(self, n_steps: int):
        self.n_steps = n_steps
        if (
            self.acquisiton_optimization_type == AcquisitionOptimizationObjectBOType.TRAILING_CANDIDATES
            or self.acquisition_function_type == AcquisitionFunctionType.RANDOM
        ):
            self.candidates = self.candidate_generator.get_initial_candidates_trailing()
        self.add_additional_metrics(0)
        self.validate()
        for i in range(self.n_steps):
            t_iter = time.perf_counter()
            query, acquisition_values_candidates = self.update(i)
            t_acq = time.perf_counter()
            new_y, query_duration = self.oracle.query(query)
            t_oracle = time.perf_counter()
            self.current_bests.append((self.get_current_best(), self.get_current_best_value()))
            self.query_list.append((query, float(new_y)))
            self.x_data.append(query)
            self.y_data = np.vstack((self.y_data, [new_y]))
            self.query_durations_x_data.append(query_duration)
            self.add_additional_metrics(i + 1)
            self.validate()
            if self.acquisiton_optimization_type == AcquisitionOptimizationObjectBOType.TRAILING_CANDIDATES:
                self.update_candidates(acquisition_values_candidates)
            self.acquisition_time_list.append(t_acq - t_iter)
            self.oracle_time_list.append(t_oracle - t_acq)
            self.iteration_time_list.append(time.perf_counter() - t_iter)
        return (
            np.array(self.validation_metrics),
            self.query_list,
            self.current_bests,
            self.additional_metrics,
            self.iteration_time_list,
            self.oracle_time_list,
            self.acquisition_time_list,
        )

    # ---- trailing-candidate bookkeeping -----------------------------------------------------
    def update_candidates(self, acquisition_values):
        self.candidates = self.get_pruned_candidates(acquisition_values)
        self.candidates = self.candidates + self.candidate_generator.get_additional_candidates_trailing(self.get_current_best())

    def get_pruned_candidates(self, acquisition_values):
        pruned = []
        best_indexes = np.argsort(-1 * acquisition_values)[: self.n_prune_trailing]
        assert acquisition_values[best_indexes[0]] >= acquisition_values[best_indexes[1]]
        for index in best_indexes:
            candidate = self.candidates[index]
            if not self.check_if_in_list(candidate, pruned) and not self.check_if_in_list(candidate, self.x_data):
                pruned.append(candidate)
        return pruned

    def check_if_in_list(self, object_element, object_list):
        name = str(object_element)
        return any(name == str(o) for o in object_list)

    def get_current_best(self):
        return self.x_data[int(np.argmax(self.y_data))]

    def get_current_best_value(self):
        return np.max(self.y_data)

    def validate(self):
        if self.validation_type == ValidationType.MAX_OBSERVED:
            self.validation_metrics.append(np.max(self.y_data))

    def add_additional_metrics(self, index):
        if index % self.additional_metrics_index_interval == 0 or index == self.n_steps:
            if hasattr(self.oracle, "query_on_test_set"):
                self.additional_metrics.append((index, *self.oracle.query_on_test_set(self.get_current_best())))
