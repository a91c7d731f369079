"""(mu + lambda)-style evolutionary acquisition optimiser over objects (mirror of
bosot/bayesian_optimization/evolutionary_optimizer_objects.py:24-71)."""
import logging

import numpy as np

from .base_candidate_generator import CandidateGenerator

logger = logging.getLogger(__name__)


class EvolutionaryOptimizerObjects:
    def __init__(self, population_size, num_offspring, candidate_generator: CandidateGenerator) -> None:
        self.candidate_generator = candidate_generator
        self.population_size = population_size
        self.num_offspring = num_offspring
        self.survival_rate = 1 / (self.num_offspring + 1)
        self.current_maximizer = None
        self.current_maximizer_value = -1 * np.inf

    def maximize(self, func, steps):
        population = self.candidate_generator.get_initial_for_evolutionary_opt(n_initial=self.population_size)
        function_values = None
        for step in range(steps):
            logger.info("EA step %d/%d", step + 1, steps)
            if step > 0:
                population = self.reproduce(self.select(population, function_values))
            function_values = func(population)
            fittest = int(np.argmax(function_values))
            if self.current_maximizer_value < function_values[fittest]:
                self.current_maximizer_value = function_values[fittest]
                self.current_maximizer = population[fittest]
        return self.current_maximizer, self.current_maximizer_value

    def select(self, population, function_values):
        # stable: equal acquisition values (duplicates in the population, commuted children) keep their population
        # order.  The reference calls np.argsort with the default kind, whose order of EQUAL values depends on the
        # NumPy build (introsort or the AVX-512 sorter); the stable order is the portable reading, and it is what the
        # device-side selection of the flat EA (bosot_topk_f64) implements.
        order = np.argsort(function_values, kind="stable")
        n_survive = int(self.population_size * self.survival_rate)
        return [population[i] for i in order[len(order) - n_survive :]]

    def reproduce(self, survivors):
        new_generation = []
        for survivor in survivors:
            new_generation += self.candidate_generator.get_around_candidate_for_evolutionary_opt(survivor, self.num_offspring)
        return new_generation + survivors
