"""Candidate lists for the BO loop and its evolutionary acquisition optimiser (mirror of
bosot/kernels/kernel_grammar/kernel_grammar_candidate_generator.py:27-117; same np.random calls in
the same order, so a seeded run proposes the same kernels as the reference)."""
import logging
from typing import List

import numpy as np

from ...bayesian_optimization.base_candidate_generator import CandidateGenerator
from .kernel_grammar_search_spaces import BaseKernelGrammarSearchSpace

logger = logging.getLogger(__name__)


class KernelGrammarCandidateGenerator(CandidateGenerator):
    def __init__(
        self,
        search_space: BaseKernelGrammarSearchSpace,
        n_initial_factor_trailing: int,
        n_exploration_trailing: int,
        exploration_p_geometric: float,
        n_exploitation_trailing: int,
        walk_length_exploitation_trailing: int,
        do_random_walk_exploitation_trailing: bool,
        **kwargs
    ):
        assert n_exploitation_trailing % walk_length_exploitation_trailing == 0
        self.search_space = search_space
        self.n_initial_trailing = n_initial_factor_trailing * search_space.get_num_base_kernels()
        self.n_exploration_trailing = n_exploration_trailing
        self.exploration_p_geometric = exploration_p_geometric
        self.n_exploitation_trailing = n_exploitation_trailing
        self.walk_length_exploitation_trailing = walk_length_exploitation_trailing
        self.do_random_walk_exploitation_trailing = do_random_walk_exploitation_trailing

    def get_random_canditates(self, n_candidates: int, seed=100, set_seed=False) -> List[object]:
        B = self.search_space.get_num_base_kernels()
        assert n_candidates % B == 0
        if set_seed:
            np.random.seed(seed)
        depth = int(n_candidates / B - 1)
        out = []
        for root in self.search_space.get_root_expressions():
            out.append(root)
            out += self.search_space.random_walk(depth, root)
        return out

    def get_initial_candidates_trailing(self) -> List[object]:
        return self.get_random_canditates(self.n_initial_trailing)

    def get_additional_candidates_trailing(self, best_current_candidate: object) -> List[object]:
        roots = self.search_space.get_root_expressions()
        out = []
        for _ in range(self.n_exploration_trailing):
            start = np.random.choice(roots)
            length = np.random.geometric(self.exploration_p_geometric)
            out += self.search_space.random_walk(length, start)
        if self.do_random_walk_exploitation_trailing:
            for _ in range(int(self.n_exploitation_trailing / self.walk_length_exploitation_trailing)):
                out += self.search_space.random_walk(self.walk_length_exploitation_trailing, best_current_candidate)
        else:
            out += self.search_space.get_neighbour_expressions(best_current_candidate)
        return out

    def get_around_candidate_for_evolutionary_opt(self, candidate: object, n_around_candidate: int):
        return [self.search_space.get_random_neighbour_expression(candidate) for _ in range(n_around_candidate)]

    def get_initial_for_evolutionary_opt(self, n_initial):
        return self.get_dataset_recursivly_generated(n_initial, 1)

    def get_dataset_recursivly_generated(self, n_data, n_per_step, filter_out_equivalent_expressions=False):
        pool = self.search_space.get_root_expressions()
        while len(pool) < n_data:
            chosen = np.random.choice(pool)
            for _ in range(n_per_step):
                if len(pool) < n_data:
                    new = self.search_space.get_random_neighbour_expression(chosen)
                    if filter_out_equivalent_expressions:
                        while self.check_if_equivalent_expression_in_list(new, pool):
                            logger.info("Expression already in list - sample new neighbour")
                            new = self.search_space.get_random_neighbour_expression(chosen)
                    pool.append(new)
        return pool

    def check_if_equivalent_expression_in_list(self, expression, expression_list):
        return any(self.search_space.check_expression_equality(e, expression) for e in expression_list)
