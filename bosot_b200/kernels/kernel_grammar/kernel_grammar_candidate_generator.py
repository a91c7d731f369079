"""Candidate lists for the BO loop and its evolutionary acquisition optimiser.

Behavioural mirror of bosot/kernels/kernel_grammar/kernel_grammar_candidate_generator.py:27-117 of the
reference: same public methods and -- because a BO trajectory is only reproducible if the global
``np.random`` stream is consumed identically -- the same random calls in the same order
(``tests/test_api_cpu.py`` pins this against trees the reference itself generated)."""
import logging
from typing import List

import numpy as np

from ...bayesian_optimization.base_candidate_generator import CandidateGenerator
from .kernel_grammar_search_spaces import BaseKernelGrammarSearchSpace

logger = logging.getLogger(__name__)


class KernelGrammarCandidateGenerator(CandidateGenerator):
    def __init__(self, search_space: BaseKernelGrammarSearchSpace, n_initial_factor_trailing: int, n_exploration_trailing: int,
                 exploration_p_geometric: float, n_exploitation_trailing: int, walk_length_exploitation_trailing: int,
                 do_random_walk_exploitation_trailing: bool, **kwargs):
        if n_exploitation_trailing % walk_length_exploitation_trailing:
            raise AssertionError("n_exploitation_trailing must be a multiple of walk_length_exploitation_trailing")
        self.search_space = search_space
        self.n_initial_trailing = n_initial_factor_trailing * search_space.get_num_base_kernels()
        self.n_exploration_trailing, self.exploration_p_geometric = n_exploration_trailing, exploration_p_geometric
        self.n_exploitation_trailing = n_exploitation_trailing
        self.walk_length_exploitation_trailing = walk_length_exploitation_trailing
        self.do_random_walk_exploitation_trailing = do_random_walk_exploitation_trailing

    # ---- random starts ------------------------------------------------------------------------
    def get_random_canditates(self, n_candidates: int, seed=100, set_seed=False) -> List[object]:
        """Every base kernel followed by one random walk of equal length from it."""
        space = self.search_space
        walk, remainder = divmod(n_candidates, space.get_num_base_kernels())
        assert remainder == 0, "n_candidates must be a multiple of the number of base kernels"
        if set_seed:
            np.random.seed(seed)
        candidates: List[object] = []
        for root in space.get_root_expressions():
            candidates += [root] + space.random_walk(walk - 1, root)
        return candidates

    def get_initial_candidates_trailing(self) -> List[object]:
        return self.get_random_canditates(self.n_initial_trailing)

    # ---- trailing-candidate refresh -----------------------------------------------------------
    def get_additional_candidates_trailing(self, best_current_candidate: object) -> List[object]:
        space = self.search_space
        roots = space.get_root_expressions()
        fresh: List[object] = []
        for _ in range(self.n_exploration_trailing):  # exploration: geometric-length walks from random roots
            start = np.random.choice(roots)
            fresh += space.random_walk(np.random.geometric(self.exploration_p_geometric), start)
        if not self.do_random_walk_exploitation_trailing:  # exploitation: the whole neighbourhood of the incumbent ...
            return fresh + space.get_neighbour_expressions(best_current_candidate)
        step = self.walk_length_exploitation_trailing  # ... or fixed-length walks from it
        for _ in range(self.n_exploitation_trailing // step):
            fresh += space.random_walk(step, best_current_candidate)
        return fresh

    # ---- evolutionary optimiser ---------------------------------------------------------------
    def get_around_candidate_for_evolutionary_opt(self, candidate: object, n_around_candidate: int):
        draw = self.search_space.get_random_neighbour_expression
        return [draw(candidate) for _ in range(n_around_candidate)]

    def get_initial_for_evolutionary_opt(self, n_initial):
        return self.get_dataset_recursivly_generated(n_initial, 1)

    def get_dataset_recursivly_generated(self, n_data, n_per_step, filter_out_equivalent_expressions=False):
        """Grow a list from the base kernels: pick a random member, append up to ``n_per_step`` of its random
        neighbours, until ``n_data`` expressions exist."""
        space = self.search_space
        grown = space.get_root_expressions()
        while len(grown) < n_data:
            parent = np.random.choice(grown)
            for _ in range(min(n_per_step, n_data - len(grown))):
                child = space.get_random_neighbour_expression(parent)
                while filter_out_equivalent_expressions and self.check_if_equivalent_expression_in_list(child, grown):
                    logger.info("Expression already in list - sample new neighbour")
                    child = space.get_random_neighbour_expression(parent)
                grown.append(child)
        return grown

    def check_if_equivalent_expression_in_list(self, expression, expression_list):
        same = self.search_space.check_expression_equality
        return any(same(member, expression) for member in expression_list)
