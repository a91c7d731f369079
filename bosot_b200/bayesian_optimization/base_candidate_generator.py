"""Contract of a candidate generator as the BO loop and its acquisition optimisers use it
(reference: bosot/bayesian_optimization/base_candidate_generator.py:21-67).  Every method returns a list
of objects of the search space (kernel-grammar expressions on this path)."""
from abc import ABC, abstractmethod


def _required(name: str, doc: str):
    def method(self, *args, **kwargs):
        raise NotImplementedError("%s.%s" % (type(self).__name__, name))

    method.__name__ = name
    method.__doc__ = doc
    return abstractmethod(method)


class CandidateGenerator(ABC):
    # trailing-candidate acquisition optimisation
    get_initial_candidates_trailing = _required("get_initial_candidates_trailing", "() -> initial candidate list")
    get_additional_candidates_trailing = _required(
        "get_additional_candidates_trailing", "(best_current_candidate) -> candidates related to the incumbent plus fresh random ones")
    # generic random objects: (n_candidates, seed=100, set_seed=False)
    get_random_canditates = _required("get_random_canditates", "(n_candidates, seed=100, set_seed=False) -> n random objects")
    # evolutionary acquisition optimisation
    get_initial_for_evolutionary_opt = _required("get_initial_for_evolutionary_opt", "(n_initial) -> initial population")
    get_around_candidate_for_evolutionary_opt = _required(
        "get_around_candidate_for_evolutionary_opt", "(candidate, n_around_candidate) -> random neighbours of one individual")
    get_dataset_recursivly_generated = _required("get_dataset_recursivly_generated", "(n_data, n_per_step) -> recursively grown list")
