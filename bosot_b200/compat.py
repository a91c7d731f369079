"""Make ``import bosot...`` resolve to this package: ``bosot_b200.compat.install_as_bosot()``.

The mirror keeps the reference's module paths (``bosot.kernels.kernel_kernel_grammar_tree``,
``bosot.models.object_gp_model``, ``bosot.bayesian_optimization.bayesian_optimizer_objects`` ...),
so existing scripts only need this one call before their imports."""
import importlib
import sys

_MODULES = [
    "kernels", "kernels.base_object_kernel", "kernels.kernel_factory", "kernels.kernel_kernel_grammar_tree",
    "kernels.kernel_grammar", "kernels.kernel_grammar.kernel_grammar", "kernels.kernel_grammar.kernel_grammar_search_spaces",
    "kernels.kernel_grammar.kernel_grammar_candidate_generator", "kernels.kernel_grammar.generator_factory",
    "models", "models.gp_model", "models.model_factory", "models.object_gp_model", "models.object_mean_functions",
    "bayesian_optimization", "bayesian_optimization.enums", "bayesian_optimization.base_candidate_generator",
    "bayesian_optimization.evolutionary_optimizer_objects", "bayesian_optimization.bayesian_optimizer_objects",
    "configs", "configs.kernels", "configs.kernels.base_kernel_config", "configs.kernels.base_elementary_kernel_config",
    "configs.kernels.rbf_configs", "configs.kernels.linear_configs", "configs.kernels.periodic_configs",
    "configs.kernels.rational_quadratic_configs", "configs.kernels.grammar_tree_kernel_kernel_configs",
    "configs.kernels.kernel_grammar_generators", "configs.kernels.kernel_grammar_generators.base_kernel_grammar_generator_config",
    "configs.kernels.kernel_grammar_generators.cks_with_rq_generator_config",
    "configs.kernels.kernel_grammar_generators.cks_high_dim_generator_config",
    "configs.kernels.kernel_grammar_generators.compositional_kernel_search_configs",
    "configs.kernels.kernel_grammar_generators.n_dim_full_kernels_generators_configs",
    "configs.models", "configs.models.object_gp_model_config",
    "configs.bayesian_optimization", "configs.bayesian_optimization.bayesian_optimizer_objects_configs",
    "oracles",
]


def install_as_bosot() -> None:
    if "bosot" in sys.modules and not getattr(sys.modules["bosot"], "__bosot_b200_alias__", False):
        raise RuntimeError("a different `bosot` package is already imported")
    root = importlib.import_module("bosot_b200")
    root.__bosot_b200_alias__ = True
    sys.modules["bosot"] = root
    for name in _MODULES:
        sys.modules["bosot." + name] = importlib.import_module("bosot_b200." + name)
