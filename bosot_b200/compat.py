"""Make ``import bosot...`` resolve to this package: ``bosot_b200.compat.install_as_bosot()``.

The mirror keeps the reference's module paths (``bosot.kernels.kernel_kernel_grammar_tree``,
``bosot.models.object_gp_model``, ``bosot.bayesian_optimization.bayesian_optimizer_objects`` ...),
so existing scripts only need this one call before their imports."""
import importlib
import importlib.abc
import importlib.util
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


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``bosot.<anything>`` -> the module object of ``bosot_b200.<anything>`` (never a second copy of a module:
    ``isinstance`` checks in the factories compare classes by identity)."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "bosot" and not fullname.startswith("bosot."):
            return None
        real = "bosot_b200" + fullname[len("bosot"):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, self)

    def create_module(self, spec):
        return importlib.import_module("bosot_b200" + spec.name[len("bosot"):])

    def exec_module(self, module):  # already executed under its real name
        pass


def install_as_bosot() -> None:
    if "bosot" in sys.modules and not getattr(sys.modules["bosot"], "__bosot_b200_alias__", False):
        raise RuntimeError("a different `bosot` package is already imported")
    root = importlib.import_module("bosot_b200")
    root.__bosot_b200_alias__ = True
    sys.modules["bosot"] = root
    if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder())
    for name in _MODULES:
        sys.modules["bosot." + name] = importlib.import_module("bosot_b200." + name)
