"""Mean functions over objects (mirror of bosot/models/object_mean_functions.py:29-74).

A mean function is ``value()`` (its constant part, which the device kernels take as ``mean_c``) plus, for
means that depend on the expression (``BICMean``), a per-expression ``offsets(...)`` that the model subtracts from
the targets before the fit / factorisation and adds back to the posterior mean."""
import numpy as np

from ..kernels.kernel_kernel_grammar_tree import Parameter


class MeanFunction:
    def value(self) -> float:
        raise NotImplementedError

    def __call__(self, X):
        return np.full((len(X), 1), self.value(), dtype=np.float64)


class Zero(MeanFunction):
    def __init__(self, output_dim=1):
        self.output_dim = output_dim

    def value(self) -> float:
        return 0.0


class ObjectConstant(MeanFunction):
    def __init__(self, base_c=0.0, trainable=True, output_dim=1):
        self.output_dim = output_dim
        self.c = Parameter(base_c, trainable=trainable)

    def value(self) -> float:
        return float(self.c)


# trainable parameters of an elementary kernel = sum of tf.size over its trainable variables (what BICMean counts):
# RBF lengthscales[nd] + variance; Linear offset + variance; Periodic period + lengthscales[nd] + variance;
# RQ alpha + lengthscales[nd] + variance; nd = 1 for a kernel "<name>_on_<dim>", else the input dimension
# (bosot/kernels/{rbf,linear,periodic,rational_quadratic}_kernel.py, base_elementary_kernel.py:30-40)
_EXTRA_PARAMS = {"RBFWithPrior": 1, "LinearWithPrior": None, "PeriodicKernelWithPrior": 2, "RQWithPrior": 2}


def elementary_parameter_count(name: str, input_dimension: int) -> int:
    base, sep, _ = name.partition("_on_")
    if base not in _EXTRA_PARAMS:
        raise ValueError("BICMean: unknown elementary kernel %r" % name)
    if _EXTRA_PARAMS[base] is None:
        return 2
    return _EXTRA_PARAMS[base] + (1 if sep else int(input_dimension))


class BICMean(MeanFunction):
    """m(x) = c - n_params(x) * log(n) / n  (or without the division), n = ``bic_n_data``: the BIC penalty of the kernel
    an expression stands for as a prior mean (reference object_mean_functions.py:51-74)."""

    def __init__(self, bic_n_data: int, divide_with_n_data=True, base_c=0.0, trainable=True):
        self.divide_with_n_data = divide_with_n_data
        self.bic_n_data = float(bic_n_data)
        self.c = Parameter(base_c, trainable=trainable)

    def value(self) -> float:
        return float(self.c)

    def penalty_per_parameter(self) -> float:
        pen = np.log(self.bic_n_data)
        return pen / self.bic_n_data if self.divide_with_n_data else pen

    @staticmethod
    def parameter_counts_of_records(records: np.ndarray, grammar) -> np.ndarray:
        """Trainable-parameter count of every flat record (uint8 [N, 64]): leaf symbols looked up in a per-grammar table."""
        table = np.zeros(256, dtype=np.int64)
        for sidx, name in enumerate(grammar.symbols):
            table[sidx] = elementary_parameter_count(name, grammar.input_dimension)
        n = records[:, 0].astype(np.int64)
        toks = records[:, 1:]
        live = (np.arange(toks.shape[1])[None, :] < n[:, None]) & (toks < 0x80)
        return np.where(live, table[toks], 0).sum(axis=1)

    def offsets(self, flat) -> np.ndarray:
        """m(x) - c for every record of ``flat`` (``FlatTrees``: device records + grammar)."""
        return -self.parameter_counts_of_records(flat.trees.cpu().numpy(), flat.grammar) * self.penalty_per_parameter()

    def get_number_of_trainable_parameters(self, kernel) -> int:
        names = [t for t in kernel.name.replace("(", " ").replace(")", " ").split() if t not in ("+", "*")]
        return int(sum(elementary_parameter_count(nm, kernel.get_input_dimension()) for nm in names))

    def __call__(self, X):
        """[len(X), 1] like the reference (:58-66), from the flat encoding of the expressions (host only)."""
        from ..encoding import Grammar, encode

        grammar = Grammar.for_expressions(X)
        counts = self.parameter_counts_of_records(encode(X, grammar), grammar)
        return (self.value() - counts * self.penalty_per_parameter()).reshape(-1, 1).astype(np.float64)
