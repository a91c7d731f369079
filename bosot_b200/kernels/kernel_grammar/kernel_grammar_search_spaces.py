"""CKS-style neighbourhoods over kernel-grammar expressions (input producer of the hot path).

Same classes, method names and -- important for reproducing a BO trajectory -- the same
enumeration order and ``np.random`` consumption as
bosot/kernels/kernel_grammar/kernel_grammar_search_spaces.py: neighbour k of an expression is
  k <  L*B            : leaf  floor(k / B) (pre-order) replaced by base kernel  k mod B
  k >= L*B (k' = k-L*B): subexpression floor(k' / (2B)) (root first, then pre-order) extended by
                         (operator floor((k' mod 2B) / B), base kernel k' mod B)          (:97-142)
"""
from typing import List

import numpy as np

from ...configs.kernels.elementary_configs import LinearWithPriorConfig, PeriodicWithPriorConfig, RBFWithPriorConfig, RQWithPriorConfig
from ..kernel_factory import KernelFactory
from .kernel_grammar import BaseKernelGrammarExpression, ElementaryKernelGrammarExpression, KernelGrammarExpression, KernelGrammarOperator


class BaseKernelGrammarSearchSpace:
    def __init__(self, input_dimension: int):
        self.input_dimension = input_dimension
        self.base_kernel_config_list = None
        self.operator_list = None
        self.name = None

    def _per_dimension(self, config_classes, input_dimension):
        self.input_dimension = input_dimension
        self.base_kernel_config_list = [
            cls(input_dimension=input_dimension, active_on_single_dimension=True, active_dimension=i)
            for cls in config_classes
            for i in range(input_dimension)
        ]
        self.operator_list = [KernelGrammarOperator.ADD, KernelGrammarOperator.MULTIPLY]
        self.name = self.__class__.__name__

    def get_root_expressions(self) -> List[ElementaryKernelGrammarExpression]:
        roots = []
        for cfg in self.base_kernel_config_list:
            root = ElementaryKernelGrammarExpression(KernelFactory.build(cfg))
            root.set_generator_name(self.name)
            roots.append(root)
        return roots

    def get_num_base_kernels(self) -> int:
        return len(self.base_kernel_config_list)

    def get_num_neighbour_expressions(self, grammar_expression: BaseKernelGrammarExpression) -> int:
        n_leaves = len(grammar_expression.get_indexes_of_elementary_expressions())
        n_sub = len(grammar_expression.get_indexes_of_subexpression())
        B = len(self.base_kernel_config_list)
        return n_leaves * B + n_sub * len(self.operator_list) * B

    def get_neighbour_at_index(self, grammar_expression: BaseKernelGrammarExpression, index: int) -> BaseKernelGrammarExpression:
        leaves = grammar_expression.get_indexes_of_elementary_expressions()
        subs = grammar_expression.get_indexes_of_subexpression()
        B, n_ops = len(self.base_kernel_config_list), len(self.operator_list)
        assert index < len(leaves) * B + len(subs) * n_ops * B
        new_expression = grammar_expression.deep_copy()
        is_leaf_root = isinstance(grammar_expression, ElementaryKernelGrammarExpression)
        if index < len(leaves) * B:
            kernel = KernelFactory.build(self.base_kernel_config_list[index % B])
            if is_leaf_root:
                new_expression.set_kernel(kernel)
            else:
                new_expression.change_elementary_expression_at_index(leaves[index // B], kernel)
        else:
            rest = index - len(leaves) * B
            where = subs[rest // (n_ops * B)]
            operator = self.operator_list[(rest % (n_ops * B)) // B]
            kernel = KernelFactory.build(self.base_kernel_config_list[rest % B])
            if is_leaf_root:
                new_expression = KernelGrammarExpression(new_expression, ElementaryKernelGrammarExpression(kernel), operator)
            else:
                new_expression.extend_sub_expression_at_index(where, operator, kernel)
        new_expression.set_generator_name(self.name)
        return new_expression

    def get_neighbour_expressions(self, grammar_expression: BaseKernelGrammarExpression) -> List[BaseKernelGrammarExpression]:
        return [self.get_neighbour_at_index(grammar_expression, k) for k in range(self.get_num_neighbour_expressions(grammar_expression))]

    def get_random_neighbour_expression(self, grammar_expression: BaseKernelGrammarExpression) -> BaseKernelGrammarExpression:
        k = np.random.choice(self.get_num_neighbour_expressions(grammar_expression))
        return self.get_neighbour_at_index(grammar_expression, k)

    def random_walk(self, length: int, initial_expression: BaseKernelGrammarExpression) -> List[BaseKernelGrammarExpression]:
        visited, current = [], initial_expression
        for _ in range(length):
            current = self.get_random_neighbour_expression(current)
            visited.append(current)
        return visited

    def check_expression_equality(self, expression1, expression2) -> bool:
        """Equivalence under the rules of addition and multiplication (commutative, associative,
        distributive), as the reference's ``get_add_mult_invariant_hash`` decides it
        (kernel_grammar.py:472-489: leaf hash, ADD -> h1 + h2, MULTIPLY -> h1 * h2 in unbounded
        Python integers).  Here every symbol gets a fixed 61-bit integer instead of a salted hash."""
        def value(e):
            if isinstance(e, ElementaryKernelGrammarExpression):
                return _symbol_value(e.get_name())
            a, b = value(e.expression1), value(e.expression2)
            return a + b if e.operator == KernelGrammarOperator.ADD else a * b

        return value(expression1) == value(expression2)


_SYMBOL_VALUES = {}


def _symbol_value(name: str) -> int:
    if name not in _SYMBOL_VALUES:
        import hashlib

        _SYMBOL_VALUES[name] = int.from_bytes(hashlib.sha256(name.encode()).digest()[:8], "little") >> 3 | 1
    return _SYMBOL_VALUES[name]


class CompositionalKernelSearchSpace(BaseKernelGrammarSearchSpace):
    def __init__(self, input_dimension: int):
        self._per_dimension([RBFWithPriorConfig, LinearWithPriorConfig, PeriodicWithPriorConfig], input_dimension)


class CKSWithRQSearchSpace(BaseKernelGrammarSearchSpace):
    def __init__(self, input_dimension: int):
        self._per_dimension([RBFWithPriorConfig, LinearWithPriorConfig, PeriodicWithPriorConfig, RQWithPriorConfig], input_dimension)


class CKSHighDimSearchSpace(BaseKernelGrammarSearchSpace):
    def __init__(self, input_dimension: int):
        self._per_dimension([RBFWithPriorConfig, RQWithPriorConfig], input_dimension)


class NDimFullKernelsSearchSpace(BaseKernelGrammarSearchSpace):
    def __init__(self, input_dimension: int):
        self.input_dimension = input_dimension
        self.base_kernel_config_list = [
            RBFWithPriorConfig(input_dimension=input_dimension),
            LinearWithPriorConfig(input_dimension=input_dimension),
            PeriodicWithPriorConfig(input_dimension=input_dimension),
        ]
        self.operator_list = [KernelGrammarOperator.ADD, KernelGrammarOperator.MULTIPLY]
        self.name = self.__class__.__name__
