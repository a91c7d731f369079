"""Kernel-grammar expression trees: host-side mirror of the reference's object model.

Same class and method names as ``bosot/kernels/kernel_grammar/kernel_grammar.py`` so that callers
(search spaces, candidate generators, the BO loop) work unchanged, but the objects are light:
a leaf wraps any object with a ``name`` (``<ConfigName>_on_<dim>``) and ``get_input_dimension()``;
nothing from GPflow is needed.  Everything the reference computes by recursive hashing on these
objects (``get_hash`` :452-470, ``get_subtree_dict`` :599-610, ``get_elementary_path_dict``
:683-696, ``get_elementary_count_dict`` :660-671) is computed on the GPU from the flat encoding
(``bosot_b200.encoding``); the dict-returning methods below are thin adapters over that path for
single-expression use, keyed by the 64-bit canonical class id instead of Python's salted hash().
"""
from __future__ import annotations

import copy
from abc import ABC, abstractmethod
from enum import Enum
from typing import Dict, List, Optional


class KernelGrammarOperator(Enum):
    ADD = 1
    MULTIPLY = 2


class SymbolicKernel:
    """Stand-in for an elementary GPflow kernel: only its identity matters on this path
    (``BaseElementaryKernel`` names, bosot/kernels/base_elementary_kernel.py:34-35)."""

    def __init__(self, name: str, input_dimension: int):
        self.name = name
        self.input_dimension = int(input_dimension)

    def get_input_dimension(self) -> int:
        return self.input_dimension

    def __add__(self, other):
        return SymbolicKernel("(" + self.name + " + " + other.name + ")", self.input_dimension)

    def __mul__(self, other):
        return SymbolicKernel("(" + self.name + " * " + other.name + ")", self.input_dimension)

    def __repr__(self):
        return "SymbolicKernel(%s)" % self.name


class BaseKernelGrammarExpression(ABC):
    """Symbolic expression for a kernel: a base kernel or a binary tree of ADD / MULTIPLY nodes
    (reference :44-180)."""

    generator_name: Optional[str] = None

    @abstractmethod
    def get_kernel(self):
        raise NotImplementedError

    @abstractmethod
    def deep_copy(self) -> "BaseKernelGrammarExpression":
        raise NotImplementedError

    @abstractmethod
    def count_elementary_expressions(self) -> int:
        raise NotImplementedError

    @abstractmethod
    def count_operators(self) -> int:
        raise NotImplementedError

    @abstractmethod
    def get_name(self) -> str:
        raise NotImplementedError

    @abstractmethod
    def get_input_dimension(self) -> int:
        raise NotImplementedError

    @abstractmethod
    def get_operator(self) -> Optional[KernelGrammarOperator]:
        raise NotImplementedError

    @abstractmethod
    def get_indexes_of_subexpression(self) -> List[List[int]]:
        raise NotImplementedError

    @abstractmethod
    def get_indexes_of_elementary_expressions(self) -> List[List[int]]:
        raise NotImplementedError

    @abstractmethod
    def get_expression_at_index(self, index_list: List[int]) -> "BaseKernelGrammarExpression":
        raise NotImplementedError

    def set_generator_name(self, generator_name: str) -> None:
        self.generator_name = generator_name

    def get_generator_name(self) -> Optional[str]:
        return self.generator_name

    def render_pickable(self) -> None:
        return None

    def __str__(self) -> str:
        return self.get_name()

    # ---- feature dictionaries, computed on the device (single-expression convenience) ----
    def _device_features(self):
        import torch

        from ... import sot
        from ...encoding import Grammar, encode

        grammar = Grammar.get(self.get_generator_name(), self.get_input_dimension())
        trees = sot.as_device_trees(encode([self], grammar), torch.device("cuda", torch.cuda.current_device()))
        return grammar, {k: v[0].cpu().numpy() for k, v in sot.featurize(trees, grammar).items()}

    def get_subtree_dict(self) -> Dict[int, List]:
        """class id -> [count, None], insertion order = pre-order first occurrence (reference :599-610)."""
        _, f = self._device_features()
        return {int(f["sub_id"][k]) & 0xFFFFFFFFFFFFFFFF: [int(f["sub_cnt"][k]), None] for k in range(int(f["sub_n"]))}

    def get_elementary_path_dict(self, operators_to_reduce=None) -> Dict[int, List]:
        """(symbol * 64 + reduced path code) -> [count, None] (reference :683-696; both operators are
        reduced, as the kernel-kernel asks, kernel_kernel_grammar_tree.py:178-180)."""
        if operators_to_reduce is not None and set(operators_to_reduce) != {KernelGrammarOperator.ADD, KernelGrammarOperator.MULTIPLY}:
            raise NotImplementedError("the device featuriser reduces both operators")
        _, f = self._device_features()
        return {int(f["path_id"][k]) & 0xFFFF: [int(f["path_cnt"][k]), None] for k in range(int(f["path_n"]))}

    def get_elementary_count_dict(self) -> Dict[str, int]:
        """leaf name -> count (reference :318-319, :660-671), in symbol-table order."""
        g, f = self._device_features()
        return {g.symbols[s]: int(c) for s, c in enumerate(f["sym_cnt"]) if c}


class ElementaryKernelGrammarExpression(BaseKernelGrammarExpression):
    def __init__(self, kernel):
        if not hasattr(kernel, "name"):
            raise TypeError("an elementary expression wraps a kernel object with a `name`")
        self.kernel = kernel
        self.generator_name = None

    def set_kernel(self, kernel) -> None:
        self.kernel = kernel

    def deep_copy(self):
        return ElementaryKernelGrammarExpression(copy.copy(self.kernel))

    def count_elementary_expressions(self) -> int:
        return 1

    def count_operators(self) -> int:
        return 0

    def get_input_dimension(self) -> int:
        return self.kernel.get_input_dimension()

    def get_name(self) -> str:
        return self.kernel.name

    def get_operator(self):
        return None

    def get_kernel(self):
        return self.kernel

    def _index_walk(self) -> List[List[Optional[int]]]:
        return [[None]]

    def get_indexes_of_subexpression(self) -> List[List[int]]:
        return [[-1]]

    def get_indexes_of_elementary_expressions(self) -> List[List[int]]:
        return [[-1]]

    def get_expression_at_index(self, index_list: List[int]):
        if list(index_list) != [-1]:
            raise IndexError("an elementary expression only has the root index [-1]")
        return self


class KernelGrammarExpression(BaseKernelGrammarExpression):
    def __init__(self, expression1: BaseKernelGrammarExpression, expression2: BaseKernelGrammarExpression, operator: KernelGrammarOperator):
        self.expression1 = expression1
        self.expression2 = expression2
        self.operator = operator
        self.generator_name = None

    def get_kernel(self):
        k1, k2 = self.expression1.get_kernel(), self.expression2.get_kernel()
        return k1 + k2 if self.operator == KernelGrammarOperator.ADD else k1 * k2

    def get_operator(self):
        return self.operator

    def get_left_expression(self):
        return self.expression1

    def get_right_expression(self):
        return self.expression2

    def deep_copy(self):
        # like the reference (:353-357) the copy does NOT carry generator_name
        return KernelGrammarExpression(self.expression1.deep_copy(), self.expression2.deep_copy(), self.operator)

    def get_input_dimension(self) -> int:
        d = self.expression1.get_input_dimension()
        assert d == self.expression2.get_input_dimension()
        return d

    def count_elementary_expressions(self) -> int:
        return self.expression1.count_elementary_expressions() + self.expression2.count_elementary_expressions()

    def count_operators(self) -> int:
        return self.expression1.count_operators() + self.expression2.count_operators() + 1

    def get_name(self) -> str:
        return "(" + self.expression1.get_name() + " " + self.operator.name + " " + self.expression2.get_name() + ")"

    # ---- tree addressing: index lists [0, 1, 1, ...] (0 = expression1, 1 = expression2), [-1] = root ----
    def _index_walk(self) -> List[List[Optional[int]]]:
        """Every node below the root in the reference's enumeration order (child 0, its subtree,
        child 1, its subtree); a trailing None marks 'this child is elementary' (reference :391-402)."""
        out: List[List[Optional[int]]] = []
        for side, child in ((0, self.expression1), (1, self.expression2)):
            out.append([side])
            out.extend([side] + tail for tail in child._index_walk())
        return out

    def get_indexes_of_subexpression(self) -> List[List[int]]:
        return [[-1]] + [ix for ix in self._index_walk() if ix[-1] is not None]

    def get_indexes_of_elementary_expressions(self) -> List[List[int]]:
        return [ix[:-1] for ix in self._index_walk() if ix[-1] is None]

    def _descend(self, index_list: List[int]):
        """(parent node, side) addressed by index_list (len >= 1, no -1)."""
        node = self
        for side in index_list[:-1]:
            node = node.expression1 if side == 0 else node.expression2
            assert isinstance(node, KernelGrammarExpression)
        return node, index_list[-1]

    def get_expression_at_index(self, index_list: List[int]):
        if list(index_list) == [-1]:
            return self
        node, side = self._descend(list(index_list))
        return node.expression1 if side == 0 else node.expression2

    # ---- mutations used by the search spaces (reference :491-542) ----
    def change_elementary_expression_at_index(self, index_list: List[int], new_kernel) -> None:
        node, side = self._descend(list(index_list))
        target = node.expression1 if side == 0 else node.expression2
        assert isinstance(target, ElementaryKernelGrammarExpression)
        if side == 0:
            node.expression1 = ElementaryKernelGrammarExpression(new_kernel)
        else:
            node.expression2 = ElementaryKernelGrammarExpression(new_kernel)

    def extend_sub_expression_at_index(self, index_list: List[int], operator: KernelGrammarOperator, new_kernel) -> None:
        new_leaf = ElementaryKernelGrammarExpression(new_kernel)
        if list(index_list) == [-1]:
            self.expression1 = self.deep_copy()
            self.expression2 = new_leaf
            self.operator = operator
            return
        node, side = self._descend(list(index_list))
        if side == 0:
            node.expression1 = KernelGrammarExpression(node.expression1, new_leaf, operator)
        else:
            node.expression2 = KernelGrammarExpression(node.expression2, new_leaf, operator)

    # ---- sum-of-products normal form (reference :708-754) ----
    def push_down_one_mult(self) -> bool:
        """Distribute ONE multiplication over an addition directly below it; True if something changed."""
        if self.operator == KernelGrammarOperator.MULTIPLY:
            left, right = self.expression1, self.expression2
            if left.get_operator() == KernelGrammarOperator.ADD:
                self.operator = KernelGrammarOperator.ADD
                self.expression1 = KernelGrammarExpression(left.expression1, right, KernelGrammarOperator.MULTIPLY)
                self.expression2 = KernelGrammarExpression(left.expression2, right, KernelGrammarOperator.MULTIPLY)
                return True
            if right.get_operator() == KernelGrammarOperator.ADD:
                self.operator = KernelGrammarOperator.ADD
                self.expression1 = KernelGrammarExpression(right.expression1, left, KernelGrammarOperator.MULTIPLY)
                self.expression2 = KernelGrammarExpression(right.expression2, left, KernelGrammarOperator.MULTIPLY)
                return True
        for child in (self.expression1, self.expression2):
            if isinstance(child, KernelGrammarExpression) and child.push_down_one_mult():
                return True
        return False


class KernelGrammarExpressionTransformer:
    @staticmethod
    def transform_to_normal_form(expression: BaseKernelGrammarExpression) -> BaseKernelGrammarExpression:
        new_expression = expression.deep_copy()
        new_expression.set_generator_name(expression.get_generator_name())
        if isinstance(new_expression, KernelGrammarExpression):
            while new_expression.push_down_one_mult():
                pass
        return new_expression
