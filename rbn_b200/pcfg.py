"""Discrete RBN bricks and the PCFG front-end (API of /root/reference/rbnet/pcfg.py).

Constructor signatures, attribute names and therefore `named_parameters()` / `state_dict()` keys are those of the
reference (`_cells.0.transition_probabilities.log_p`, `_cells.0._transitions.{0,1}...`,
`_prior.structural_distribution.log_p`, `_prior.prior_distributions.0.log_p`), and the same argument validation
raises the same exception types (pcfg.py:260-267, 296-301, 333-341, 370-382).

The per-location hooks (`inside_marginals`, `inside_mixture`, `marginal_likelihood`) are NOT evaluated on the host:
a model made of these bricks is run as one fused CUDA chart pass by `SequentialRBN.inside` (rbn_b200/engine.py).
Calling a hook directly raises `HostEvaluationError` instead of silently computing on the CPU.
"""
from typing import Iterable

import numpy as np
import torch

from .base import Cell, NonTermVar, Prior, Transition
from .sequential import SequentialBinaryTransition, SequentialRBN, SequentialTerminalTransition
from .tmap import TMap
from .util import ConstrainedModuleList, ConstrainedModuleMixin, LogProb, as_detached_tensor, normalize_non_zero


class HostEvaluationError(NotImplementedError):
    """Raised when a discrete brick is asked to evaluate one chart location on the host."""

    def __init__(self, what):
        super().__init__(f"{what} is evaluated inside the fused CUDA chart pass (SequentialRBN.inside); "
                         f"rbn_b200 has no host/CPU implementation of it")


def _non_negative(weights):
    if bool((weights < 0).any()):
        raise ValueError("All weights have to be non-negative")


class DiscreteNonTermVar(NonTermVar):
    """Discrete non-terminal variable with `cardinality` values (pcfg.py:193-238)."""

    def __init__(self, cardinality, chart_type="TMap", *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.cardinality = cardinality
        self.chart_type = chart_type

    def get_chart(self, n, *args, **kwargs):
        """Empty chart for a sequence of length n (pcfg.py:206-217)."""
        if self.chart_type == "TMap":
            return TMap(torch.zeros((TMap.size_from_n(n), self.cardinality)))
        if self.chart_type == "dict":
            return {}
        raise ValueError(f"Unknown chart type '{self.chart_type}'")

    def mixture(self, components, weights=None, dim=0):
        """(Weighted) sum of discrete distributions along `dim` (pcfg.py:219-238); tiny host-side utility kept for
        API compatibility -- the chart pass itself does its mixtures on the GPU."""
        if len(components) == 0:
            return torch.zeros(self.cardinality)
        stacked = components if torch.is_tensor(components) else torch.stack(list(components))
        if weights is not None:
            stacked = torch.as_tensor(weights) * stacked
        return stacked.sum(dim=dim if isinstance(dim, tuple) else (dim,))


class DiscretePrior(Prior, ConstrainedModuleMixin):
    """Prior over which variable sits at the root and over its value (pcfg.py:241-273)."""

    def __init__(self, struc_weights, prior_weights, prob_rep=LogProb, *args, **kwargs):
        super().__init__(*args, **kwargs)
        struc = as_detached_tensor(struc_weights)
        dists = [as_detached_tensor(w) for w in prior_weights]
        if struc.dim() != 1:
            raise ValueError("'struc_weights` must be one-dimensional")
        _non_negative(struc)
        self.structural_distribution = prob_rep(p=struc)
        if len(dists) != len(struc):
            raise ValueError(f"Expected as many distributions as weights, but got: {len(dists)} and {len(struc)}")
        self.prior_distributions = ConstrainedModuleList([prob_rep(p=w) for w in dists])

    def marginal_likelihood(self, root_location, inside_chart, **kwargs):
        raise HostEvaluationError("DiscretePrior.marginal_likelihood")


class DiscreteBinaryNonTerminalTransition(SequentialBinaryTransition, ConstrainedModuleMixin):
    """p(left, right | parent) with weights[left, right, parent] -- parent LAST (pcfg.py:276-314)."""

    def __init__(self, weights, left_idx=0, right_idx=0, prob_rep=LogProb, *args, **kwargs):
        super().__init__(*args, **kwargs)
        weights = as_detached_tensor(weights)
        if weights.dim() != 3:
            raise ValueError("'weights' has to be three-dimensional")
        _non_negative(weights)
        self.transition_probabilities = prob_rep(p=weights, dim=(0, 1))
        self.left_idx = left_idx
        self.right_idx = right_idx

    def inside_marginals(self, location, inside_chart, terminal_chart, **kwargs):
        raise HostEvaluationError("DiscreteBinaryNonTerminalTransition.inside_marginals")


class DiscreteTerminalTransition(SequentialTerminalTransition, ConstrainedModuleMixin):
    """p(terminal | parent) with weights[terminal, parent] (pcfg.py:318-352)."""

    def __init__(self, weights, term_idx=0, prob_rep=LogProb, *args, **kwargs):
        super().__init__(*args, **kwargs)
        weights = as_detached_tensor(weights)
        if weights.dim() != 2:
            raise ValueError("'weights' has to be two-dimensional")
        _non_negative(weights)
        empty = (weights.sum(dim=0) == 0).nonzero().flatten().tolist()
        if empty:
            raise ValueError("All transition weights for the following indices are zero: "
                             + ", ".join(str(int(i)) for i in empty))
        self.transition_probabilities = prob_rep(p=weights, dim=0)
        self.term_idx = term_idx

    def inside_marginals(self, location, inside_chart, terminal_chart, **kwargs):
        raise HostEvaluationError("DiscreteTerminalTransition.inside_marginals")


class DiscreteCell(Cell, ConstrainedModuleMixin):
    """Cell of a discrete variable; the structural distribution p(z | x) may depend on the value x (pcfg.py:356-398)."""

    def __init__(self, variable, weights, transitions, prob_rep=LogProb, *args, **kwargs):
        super().__init__(variable=variable, *args, **kwargs)
        weights = as_detached_tensor(weights)
        if weights.dim() != 2:
            raise ValueError("'weights' must be two-dimensional")
        _non_negative(weights)
        self.transition_probabilities = prob_rep(p=weights, dim=0)
        if len(transitions) != weights.shape[0]:
            raise ValueError(f"Number of transitions and size of first weight dimension must be the same, but got: "
                             f"{len(transitions)} and {weights.shape[0]}")
        self._transitions = ConstrainedModuleList(transitions)
        if variable.cardinality != weights.shape[1]:
            raise ValueError(f"Cardinality of variable and size of second weight dimension must be the same, but got: "
                             f"{variable.cardinality} and {weights.shape[1]}")

    def transitions(self) -> Iterable[Transition]:
        return iter(self._transitions)

    def inside_mixture(self, inside_marginals):
        raise HostEvaluationError("DiscreteCell.inside_mixture")


class StaticCell(DiscreteCell):
    """Cell whose structural distribution does not depend on the variable's value (pcfg.py:401-410)."""

    def __init__(self, variable, weights, transitions, prob_rep=LogProb, *args, **kwargs):
        per_value = as_detached_tensor(weights).unsqueeze(1).expand(-1, variable.cardinality)
        super().__init__(variable=variable, weights=per_value, transitions=transitions, prob_rep=prob_rep,
                         *args, **kwargs)


class PCFG(SequentialRBN):
    """A SequentialRBN over symbol sequences with a tokeniser and a chart pretty-printer (pcfg.py:14-41)."""

    def __init__(self, cells, prior, terminal_indices, non_terminals, auto_tokenise=True, *args, **kwargs):
        super().__init__(cells=cells, prior=prior, *args, **kwargs)
        self.terminal_indices = terminal_indices
        self._non_terminals = non_terminals
        self.auto_tokenise = auto_tokenise

    def tokenise(self, sequence):
        """symbols -> [[index], ...]; unknown symbols raise KeyError (pcfg.py:22-23)."""
        lookup = self.terminal_indices
        return [[lookup[symbol]] for symbol in sequence]

    def init_inside(self, sequence):
        super().init_inside(sequence=self.tokenise(sequence) if self.auto_tokenise else sequence)

    def _prepare_sequence(self, sequence):
        return self.tokenise(sequence) if self.auto_tokenise else sequence

    def map_inside_chart(self, precision=None):
        """Chart of 'Symbol:probability' strings for the entries with positive probability (pcfg.py:30-41)."""
        rows = []
        for probs in self.inside_chart[0].arr:
            entries = [f"{self._non_terminals[k]}:{np.format_float_scientific(float(p), precision=precision)}"
                       for k, p in enumerate(probs) if p > 0]
            rows.append(entries[0] if len(entries) == 1 else entries)
        return TMap(rows)


def _parse_rule(rule):
    """'X --> Y Z' / 'X --> y' strings or (X, (Y, Z)) / (X, (y,)) tuples -> (lhs, rhs) (pcfg.py:75-84)."""
    if not isinstance(rule, str):
        lhs, rhs = rule
        return lhs, tuple(rhs)
    parts = rule.split()
    if len(parts) not in (3, 4) or parts[1] != "-->":
        raise ValueError(f"Rules must be of the form 'X --> Y' or 'X --> Y Z', but got '{rule}'")
    return parts[0], tuple(parts[2:])


class AbstractedPCFG(PCFG, ConstrainedModuleMixin):
    """PCFG as an RBN with ONE discrete non-terminal variable and one discrete terminal variable (pcfg.py:44-108).

    The reference additionally derives from `pytorch_lightning.LightningModule`; that base contributes nothing to
    the inside path and is not required here -- the class is a plain `torch.nn.Module` with
    the same parameter names.
    """

    def __init__(self, non_terminals, terminals, rules, start, prob_rep=LogProb, *args, **kwargs):
        self.terminals = terminals
        self.non_terminal_indices = {symbol: k for k, symbol in enumerate(non_terminals)}
        terminal_indices = {symbol: k for k, symbol in enumerate(terminals)}
        n_nt, n_t = len(non_terminals), len(terminals)

        binary = np.zeros((n_nt, n_nt, n_nt))      # [left, right, parent]
        emission = np.zeros((n_t, n_nt))            # [terminal, parent]
        for rule, weight in rules:
            lhs, rhs = _parse_rule(rule)
            parent = self.non_terminal_indices[lhs]
            if len(rhs) == 1:
                emission[terminal_indices[rhs[0]], parent] = weight
            elif len(rhs) == 2:
                binary[self.non_terminal_indices[rhs[0]], self.non_terminal_indices[rhs[1]], parent] = weight
            else:
                raise ValueError(f"Right-hand side of rules must consist of one (terminal) or two (non-terminal) "
                                 f"symbols; this one has {len(rhs)} symbols: rhs={rhs}, lhs={lhs}, w={weight}")
        # structural distribution per parent: row 0 = terminal mass, row 1 = binary mass (pcfg.py:97, 106)
        structural = [emission.sum(axis=0), binary.sum(axis=(0, 1))]
        # parents with only one kind of rule get a uniform dummy for the other kind; its structural weight is 0
        emission = normalize_non_zero(emission, axis=0, make_zeros_uniform=True)
        binary = normalize_non_zero(binary, axis=(0, 1), make_zeros_uniform=True)

        start_one_hot = np.zeros(n_nt)
        start_one_hot[self.non_terminal_indices[start]] = 1
        prior = DiscretePrior(struc_weights=np.ones(1), prior_weights=[start_one_hot], prob_rep=prob_rep)
        cell = DiscreteCell(variable=DiscreteNonTermVar(n_nt), weights=structural,
                            transitions=[DiscreteTerminalTransition(weights=emission, prob_rep=prob_rep),
                                         DiscreteBinaryNonTerminalTransition(weights=binary, prob_rep=prob_rep)],
                            prob_rep=prob_rep)
        super().__init__(cells=[cell], prior=prior, non_terminals=non_terminals, terminal_indices=terminal_indices,
                         *args, **kwargs)
