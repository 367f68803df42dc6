"""Abstract RBN template and brick interfaces (API of /root/reference/rbnet/base.py).

`RBN` declares the hooks an RBN instantiation provides (schedule, cells, chart access, prior) and offers the
generic template method `inside()` (base.py:93-121).  The concrete discrete instantiation in this package does NOT
run that Python loop nest: `rbn_b200.sequential.SequentialRBN.inside` hands the whole chart pass to the CUDA
kernels.  The template is kept because it is the reference's plugin surface (tests/test_base.py:22-45 pins that
every abstract hook raises NotImplementedError and that `init_inside()` returns None).
"""
from abc import ABC, abstractmethod
from typing import Iterable

import torch


class RBN(ABC):
    """Base class for recursive Bayesian networks."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    @abstractmethod
    def inside_schedule(self, *args, **kwargs):
        """Yield (batches of) non-terminal locations such that every location comes after the locations it is
        generated into, i.e. children before parents (base.py:13-23)."""
        raise NotImplementedError

    @property
    @abstractmethod
    def root_location(self):
        """Location of the root variables (base.py:25-34)."""
        raise NotImplementedError

    @abstractmethod
    def cells(self):
        """Iterable over the cells, one per non-terminal template variable (base.py:36-41)."""
        raise NotImplementedError

    def init_inside(self, *args, **kwargs):
        """Called by `inside` with all its arguments before parsing a new input (base.py:43-48)."""
        return None

    @abstractmethod
    def update_inside_chart(self, var_idx, locations, values):
        """Store `values` for variable `var_idx` at `locations` (base.py:50-59)."""
        raise NotImplementedError

    @property
    @abstractmethod
    def inside_chart(self):
        """Charts of inside probabilities, one per variable (base.py:61-69)."""
        raise NotImplementedError

    @property
    @abstractmethod
    def terminal_chart(self):
        """Chart of terminal variables (base.py:71-79)."""
        raise NotImplementedError

    @property
    @abstractmethod
    def prior(self):
        """Prior transition implementing `Prior.marginal_likelihood` (base.py:81-91)."""
        raise NotImplementedError

    def inside(self, *args, **kwargs):
        """Template method: run the hooks bottom-up and return the marginal likelihood (base.py:93-121).

        Instantiations built from custom bricks run through here; the discrete bricks of `rbn_b200.pcfg` never do
        (see `SequentialRBN.inside`)."""
        self.init_inside(*args, **kwargs)
        for location in self.inside_schedule():
            for var_idx, cell in enumerate(self.cells()):
                marginals = [tr.inside_marginals(location=location, inside_chart=self.inside_chart,
                                                 terminal_chart=self.terminal_chart) for tr in cell.transitions()]
                self.update_inside_chart(var_idx=var_idx, locations=location, values=cell.inside_mixture(marginals))
        return self.prior.marginal_likelihood(root_location=self.root_location, inside_chart=self.inside_chart)


class Transition(ABC, torch.nn.Module):
    """A transition of an RBN cell; provides the marginals over inside probabilities for every split (base.py:124-165)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    @abstractmethod
    def inside_marginals(self, location, inside_chart, terminal_chart, **kwargs):
        raise NotImplementedError


class Prior(ABC, torch.nn.Module):
    """Prior transition: no parent, cannot emit terminals; provides the marginal likelihood (base.py:168-191)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    @abstractmethod
    def marginal_likelihood(self, root_location, inside_chart, **kwargs):
        raise NotImplementedError


class NonTermVar(ABC):
    """A type of non-terminal template variable: makes charts, computes mixtures (base.py:194-219)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    @abstractmethod
    def get_chart(self, *args, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def mixture(self, *args, **kwargs):
        raise NotImplementedError


class Cell(ABC, torch.nn.Module):
    """The transitions available from one non-terminal variable and their mixture (base.py:222-265)."""

    def __init__(self, variable, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.variable = variable

    @abstractmethod
    def transitions(self) -> Iterable[Transition]:
        raise NotImplementedError

    @abstractmethod
    def inside_mixture(self, inside_marginals):
        raise NotImplementedError
