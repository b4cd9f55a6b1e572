"""Minimal stand-ins for `cobaya.theory` so that the reference's eftpipe/theory.py imports (golden generation only:
its PlkInterpolator class is exercised, not the Cobaya Theory protocol)."""
from .log import HasLogger


class Theory(HasLogger):
    def initialize(self):
        pass


class HelperTheory(Theory):
    pass


class Provider:
    pass
